#include "DeviceGroup.hpp"
#include <algorithm>
#include <cstring>

using namespace KITGPI;

namespace
{
    void check(int rc)
    {
        if (rc != WS_OK)
            COMMON_THROWEXCEPTION(ws_last_error())
    }
}

ForwardSolver::DeviceGroup::DeviceGroup(std::vector<IndexType> const &devs) : devices(devs)
{
    SCAI_ASSERT_ERROR(!devices.empty(), "a shot domain needs at least one GPU")
    handles.assign(devices.size(), nullptr);
    y0.assign(devices.size(), 0);
    nyl.assign(devices.size(), 0);
    errors.assign(devices.size(), "");
    if (devices.size() > 1)
        for (IndexType r = 0; r < size(); r++)
            workers.emplace_back(&DeviceGroup::workerLoop, this, r);
}

ForwardSolver::DeviceGroup::~DeviceGroup()
{
    *alive = false;
    destroy();
    {
        std::lock_guard<std::mutex> lock(m);
        stop = true;
    }
    cvStart.notify_all();
    for (auto &t : workers)
        t.join();
}

void ForwardSolver::DeviceGroup::workerLoop(IndexType rank)
{
    unsigned long seen = 0;
    for (;;) {
        std::function<void(IndexType)> const *fn = nullptr;
        {
            std::unique_lock<std::mutex> lock(m);
            cvStart.wait(lock, [&] { return stop || generation != seen; });
            if (stop)
                return;
            seen = generation;
            fn = task;
        }
        std::string err;
        try {
            (*fn)(rank);
        } catch (std::exception const &e) {
            err = e.what();
            if (err.empty())
                err = "unknown error";
        }
        {
            std::lock_guard<std::mutex> lock(m);
            errors[rank] = err;
            if (--pending == 0)
                cvDone.notify_all();
        }
    }
}

void ForwardSolver::DeviceGroup::forEach(std::function<void(IndexType)> const &fn)
{
    if (devices.size() == 1) {
        fn(0);
        return;
    }
    {
        std::unique_lock<std::mutex> lock(m);
        task = &fn;
        pending = size();
        generation++;
        cvStart.notify_all();
        cvDone.wait(lock, [&] { return pending == 0; });
        task = nullptr;
    }
    for (IndexType r = 0; r < size(); r++)
        if (!errors[r].empty())
            COMMON_THROWEXCEPTION("GPU " << devices[r] << " (slab " << r << " of " << size() << "): " << errors[r])
}

void ForwardSolver::DeviceGroup::create(ws_desc desc)
{
    destroy();
    unsigned char id[128];
    std::memset(id, 0, sizeof(id));
    if (size() > 1)
        check(ws_comm_unique_id(id));
    const IndexType nz = desc.dim == 3 ? desc.nz : 1;
    planeSize = (size_t)desc.nx * nz;
    nGlobal = planeSize * desc.ny;
    forEach([&](IndexType r) {
        ws_desc d = desc;
        d.rank = r;
        d.nranks = size();
        d.device = devices[r];
        check(ws_create(&d, &handles[r]));
        if (size() > 1)
            check(ws_comm_init(handles[r], id));
        int32_t a = 0, b = 0;
        check(ws_local_range(handles[r], &a, &b));
        y0[r] = a;
        nyl[r] = b;
    });
}

void ForwardSolver::DeviceGroup::createSparse(ws_desc desc, size_t nPoints)
{
    destroy();
    // the reference partitions irregular grids with a graph partitioner; here a shot on an irregular grid runs on one GPU
    SCAI_ASSERT_ERROR(size() == 1, "variable grids run on one GPU per shot domain (set GPUsPerShotDomain=1)")
    desc.rank = 0;
    desc.nranks = 1;
    desc.device = devices[0];
    check(ws_create_sparse(&desc, (int64_t)nPoints, &handles[0]));
    planeSize = nPoints;
    nGlobal = nPoints;
    y0[0] = 0;
    nyl[0] = 1;
}

void ForwardSolver::DeviceGroup::destroy()
{
    bool any = false;
    for (auto h : handles)
        any = any || h != nullptr;
    if (!any)
        return;
    for (auto *w : liveSets) // stored wavefield objects whose owner is still around: their memory goes with the solver
        ws_wavefields_destroy(w);
    liveSets.clear();
    forEach([&](IndexType r) {
        if (handles[r])
            ws_destroy(handles[r]);
        handles[r] = nullptr;
    });
}

std::vector<float> ForwardSolver::DeviceGroup::getWavefield(std::string const &component)
{
    std::vector<float> out(nGlobal);
    // "CURL" / "DIV" refresh ghost planes over the communicator: a collective call
    forEach([&](IndexType r) { check(ws_get_wavefield(handles[r], component.c_str(), out.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
    return out;
}

void ForwardSolver::DeviceGroup::setWavefield(std::string const &component, std::vector<float> const &values)
{
    SCAI_ASSERT_ERROR(values.size() == nGlobal, "wavefield vector must hold NX*NY*NZ values")
    forEach([&](IndexType r) { check(ws_set_wavefield(handles[r], component.c_str(), values.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
}

std::vector<float> ForwardSolver::DeviceGroup::getMaterial(std::string const &name)
{
    std::vector<float> out(nGlobal);
    forEach([&](IndexType r) { check(ws_get_material(handles[r], name.c_str(), out.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
    return out;
}

void ForwardSolver::DeviceGroup::getSeismogram(std::vector<float> &all)
{
    // rows of receivers that live on another slab are left untouched by ws_get_seismogram, so the ranks fill one buffer in turn
    std::fill(all.begin(), all.end(), 0.0f);
    for (IndexType r = 0; r < size(); r++)
        check(ws_get_seismogram(handles[r], all.data()));
}

ForwardSolver::DeviceGroup::FieldSet ForwardSolver::DeviceGroup::createFieldSet()
{
    FieldSet set(size(), nullptr);
    forEach([&](IndexType r) { check(ws_wavefields_create(handles[r], &set[r])); });
    liveSets.insert(liveSets.end(), set.begin(), set.end());
    return set;
}

void ForwardSolver::DeviceGroup::destroyFieldSet(FieldSet &set)
{
    for (auto *w : set) {
        auto it = std::find(liveSets.begin(), liveSets.end(), w);
        if (it == liveSets.end())
            continue; // already released with the solver
        liveSets.erase(it);
        ws_wavefields_destroy(w);
    }
    set.clear();
}

void ForwardSolver::DeviceGroup::fieldSetBinary(FieldSet const *dst, FieldSet const *src, int op)
{
    forEach([&](IndexType r) {
        ws_wavefields *d = dst ? (*dst)[r] : nullptr;
        const ws_wavefields *s = src ? (*src)[r] : nullptr;
        check(op == 0 ? ws_wavefields_assign(handles[r], d, s) : (op == 1 ? ws_wavefields_plus_assign(handles[r], d, s) : ws_wavefields_minus_assign(handles[r], d, s)));
    });
}

void ForwardSolver::DeviceGroup::fieldSetScale(FieldSet const *dst, float rhs)
{
    forEach([&](IndexType r) { check(ws_wavefields_times_assign(handles[r], dst ? (*dst)[r] : nullptr, rhs)); });
}

void ForwardSolver::DeviceGroup::fieldSetScale(FieldSet const *dst, std::vector<float> const &rhs)
{
    SCAI_ASSERT_ERROR(rhs.size() == nGlobal, "the vector must hold NX*NY*NZ values")
    forEach([&](IndexType r) { check(ws_wavefields_times_assign_vector(handles[r], dst ? (*dst)[r] : nullptr, rhs.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
}

std::vector<float> ForwardSolver::DeviceGroup::getWavefield(FieldSet const &set, std::string const &component)
{
    std::vector<float> out(nGlobal);
    forEach([&](IndexType r) { check(ws_wavefields_get(handles[r], set[r], component.c_str(), out.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
    return out;
}

void ForwardSolver::DeviceGroup::setStepScaling(std::vector<float> const &vec)
{
    SCAI_ASSERT_ERROR(vec.empty() || vec.size() == nGlobal, "the vector must hold NX*NY*NZ values")
    forEach([&](IndexType r) { check(ws_set_step_scaling(handles[r], vec.empty() ? nullptr : vec.data() + (size_t)y0[r] * planeSize, (size_t)nyl[r] * planeSize)); });
}

bool ForwardSolver::DeviceGroup::isFinite()
{
    std::vector<int32_t> flags(size(), 0);
    forEach([&](IndexType r) { check(ws_is_finite(handles[r], &flags[r])); }); // all-reduced inside the library
    bool ok = true;
    for (auto f : flags)
        ok = ok && f != 0;
    return ok;
}

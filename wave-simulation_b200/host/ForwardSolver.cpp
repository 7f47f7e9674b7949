#include "ForwardSolver.hpp"
#include "IrregularGrid.hpp"
#include "../../include/wavesim.h"
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
    int eqCode(std::string const &t)
    {
        const char *names[] = {"acoustic", "elastic", "viscoelastic", "sh", "viscosh", "tmem", "emem", "viscotmem", "viscoemem"};
        for (int k = 0; k < 9; k++)
            if (t == names[k])
                return k;
        COMMON_THROWEXCEPTION("Unkown type")
    }
    ws_desc makeDesc(Configuration::Configuration const &config, std::string const &dimension, std::string const &type, int device)
    {
        ws_desc d;
        std::memset(&d, 0, sizeof(d));
        d.dim = dimension == "3d" ? 3 : 2;
        d.eq = eqCode(type);
        d.nx = config.get<IndexType>("NX");
        d.ny = config.get<IndexType>("NY");
        d.nz = d.dim == 3 ? config.get<IndexType>("NZ") : 1;
        d.dh = config.get<ValueType>("DH");
        d.dt = config.get<ValueType>("DT");
        d.nt = Common::time2index(config.get<ValueType>("T"), d.dt); // Simulation.cpp:304
        d.fd_order = config.get<IndexType>("spatialFDorder");
        d.edge_policy = config.getAndCatch("edgePolicy", config.getAndCatch("useStencilMatrix", 0) ? 0 : 1);
        // 0 = off, 1 = image method, 2 = improved vacuum formulation: plain operators (the vacuum is part of the model: averaged shear moduli
        // below 4 Pa become 0, ModelparameterSeismic.cpp:430) and, as with 1, no absorbing frame at the top (ABS*::init / CPML*::init test == 0)
        d.free_surface = config.get<IndexType>("FreeSurface");
        SCAI_ASSERT_ERROR(d.free_surface >= 0 && d.free_surface <= 2, "FreeSurface must be 0, 1 or 2")
        d.damping = config.get<IndexType>("DampingBoundary");
        d.boundary_width = config.getAndCatch("BoundaryWidth", 0);
        d.damping_coeff = config.getAndCatch("DampingCoeff", ValueType(0));
        d.vmax_cpml = config.getAndCatch("VMaxCPML", ValueType(0));
        d.fc_cpml = config.getAndCatch("CenterFrequencyCPML", ValueType(0));
        d.npower = config.getAndCatch("NPower", ValueType(0));
        const bool visco = type.compare(0, 5, "visco") == 0;
        d.n_relax = visco ? config.get<IndexType>("numRelaxationMechanisms") : 0;
        const char *keys[4] = {"relaxationFrequency", "relaxationFrequency2", "relaxationFrequency3", "relaxationFrequency4"};
        for (int l = 0; l < d.n_relax && l < 4; l++)
            d.relax_freq[l] = config.get<ValueType>(keys[l]);
        d.exact_arith = config.getAndCatch("exactArithmetic", 0); // B200 extension: reference operation order, no FMA contraction
        d.kernel_variant = config.getAndCatch("kernelVariant", 0);
        d.rank = 0;
        d.nranks = 1;
        d.device = device;
        return d;
    }
}

template <typename ValueType> ForwardSolver::ForwardSolver<ValueType>::ForwardSolver(std::string const &dim, std::string const &type) : dimension(dim), equationType(type)
{
    SCAI_ASSERT_ERROR(dimension == "2d" || dimension == "3d", "Unkown dimension")
    const int eq = eqCode(type);
    // ForwardSolverFactory.cpp:4-66: sh, viscosh, tmem and viscotmem exist in 2D only
    if (dimension == "3d" && (eq == WS_EQ_SH || eq == WS_EQ_VISCOSH || eq == WS_EQ_TMEM || eq == WS_EQ_VISCOTMEM))
        COMMON_THROWEXCEPTION("Unkown type")
}

template <typename ValueType> ForwardSolver::ForwardSolver<ValueType>::~ForwardSolver() {}

template <typename ValueType>
ValueType ForwardSolver::ForwardSolver<ValueType>::estimateMemory(Configuration::Configuration const &config, Acquisition::Coordinates<ValueType> const &modelCoordinates)
{
    ws_desc d = makeDesc(config, dimension, equationType, deviceIds[0]);
    if (modelCoordinates.isVariable()) {
        // operator-given mode: wavefields, model vectors and scratch as plain vectors of N points, the 2 x dim operators in ELL form
        // (column + value per tap, as many taps as the highest FD order), CPML / interpolation lists are O(surface)
        const double N = (double)modelCoordinates.getNGridpoints();
        const int dim = d.dim;
        IndexType taps = d.fd_order;
        if (config.getAndCatch("useVariableFDoperators", 0) != 0)
            for (IndexType o : Acquisition::readColumnFromFile(config.get<std::string>("gridConfigurationFilename"), 2))
                taps = std::max(taps, o);
        const double vectors = (dim + 1) + (dim + 1) + 2 + 1; // p, v;  pWaveModulus, inverse density averages;  vp, rho;  scratch
        return (ValueType)((vectors * 4.0 + 2.0 * dim * taps * 8.0) * N / 1024.0 / 1024.0);
    }
    d.nranks = (int32_t)deviceIds.size(); // memory of the first slab (the largest one)
    return (ValueType)(ws_estimate_memory(&d) / 1024.0 / 1024.0);
}

template <typename ValueType>
void ForwardSolver::ForwardSolver<ValueType>::initForwardSolver(Configuration::Configuration const &config, Derivatives::Derivatives<ValueType> &derivatives,
                                                                Wavefields::Wavefields<ValueType> &wavefield, Modelparameter::Modelparameter<ValueType> &model,
                                                                Acquisition::Coordinates<ValueType> const &modelCoordinates, ValueType DT)
{
    SCAI_ASSERT_ERROR(derivatives.getSpatialFDorder() == config.get<IndexType>("spatialFDorder"), "Derivatives::init must be called with the same configuration")
    SCAI_ASSERT_ERROR(model.getEquationType() == equationType && wavefield.getEquationType() == equationType, "model / wavefield type differs from the solver type")
    group.reset(new DeviceGroup(deviceIds));
    ws_desc d = makeDesc(config, dimension, equationType, deviceIds[0]);
    d.dt = DT;
    if (modelCoordinates.isVariable()) {
        initIrregular(config, d, derivatives, wavefield, model, modelCoordinates);
        return;
    }
    group->create(d);
    NT = d.nt;
    SCAI_ASSERT_ERROR(group->getNGlobal() == (size_t)modelCoordinates.getNGridpoints(), "grid of the configuration and of the coordinates differ")
    // every rank receives the GLOBAL vectors and keeps its slab plus the ghost planes the averaging needs
    for (auto const &kv : model.getRawParameters())
        group->forEach([&](IndexType r) { check(ws_set_material(group->handle(r), kv.first.c_str(), kv.second.data(), kv.second.size())); });
    wavefield.init(d.n_relax);
    wavefield.bind(group.get());
    model.bind(group.get());
    srcVersion = recVersion = ~0ul;
}

template <typename ValueType> void ForwardSolver::ForwardSolver<ValueType>::updateModel(Modelparameter::Modelparameter<ValueType> &model)
{
    SCAI_ASSERT_ERROR(group, "initForwardSolver must be called before updateModel")
    SCAI_ASSERT_ERROR(model.getEquationType() == equationType, "model type differs from the solver type")
    for (auto const &kv : model.getRawParameters()) {
        SCAI_ASSERT_ERROR(kv.second.size() == group->getNGlobal(), "the model does not fit the grid of the solver")
        group->forEach([&](IndexType r) { check(ws_set_material(group->handle(r), kv.first.c_str(), kv.second.data(), kv.second.size())); });
    }
    model.bind(group.get());
}

// Variable grid / variable FD order: the operators are assembled here by the reference's rules (IrregularGrid.cpp) and handed to
// the library (operator-given mode); replaces Derivatives::init + prepareBoundaryConditions + Modelparameter::prepareForModelling
// of FDTD2D.cpp:186-232, CPML2DAcoustic.cpp:99-200, Acoustic.cpp prepareForModelling for this case.
template <typename ValueType>
void ForwardSolver::ForwardSolver<ValueType>::initIrregular(Configuration::Configuration const &config, ws_desc d, Derivatives::Derivatives<ValueType> &derivatives,
                                                            Wavefields::Wavefields<ValueType> &wavefield, Modelparameter::Modelparameter<ValueType> &model,
                                                            Acquisition::Coordinates<ValueType> const &mc)
{
    SCAI_ASSERT_ERROR(equationType == "acoustic", "variable grids / variable FD orders are available for the acoustic solvers (equationType=" << equationType << ")")
    const size_t N = (size_t)mc.getNGridpoints();
    group->createSparse(d, N);
    NT = d.nt;
    ws_solver *h = group->handle(0);
    const bool d3 = dimension == "3d", fs = d.free_surface == 1;
    IrregularOperators<ValueType> ops(mc, derivatives, d.dt);
    auto setOp = [&](const char *name, EllRows const &m) { check(ws_set_operator(h, name, (int32_t)m.taps, m.cols.data(), m.vals.data())); };
    setOp("Dxf", ops.derivative(0, true));
    setOp("Dxb", ops.derivative(0, false));
    setOp("Dyf", ops.derivative(1, true, fs)); // DyfFreeSurface takes the place of Dyf (ForwardSolver2Dacoustic.cpp:141-146)
    setOp("Dyb", ops.derivative(1, false));
    if (d3) {
        setOp("Dzf", ops.derivative(2, true));
        setOp("Dzb", ops.derivative(2, false));
    }
    if (mc.hasVariableSpacing()) {
        const char *names[3] = {"InterpolationFull", "InterpolationStaggeredX", "InterpolationStaggeredZ"};
        for (IndexType mode = 0; mode < (d3 ? 3 : 2); mode++) {
            const EllRows m = ops.interpolation(mode);
            check(ws_set_interpolation(h, names[mode], (int64_t)m.rows.size(), m.rows.data(), (int32_t)m.taps, m.cols.data(), m.vals.data()));
        }
    }
    if (d.damping == 2)
        for (IndexType axis = 0; axis < (d3 ? 3 : 2); axis++) {
            const IndexType ax = (!d3 && axis == 1) ? 1 : axis; // 2-D: axes x, y
            const CpmlProfile p = ops.cpml(ax, d.boundary_width, d.npower, d.fc_cpml, d.vmax_cpml, d.free_surface != 0);
            check(ws_set_cpml_profile(h, (int32_t)ax, (int64_t)p.idx.size(), p.idx.data(), p.a.data(), p.b.data(), p.aHalf.data(), p.bHalf.data()));
        }
    if (d.damping == 1) {
        const AbsProfile p = ops.abs(d.boundary_width, d.damping_coeff, d.free_surface, d3);
        check(ws_set_abs_profile(h, (int64_t)p.idx.size(), p.idx.data(), p.damping.data()));
    }
    if (fs) {
        const std::vector<int32_t> surf = ops.surfacePoints();
        check(ws_set_surface(h, (int64_t)surf.size(), surf.data()));
    }
    // prepareForModelling products on the host: P-wave modulus rho vp^2 (ModelparameterSeismic.cpp:131-136), staggered inverse densities
    auto const &vp = model.getVelocityP();
    auto const &rho = model.getDensity();
    SCAI_ASSERT_ERROR(vp.size() == N && rho.size() == N, "the model must hold one value per point of the variable grid")
    std::vector<ValueType> pw(N);
    for (size_t i = 0; i < N; i++)
        pw[i] = (rho[i] * vp[i]) * vp[i];
    check(ws_set_material(h, "pWaveModulus", pw.data(), N));
    const char *avgNames[3] = {"inverseDensityAverageX", "inverseDensityAverageY", "inverseDensityAverageZ"};
    for (IndexType axis = 0; axis < (d3 ? 3 : 2); axis++) {
        const std::vector<ValueType> r = ops.inverseAverage(rho, axis);
        check(ws_set_material(h, avgNames[axis], r.data(), N));
    }
    for (auto const &kv : model.getRawParameters())
        check(ws_set_material(h, kv.first.c_str(), kv.second.data(), kv.second.size()));
    wavefield.init(d.n_relax);
    wavefield.bind(group.get());
    model.bind(group.get());
    srcVersion = recVersion = ~0ul;
}

template <typename ValueType> void ForwardSolver::ForwardSolver<ValueType>::prepareForModelling(Modelparameter::Modelparameter<ValueType> const &, ValueType)
{
    SCAI_ASSERT_ERROR(group, "initForwardSolver must be called before prepareForModelling")
    group->forEach([&](IndexType r) { check(ws_prepare(group->handle(r))); });
}

template <typename ValueType> void ForwardSolver::ForwardSolver<ValueType>::resetCPML()
{
    SCAI_ASSERT_ERROR(group, "initForwardSolver must be called first")
    group->forEach([&](IndexType r) { check(ws_reset(group->handle(r))); }); // memory variables (and wavefields, traces: both are re-initialised per shot anyway)
}

template <typename ValueType>
void ForwardSolver::ForwardSolver<ValueType>::bindAcquisition(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources)
{
    if (srcObj != &sources || srcVersion != sources.getVersion()) {
        auto const &types = sources.getSeismogramTypes();
        auto const &idx = sources.get1DCoordinates();
        std::vector<ValueType> signals((size_t)types.size() * NT);
        for (size_t k = 0; k < types.size(); k++) {
            auto const &sg = sources.getSeismogramHandler().getSeismogram(types[k] - 1);
            SCAI_ASSERT_ERROR(sg.getNumSamples() == NT, "source signals must hold NT samples")
            std::memcpy(&signals[k * NT], &sg.getData()[(size_t)sources.getRowOfEntry((IndexType)k) * NT], sizeof(ValueType) * NT);
        }
        group->forEach([&](IndexType r) { check(ws_set_sources(group->handle(r), (int32_t)types.size(), types.data(), idx.data(), signals.data())); });
        srcObj = &sources;
        srcVersion = sources.getVersion();
    }
    if (recObj != &receiver || recVersion != receiver.getVersion()) {
        group->forEach([&](IndexType r) {
            check(ws_set_receivers(group->handle(r), (int32_t)receiver.getSeismogramTypes().size(), receiver.getSeismogramTypes().data(), receiver.get1DCoordinates().data()));
        });
        recObj = &receiver;
        recVersion = receiver.getVersion();
    }
}

template <typename ValueType> void ForwardSolver::ForwardSolver<ValueType>::fetchSeismograms(Acquisition::Receivers<ValueType> &receiver)
{
    auto const &types = receiver.getSeismogramTypes();
    std::vector<ValueType> all((size_t)types.size() * NT);
    group->getSeismogram(all);
    for (size_t k = 0; k < types.size(); k++) {
        auto &sg = receiver.getSeismogramHandler().getSeismogram(types[k] - 1);
        std::memcpy(&sg.getData()[(size_t)receiver.getRowOfEntry((IndexType)k) * NT], &all[k * NT], sizeof(ValueType) * NT);
    }
}

template <typename ValueType>
void ForwardSolver::ForwardSolver<ValueType>::run(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources,
                                                  Modelparameter::Modelparameter<ValueType> const &, Wavefields::Wavefields<ValueType> &,
                                                  Derivatives::Derivatives<ValueType> const &, IndexType t)
{
    SCAI_ASSERT_ERROR(group, "initForwardSolver must be called before run")
    bindAcquisition(receiver, sources);
    group->forEach([&](IndexType r) { check(ws_step(group->handle(r), t)); });
    if (t == NT - 1)
        fetchSeismograms(receiver);
}

template <typename ValueType>
void ForwardSolver::ForwardSolver<ValueType>::run(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources, IndexType t0, IndexType t1)
{
    SCAI_ASSERT_ERROR(group, "initForwardSolver must be called before run")
    bindAcquisition(receiver, sources);
    group->forEach([&](IndexType r) { check(ws_run(group->handle(r), t0, t1)); });
    if (t1 == NT)
        fetchSeismograms(receiver);
}

template <typename ValueType> void ForwardSolver::ForwardSolver<ValueType>::sync()
{
    if (group)
        group->forEach([&](IndexType r) { check(ws_sync(group->handle(r))); });
}

template <typename ValueType> typename ForwardSolver::ForwardSolver<ValueType>::ForwardSolverPtr ForwardSolver::Factory<ValueType>::Create(std::string dimension, std::string type)
{
    std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
    std::transform(type.begin(), type.end(), type.begin(), ::tolower);
    return std::make_shared<ForwardSolver<ValueType>>(dimension, type);
}

template class ForwardSolver::ForwardSolver<float>;
template class ForwardSolver::Factory<float>;

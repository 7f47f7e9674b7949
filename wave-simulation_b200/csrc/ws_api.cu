// ws_api.cu — C ABI (include/wavesim.h) of the B200-native FD time-stepping library: solver object, HBM layout,
// model preparation, acquisition, time-step orchestration (streams, CUDA graphs, y-slab halo exchange over NCCL).
//
// Reference interfaces replaced (relative to src/): ForwardSolver/ForwardSolver.hpp:38-52 (run, prepareForModelling,
// resetCPML, initForwardSolver), Wavefields/Wavefields.hpp (resetWavefields, getRef*), Modelparameter/*.cpp
// (prepareForModelling), ForwardSolver/SourceReceiverImpl/*.cpp, Partitioning/Partitioning.hpp (replaced by y-slabs).
#include "../../include/wavesim.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <unistd.h>
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <vector>

struct WsError : public std::exception {
    int code;
    std::string msg;
    WsError(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char *what() const noexcept override { return msg.c_str(); }
};

#include "ws_common.cuh"
#include "ws_acquisition.cuh"
#include "ws_launch.hpp"
#include "ws_prepare.cuh"
#include "ws_tables.hpp"
#include "ws_kernels_tma.cuh" // operand table (wsmarch::spec) and the TMA program type
#include "ws_kernels_tile2d.cuh" // program type of the 2-D tile kernels
#include "ws_kernels_sparse.cuh"

namespace {

thread_local std::string g_lastError;

#define WS_REQUIRE(cond, code, msg)                                                                                    \
    do {                                                                                                               \
        if (!(cond))                                                                                                   \
            throw WsError(code, msg);                                                                                  \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// NCCL through dlopen: if the host process already carries a libnccl.so.2 (e.g. torch's), that one is reused.
// ---------------------------------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void load()
    {
        if (lib)
            return;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib)
                break;
        }
        WS_REQUIRE(lib, WS_ECOMM, std::string("cannot load libnccl: ") + dlerror());
#define NCCL_SYM(field, name)                                                                                          \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                                                       \
    WS_REQUIRE(field, WS_ECOMM, std::string("missing NCCL symbol ") + name);
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(Send, "ncclSend")
        NCCL_SYM(Recv, "ncclRecv")
        NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(AllReduce, "ncclAllReduce")
        NCCL_SYM(AllGather, "ncclAllGather")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    }
    void check(int rc, const char *what)
    {
        if (rc != 0)
            throw WsError(WS_ECOMM, std::string(what) + ": " + (GetErrorString ? GetErrorString(rc) : "nccl error"));
    }
};
NcclApi g_nccl;
constexpr int kNcclFloat = 7; // ncclFloat32
constexpr int kNcclInt = 2;   // ncclInt32
constexpr int kNcclChar = 0;  // ncclInt8
constexpr int kNcclMax = 2;   // ncclMax

// ---------------------------------------------------------------------------------------------------------------------
// acquisition kernels (SourceReceiverImpl.cpp:12-37, FDTD3Delastic.cpp:12-53, FDTD2Delastic.cpp, FDTDacoustic.cpp,
// ForwardSolverEM/SourceReceiverImpl/SourceReceiverImplEM.cpp)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void kSources(const __grid_constant__ WsParams P, WsAcq a, int sequential)
{
    const int t = *a.tdev;
    if (sequential) {
        if (blockIdx.x != 0 || threadIdx.x != 0)
            return;
        wsSourcesSequential(P, a, t);
    } else
        wsSourceOne(P, a, t, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void kReceivers(const __grid_constant__ WsParams P, WsAcq a) { wsReceiverOne(P, a, *a.tdev, blockIdx.x * blockDim.x + threadIdx.x); }
// the time index lives in device memory so that a captured CUDA graph is step-invariant
__global__ void kAdvance(int *tdev) { *tdev = *tdev + 1; }
// Sources, receivers and the time index of a step in ONE launch of one thread block (small acquisition geometries: three
// launches of 2-5 us each weigh 3 % of a 2-D step of 0.25 ms).  Same order as the three kernels: all sources (block barrier
// + fence), then the receivers, then the time index.
#ifndef WS_EMULATE
constexpr int WS_ACQ_THREADS = 512, WS_ACQ_MAX_SRC = 2048, WS_ACQ_MAX_REC = 8192;
__global__ void __launch_bounds__(WS_ACQ_THREADS) kAcquisition(const __grid_constant__ WsParams P, WsAcq a, int sequential)
{
    // programmatic dependent launch (see ws_kernels_tile2d.cuh): scheduled while the second half-step drains, waits for it here
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    wsAcquisitionBlock(P, a, sequential);
}
#endif

// element-wise operators of the wavefield objects (Wavefields.hpp:62-80): op 0 dst = src, 1 dst += src, 2 dst -= src
__global__ void kWfBinary(float *__restrict__ dst, const float *__restrict__ src, size_t n, int op)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float b = src[i];
    dst[i] = op == 0 ? b : (op == 1 ? __fadd_rn(dst[i], b) : __fsub_rn(dst[i], b));
}
__global__ void kWfScale(float *__restrict__ dst, size_t n, float a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = __fmul_rn(dst[i], a);
}
__global__ void kWfScaleVec(float *__restrict__ dst, const float *__restrict__ vec, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = __fmul_rn(dst[i], vec[i]);
}
// every component of the live wavefields times a grid vector, own planes only (the ghost planes belong to the neighbours)
struct WsFieldTable {
    float *p[F_COUNT];
    int n;
};
__global__ void kWfScaleVecAll(WsFieldTable t, const float *__restrict__ vec, size_t first, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float v = vec[first + i];
    for (int k = 0; k < t.n; k++)
        t.p[k][first + i] = __fmul_rn(t.p[k][first + i], v);
}

// ---------------------------------------------------------------------------------------------------------------------
// halo exchange over peer memory (NVLink / NVSwitch): the sender's kernel stores its edge planes straight into the ghost planes of
// the neighbour's arrays and then raises a counter in the neighbour's memory; the neighbour's edge-slab kernels are preceded by a
// one-thread kernel that waits for that counter.  No send / recv pairing, no host involvement, nothing but kernels: the exchange is
// captured in the step graph like any other launch.  Counters (one int each, device memory of the RECEIVER, written by its
// neighbours; P2P_* below) count exchanges since the solver was created; who expects which count lives in device memory too, so
// the captured graph is step-invariant.  Write-after-read on the ghost planes needs no acknowledgement: a rank can only push
// exchange A of step n after it has received exchange B of step n-1, which its neighbour sends after the kernels that read the
// previous ghost planes (and likewise for B after A).
// ---------------------------------------------------------------------------------------------------------------------
enum { P2P_FLAG_UP_A = 0, P2P_FLAG_UP_B = 1, P2P_FLAG_DOWN_A = 2, P2P_FLAG_DOWN_B = 3, P2P_WAIT_A = 4, P2P_WAIT_B = 5, P2P_PUSH_A = 6, P2P_PUSH_B = 7,
       P2P_DONE = 8, P2P_ERROR = 9, P2P_NINTS = 16 };
#ifndef WS_EMULATE
constexpr int WS_P2P_MAXF = 6;
struct WsPushArgs {
    const float *src[2][WS_P2P_MAXF]; // [0: to the upper neighbour (rank - 1), 1: to the lower neighbour][field]
    float *dst[2][WS_P2P_MAXF];       // ghost planes inside the neighbour's arrays (peer-mapped)
    int *peerFlag[2];                 // the counter this exchange raises at the neighbour (null: no neighbour on that side)
    int *local;                       // this rank's P2P_* block
    int nf, type;                     // fields, 0 = exchange A (after the first half-step), 1 = exchange B
    size_t count4;                    // float4 per field and direction
};
__device__ __forceinline__ int ldAcquireSys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseSys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__global__ void __launch_bounds__(256) kHaloPush(const WsPushArgs a)
{
    const int dir = blockIdx.y / a.nf, f = blockIdx.y - dir * a.nf;
    if (a.peerFlag[dir]) {
        const float4 *__restrict__ src = reinterpret_cast<const float4 *>(a.src[dir][f]);
        float4 *__restrict__ dst = reinterpret_cast<float4 *>(a.dst[dir][f]);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count4; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
    __threadfence_system(); // the planes are visible at the neighbour before the counter is
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total = gridDim.x * gridDim.y;
        if (atomicAdd(a.local + P2P_DONE, 1) == total - 1) { // the last thread block of the grid raises the counters
            __threadfence_system();
            const int n = a.local[P2P_PUSH_A + a.type] + 1;
            // at the upper neighbour this rank is the "down" side, at the lower neighbour the "up" side
            if (a.peerFlag[0])
                stReleaseSys(a.peerFlag[0] + P2P_FLAG_DOWN_A + a.type, n);
            if (a.peerFlag[1])
                stReleaseSys(a.peerFlag[1] + P2P_FLAG_UP_A + a.type, n);
            a.local[P2P_PUSH_A + a.type] = n;
            a.local[P2P_DONE] = 0;
        }
    }
}
// waits until the ghost planes of exchange `type` have landed from both neighbours: A is expected once more than the waits done so
// far (the exchange of this step), B as often as the waits done so far (the exchange of the step before; none before the first step)
__global__ void kHaloWait(int *local, int type, int hasUp, int hasDown)
{
    const int e = type == 0 ? local[P2P_WAIT_A] + 1 : local[P2P_WAIT_B];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        const bool up = !hasUp || ldAcquireSys(local + P2P_FLAG_UP_A + type) >= e;
        const bool down = !hasDown || ldAcquireSys(local + P2P_FLAG_DOWN_A + type) >= e;
        if (up && down)
            break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { // 20 s: a neighbour is gone; report instead of hanging the GPU
            local[P2P_ERROR] = 1;
            break;
        }
        __nanosleep(200);
    }
    local[P2P_WAIT_A + type] = type == 0 ? e : e + 1;
}
#endif

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool owned = true;
    // slot of an arena owned by another DevBuf
    void borrow(T *ptr, size_t count)
    {
        release();
        p = ptr;
        n = count;
        owned = false;
    }
    void alloc(size_t count)
    {
        release();
        owned = true;
        n = count;
        if (count) {
            cudaError_t e = cudaMalloc(&p, count * sizeof(T));
            if (e != cudaSuccess)
                throw WsError(WS_ENOMEM, std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e));
        }
    }
    void upload(const std::vector<T> &h)
    {
        alloc(h.size());
        if (!h.empty()) {
            WS_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
            // a pageable host-to-device copy may return before its DMA has landed; the solver's kernels run on a
            // non-blocking stream that is not ordered with the legacy default stream
            WS_CUDA_CHECK(cudaStreamSynchronize(0));
        }
    }
    // always on the solver's own stream: it is a non-blocking stream, so work on the legacy default stream is not ordered with it
    void zero(cudaStream_t st)
    {
        if (p)
            WS_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    }
    void release()
    {
        if (p && owned)
            cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct NameSlot {
    const char *name;
    int slot;
};
const NameSlot kMatNames[] = {
    {"velocityP", M_VP}, {"velocityS", M_VS}, {"density", M_RHO}, {"tauP", M_TAUP}, {"tauS", M_TAUS},
    {"pWaveModulus", M_PW}, {"sWaveModulus", M_MU},
    {"inverseDensityAverageX", M_RIX}, {"inverseDensityAverageY", M_RIY}, {"inverseDensityAverageZ", M_RIZ},
    {"sWaveModulusAverageXY", M_MUXY}, {"sWaveModulusAverageXZ", M_MUXZ}, {"sWaveModulusAverageYZ", M_MUYZ},
    {"tauSAverageXY", M_TSXY}, {"tauSAverageXZ", M_TSXZ}, {"tauSAverageYZ", M_TSYZ}, {"inverseDensity", M_INVRHO},
    {"dielectricPermittivity", M_EPS}, {"electricConductivity", M_SIG}, {"magneticPermeability", M_MUM},
    {"tauDielectricPermittivity", M_TAUEPS}, {"tauElectricConductivity", M_TAUSIG},
    {"inverseMagneticPermeabilityAverageXY", M_MIXY}, {"inverseMagneticPermeabilityAverageXZ", M_MIXZ},
    {"inverseMagneticPermeabilityAverageYZ", M_MIYZ},
    {"CaAverageX", M_CAX}, {"CaAverageY", M_CAY}, {"CaAverageZ", M_CAZ},
    {"CbAverageX", M_CBX}, {"CbAverageY", M_CBY}, {"CbAverageZ", M_CBZ},
    {"CdAverageX1", M_CD0 + 0}, {"CdAverageY1", M_CD0 + 1}, {"CdAverageZ1", M_CD0 + 2},
    {"CdAverageX2", M_CD0 + 3}, {"CdAverageY2", M_CD0 + 4}, {"CdAverageZ2", M_CD0 + 5},
    {"CdAverageX3", M_CD0 + 6}, {"CdAverageY3", M_CD0 + 7}, {"CdAverageZ3", M_CD0 + 8},
    {"CdAverageX4", M_CD0 + 9}, {"CdAverageY4", M_CD0 + 10}, {"CdAverageZ4", M_CD0 + 11},
};

} // namespace

struct ws_solver {
    ws_desc d{};
    int nx = 0, gny = 0, nz = 0, y0 = 0, nyl = 0;
    int pitch = 0, nzp = 0;
    long long plane = 0, base = 0, total = 0;
    int q = 0, h = 0, L = 0, W = 0;
    bool seismic = true, visco = false, exact = false;
    cudaStream_t stream = nullptr, commStream = nullptr;
    cudaEvent_t evCompute = nullptr, evComm = nullptr;
    cudaEvent_t evComputeG = nullptr, evCommG = nullptr; // the same roles inside a stream capture (a captured event cannot be waited on outside)
    bool capturing = false;
    DevBuf<float> fldArena, matArena, psiXArena, psiZArena; // declared first: the slots below borrow from them
    DevBuf<float> arena; // every other solver: wavefields, then the model parameters the kernels read (TMA marching kernels)
    WsArenaInfo ainfo;
    wstma::TmaProg tmaProg[2];
    int tmaNL[2] = {4, 4};
    void *tmaMaps = nullptr;
    bool useTile = false; // 2-D tile kernels (ws_kernels_tile2d.cu) serve this configuration
    wstile::TileProg tileProg[2];
    void *tileMaps = nullptr;
    bool useTma = false; // TMA marching kernels (ws_kernels_tma.cu) serve this configuration
    DevBuf<float> fld[F_COUNT], mat[M_COUNT], psi[PSI_COUNT];
    bool matGiven[M_COUNT] = {};
    int psiAxis[PSI_COUNT];
    std::map<std::string, int> fldSlot;
    DevBuf<float> cxTab;
    DevBuf<float> tab, cax, cbx, caxh, cbxh, cay, cby, cayh, cbyh, caz, cbz, cazh, cbzh, absCoeff;
    DevBuf<float> sH, sV, sRH[4], sRV[4];
    DevBuf<float> scratch; // dense staging buffer for pack/unpack
    DevBuf<float> stepScale; // `*wavefields *= compensation` after every step (ws_set_step_scaling), padded like a wavefield; null = off
    DevBuf<int> flag;
    WsParams P{};
    bool prepared = false;
    // acquisition
    int nsrc = 0, nrec = 0;
    bool srcSequential = true;
    DevBuf<int> srcType, recType, tdev;
    DevBuf<long long> srcOff, recOff;
    DevBuf<float> srcSig, srcStep, seis, recStep;
    std::vector<int> recOwned; // 1 if the receiver lives on this rank
    float *pinSrc = nullptr, *pinRec = nullptr;
    WsAcq acq{};
    // halo exchange lists
    std::vector<int> exchA, exchB; // field slots whose y-ghost planes are needed after pass A / pass B
    void *ncclComm = nullptr;
    // halo exchange over peer memory (ws_comm_init sets it up when every neighbour's memory can be mapped)
    bool p2p = false;
    DevBuf<int> p2pLocal;                      // P2P_* block of this rank
    float *peerArena[2] = {nullptr, nullptr};  // wavefield arena of the upper / lower neighbour, mapped into this process
    int *peerFlags[2] = {nullptr, nullptr};    // their P2P_* blocks
    bool peerIpc[2] = {false, false};          // opened with cudaIpcOpenMemHandle (another process)
    int peerNyl[2] = {0, 0};
    ws_sendrecv_fn extFn = nullptr; // bring-your-own transport (ws_comm_init_external)
    void *extUser = nullptr;
    // instrumentation
    uint64_t launches = 0;
    bool timing = false;
    std::vector<cudaEvent_t> evPool;
    float msA = 0, msB = 0, msStep = 0;
    bool useFast = false;
    bool useFastA = false; // 3-D viscoelastic: the velocity half-step runs the tiled elastic kernel (same statements)
    bool useMarch = false; // marching kernels (ws_kernels_march.cu) serve this configuration
    void *fastMaps = nullptr;
    // operator-given mode (ws_create_sparse): irregular grids, the operators come from the caller in ELL form
    bool sparse = false;
    struct Interp {
        long long nrows = 0;
        int taps = 0;
        DevBuf<int> rows, cols;
        DevBuf<float> vals;
    };
    DevBuf<int> spCol[wssparse::SP_NOPS], spCpK[3], spSurf;
    DevBuf<float> spVal[wssparse::SP_NOPS], spCa[3], spCb[3], spCah[3], spCbh[3], spPsi[wssparse::SPSI_COUNT], spTmp;
    int spTaps[wssparse::SP_NOPS] = {};
    long long spSurfN = 0;
    DevBuf<int> spAbsIdx; // ABS frame in operator-given mode (ws_set_abs_profile)
    DevBuf<float> spAbsVal;
    long long spAbsN = 0;
    Interp spInterp[3]; // full grid, staggered in x, staggered in z
    // CUDA graph of one time step
    cudaGraphExec_t graphExec = nullptr;
    int graphSteps = 0;
    uint64_t graphLaunchesPerStep = 0;

    ~ws_solver()
    {
        if (graphExec)
            cudaGraphExecDestroy(graphExec);
        if (fastMaps)
            wsFastRelease(fastMaps);
#ifndef WS_EMULATE
        if (tmaMaps)
            wsTmaRelease(tmaMaps);
        if (tileMaps)
            wsTileRelease(tileMaps);
#endif
        for (auto e : evPool)
            cudaEventDestroy(e);
        if (evCompute)
            cudaEventDestroy(evCompute);
        if (evComm)
            cudaEventDestroy(evComm);
        if (evComputeG)
            cudaEventDestroy(evComputeG);
        if (evCommG)
            cudaEventDestroy(evCommG);
#ifndef WS_EMULATE
        for (int d = 0; d < 2; d++)
            if (peerIpc[d]) {
                cudaIpcCloseMemHandle(peerArena[d]);
                cudaIpcCloseMemHandle(peerFlags[d]);
            }
#endif
        if (ncclComm && g_nccl.CommDestroy)
            g_nccl.CommDestroy(ncclComm);
        if (pinSrc)
            cudaFreeHost(pinSrc);
        if (pinRec)
            cudaFreeHost(pinRec);
        if (commStream)
            cudaStreamDestroy(commStream);
        if (stream)
            cudaStreamDestroy(stream);
    }

    wsprep::Geo geo(int ylo, int yhi) const
    {
        wsprep::Geo g;
        g.nx = nx; g.nyl = nyl; g.nz = nz; g.gny = gny; g.gy0 = y0; g.pitch = pitch; g.nzp = nzp;
        g.plane = plane; g.base = base; g.ylo = ylo; g.yhi = yhi;
        return g;
    }
    void gridFor(int ylo, int yhi, dim3 &grid, dim3 &block) const
    {
        wsPointGrid(nx, nz, std::max(0, yhi - ylo), grid, block);
    }
    long long offsetOf(int x, int ly, int z) const { return base + x + (long long)z * pitch + (long long)ly * plane; }
};

namespace {

void setDevice(const ws_solver *s) { WS_CUDA_CHECK(cudaSetDevice(s->d.device)); }

bool eqIsVisco(int eq) { return eq == WS_EQ_VISCOELASTIC || eq == WS_EQ_VISCOSH || eq == WS_EQ_VISCOTMEM || eq == WS_EQ_VISCOEMEM; }

void validateDesc(const ws_desc &d)
{
    WS_REQUIRE(d.dim == 2 || d.dim == 3, WS_EINVAL, "dimension must be 2D or 3D");
    WS_REQUIRE(d.eq >= WS_EQ_ACOUSTIC && d.eq <= WS_EQ_VISCOEMEM, WS_EINVAL, "unknown equationType");
    if (d.dim == 3)
        WS_REQUIRE(d.eq != WS_EQ_SH && d.eq != WS_EQ_VISCOSH && d.eq != WS_EQ_TMEM && d.eq != WS_EQ_VISCOTMEM, WS_EINVAL,
                   "sh, viscosh, tmem and viscotmem exist in 2D only (ForwardSolverFactory.cpp:4-66)");
    WS_REQUIRE(d.nx > 0 && d.ny > 0 && (d.dim == 2 || d.nz > 0), WS_EINVAL, "NX, NY, NZ must be positive");
    WS_REQUIRE(d.dh > 0 && d.dt > 0 && d.nt > 0, WS_EINVAL, "DH, DT, NT must be positive");
    WS_REQUIRE(d.fd_order >= 2 && d.fd_order <= WS_MAXQ && d.fd_order % 2 == 0, WS_EINVAL,
               "spatialFDorder = " + std::to_string(d.fd_order) + " Unsupported spatialFDorder value.");
    WS_REQUIRE(d.edge_policy == 0 || d.edge_policy == 1, WS_EINVAL, "edge_policy must be 0 or 1");
    WS_REQUIRE(d.free_surface >= 0 && d.free_surface <= 2, WS_EINVAL, "FreeSurface must be 0, 1 or 2");
    WS_REQUIRE(d.damping >= 0 && d.damping <= 2, WS_EINVAL, "DampingBoundary must be 0, 1 or 2");
    if (d.damping) {
        WS_REQUIRE(d.boundary_width > 0, WS_EINVAL, "BoundaryWidth must be positive");
        WS_REQUIRE(2 * d.boundary_width <= d.nx && 2 * d.boundary_width <= d.ny && (d.dim == 2 || 2 * d.boundary_width <= d.nz),
                   WS_EINVAL, "2*BoundaryWidth must not exceed the grid extent");
    }
    if (eqIsVisco(d.eq))
        WS_REQUIRE(d.n_relax >= 1 && d.n_relax <= WS_MAX_RELAX, WS_EINVAL, "numRelaxationMechanisms more than 4 is not available here!");
    WS_REQUIRE(d.nranks >= 1 && d.rank >= 0 && d.rank < d.nranks, WS_EINVAL, "invalid rank / nranks");
    const int nz = d.dim == 2 ? 1 : d.nz;
    (void)nz; // grids beyond 2^31 points are accepted; they need the 64-bit acquisition entry points
    WS_REQUIRE(d.ny / d.nranks >= std::max(d.fd_order / 2, 1) * 2 || d.nranks == 1, WS_EINVAL, "y-slabs thinner than the stencil");
    // the raw model parameters exchange WS_HALO ghost planes (ws_set_material_device): a thinner slab would send its own ghosts
    WS_REQUIRE(d.ny / d.nranks >= WS_HALO || d.nranks == 1, WS_EINVAL, "y-slabs thinner than " + std::to_string(WS_HALO) + " planes are not supported");
}

void slabRange(const ws_desc &d, int rank, int &y0, int &nyl)
{
    // block distribution of planes, remainder to the first ranks (same rule as dmemo::BlockDistribution)
    const int base = d.ny / d.nranks, rem = d.ny % d.nranks;
    nyl = base + (rank < rem ? 1 : 0);
    y0 = rank * base + std::min(rank, rem);
}

struct FieldList {
    std::vector<std::pair<std::string, int>> f;
    void add(const std::string &n, int slot) { f.emplace_back(n, slot); }
};

FieldList fieldsFor(const ws_desc &d)
{
    FieldList fl;
    const int L = eqIsVisco(d.eq) ? d.n_relax : 0;
    const bool d3 = d.dim == 3;
    auto rname = [](const char *b, int l) { return std::string(b) + std::to_string(l + 1); };
    switch (d.eq) {
    case WS_EQ_ACOUSTIC:
        fl.add("VX", F_VX); fl.add("VY", F_VY); if (d3) fl.add("VZ", F_VZ); fl.add("P", F_P);
        break;
    case WS_EQ_ELASTIC:
    case WS_EQ_VISCOELASTIC:
        fl.add("VX", F_VX); fl.add("VY", F_VY); fl.add("Sxx", F_SXX); fl.add("Syy", F_SYY); fl.add("Sxy", F_SXY);
        if (d3) { fl.add("VZ", F_VZ); fl.add("Szz", F_SZZ); fl.add("Sxz", F_SXZ); fl.add("Syz", F_SYZ); }
        for (int l = 0; l < L; l++) {
            fl.add(rname("Rxx", l), F_R0 + 6 * l + RC_XX); fl.add(rname("Ryy", l), F_R0 + 6 * l + RC_YY); fl.add(rname("Rxy", l), F_R0 + 6 * l + RC_XY);
            if (d3) { fl.add(rname("Rzz", l), F_R0 + 6 * l + RC_ZZ); fl.add(rname("Rxz", l), F_R0 + 6 * l + RC_XZ); fl.add(rname("Ryz", l), F_R0 + 6 * l + RC_YZ); }
        }
        break;
    case WS_EQ_SH:
    case WS_EQ_VISCOSH:
        fl.add("VZ", F_VZ); fl.add("Sxz", F_SXZ); fl.add("Syz", F_SYZ);
        for (int l = 0; l < L; l++) { fl.add(rname("Rxz", l), F_R0 + 6 * l + RC_XZ); fl.add(rname("Ryz", l), F_R0 + 6 * l + RC_YZ); }
        break;
    case WS_EQ_TMEM:
    case WS_EQ_VISCOTMEM:
        fl.add("HX", F_HX); fl.add("HY", F_HY); fl.add("EZ", F_EZ);
        for (int l = 0; l < L; l++) fl.add(rname("RZ", l), F_R0 + 6 * l + RC_Z);
        break;
    case WS_EQ_EMEM:
    case WS_EQ_VISCOEMEM:
        fl.add("HZ", F_HZ); fl.add("EX", F_EX); fl.add("EY", F_EY);
        if (d3) { fl.add("HX", F_HX); fl.add("HY", F_HY); fl.add("EZ", F_EZ); }
        for (int l = 0; l < L; l++) {
            fl.add(rname("RX", l), F_R0 + 6 * l + RC_X); fl.add(rname("RY", l), F_R0 + 6 * l + RC_Y);
            if (d3) fl.add(rname("RZ", l), F_R0 + 6 * l + RC_Z);
        }
        break;
    }
    return fl;
}

// CPML memory variables used by an equation type and the axis each one lives on
std::vector<std::pair<int, int>> psiFor(const ws_desc &d)
{
    std::vector<std::pair<int, int>> v;
    const bool d3 = d.dim == 3;
    auto add = [&](int slot, int axis) { if (axis != 2 || d3) v.emplace_back(slot, axis); };
    switch (d.eq) {
    case WS_EQ_ACOUSTIC:
        add(PSI_P_X, 0); add(PSI_P_Y, 1); add(PSI_P_Z, 2); add(PSI_VXX, 0); add(PSI_VYY, 1); add(PSI_VZZ, 2);
        break;
    case WS_EQ_ELASTIC:
    case WS_EQ_VISCOELASTIC:
        add(PSI_SXX_X, 0); add(PSI_SXY_X, 0); add(PSI_SXY_Y, 1); add(PSI_SYY_Y, 1);
        add(PSI_VXX, 0); add(PSI_VYX, 0); add(PSI_VXY, 1); add(PSI_VYY, 1);
        if (d3) {
            add(PSI_SXZ_X, 0); add(PSI_SYZ_Y, 1); add(PSI_SXZ_Z, 2); add(PSI_SYZ_Z, 2); add(PSI_SZZ_Z, 2);
            add(PSI_VZX, 0); add(PSI_VZY, 1); add(PSI_VXZ, 2); add(PSI_VYZ, 2); add(PSI_VZZ, 2);
        }
        break;
    case WS_EQ_SH:
    case WS_EQ_VISCOSH:
        add(PSI_SXZ_X, 0); add(PSI_SYZ_Y, 1); add(PSI_VZX, 0); add(PSI_VZY, 1);
        break;
    case WS_EQ_TMEM:
    case WS_EQ_VISCOTMEM:
        add(PSI_EZX, 0); add(PSI_EZY, 1); add(PSI_HYX, 0); add(PSI_HXY, 1);
        break;
    case WS_EQ_EMEM:
    case WS_EQ_VISCOEMEM:
        add(PSI_EYX, 0); add(PSI_EXY, 1); add(PSI_HZX, 0); add(PSI_HZY, 1);
        if (d3) {
            add(PSI_EZX, 0); add(PSI_EZY, 1); add(PSI_EXZ, 2); add(PSI_EYZ, 2);
            add(PSI_HYX, 0); add(PSI_HXY, 1); add(PSI_HXZ, 2); add(PSI_HYZ, 2);
        }
        break;
    }
    return v;
}

// row layout of the x-term slabs (ws_common.cuh wsPsiXIndex): shift D of the high side and row length PX
void psiXLayout(int nx, int W, int &D, int &PX)
{
    const int W4 = (W + 3) / 4 * 4;
    D = nx - W - W4 >= 0 ? (nx - W - W4) / 4 * 4 : 0;
    PX = (nx - D + 3) / 4 * 4;
}

size_t psiSize(const ws_solver *s, int axis)
{
    const size_t W2 = 2 * (size_t)s->W;
    if (axis == 0) {
        int D, PX;
        psiXLayout(s->nx, s->W, D, PX);
        return (size_t)s->nyl * s->nz * PX;
    }
    if (axis == 1)
        return W2 * s->nz * s->nx;
    return (size_t)s->nyl * W2 * s->nx;
}

int matSlotOf(const char *name)
{
    for (const auto &ns : kMatNames)
        if (std::strcmp(ns.name, name) == 0)
            return ns.slot;
    throw WsError(WS_EINVAL, std::string("unknown model parameter '") + name + "'");
}

float *matBuf(ws_solver *s, int slot)
{
    if (!s->mat[slot].p) {
        s->mat[slot].alloc((size_t)s->total);
        s->mat[slot].zero(s->stream);
    }
    return s->mat[slot].p;
}

void ensureScratch(ws_solver *s, size_t n)
{
    if (s->scratch.n < n)
        s->scratch.alloc(n);
}

// dense (reference linear order) host/device slab -> padded array; planes [ylo,yhi) local, dense plane 0 = local ylo
void packPlanes(ws_solver *s, const float *denseDev, float *padded, int ylo, int yhi)
{
    dim3 grid, block;
    s->gridFor(ylo, yhi, grid, block);
    if (grid.z == 0)
        return;
    WS_LAUNCH(wsprep::kPack, grid, block, 0, s->stream, s->geo(ylo, yhi), denseDev, padded, ylo);
    WS_CUDA_CHECK(cudaGetLastError()); // set-up path: a failed launch must not pass for a copied array
    s->launches++;
}

void uploadGlobal(ws_solver *s, const float *hostGlobal, float *padded)
{
    // planes [y0-HALO, y0+nyl+HALO) ∩ [0, NY) of the global vector, staged in chunks
    const int glo = std::max(0, s->y0 - WS_HALO), ghi = std::min(s->gny, s->y0 + s->nyl + WS_HALO);
    const size_t planeDense = (size_t)s->nx * s->nz;
    const int chunk = std::max(1, (int)std::min<size_t>((size_t)(ghi - glo), (size_t)(256u << 20) / (planeDense * sizeof(float)) + 1));
    ensureScratch(s, planeDense * chunk);
    for (int g0 = glo; g0 < ghi; g0 += chunk) {
        const int g1 = std::min(ghi, g0 + chunk);
        WS_CUDA_CHECK(cudaMemcpyAsync(s->scratch.p, hostGlobal + (size_t)g0 * planeDense, (size_t)(g1 - g0) * planeDense * sizeof(float),
                                      cudaMemcpyHostToDevice, s->stream));
        packPlanes(s, s->scratch.p, padded, g0 - s->y0, g1 - s->y0);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    }
}

void downloadLocal(ws_solver *s, const float *padded, float *hostLocal)
{
    const size_t planeDense = (size_t)s->nx * s->nz;
    const int chunk = std::max(1, (int)std::min<size_t>((size_t)s->nyl, (size_t)(256u << 20) / (planeDense * sizeof(float)) + 1));
    ensureScratch(s, planeDense * chunk);
    for (int l0 = 0; l0 < s->nyl; l0 += chunk) {
        const int l1 = std::min(s->nyl, l0 + chunk);
        dim3 grid, block;
        s->gridFor(l0, l1, grid, block);
        WS_LAUNCH(wsprep::kUnpack, grid, block, 0, s->stream, s->geo(l0, l1), padded, s->scratch.p);
        WS_CUDA_CHECK(cudaGetLastError());
        s->launches++;
        WS_CUDA_CHECK(cudaMemcpyAsync(hostLocal + (size_t)l0 * planeDense, s->scratch.p, (size_t)(l1 - l0) * planeDense * sizeof(float),
                                      cudaMemcpyDeviceToHost, s->stream));
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    }
}

// y-ghost exchange of one padded array set (h planes each way) on stream `st`
void exchangeHalos(ws_solver *s, const std::vector<float *> &arrays, int h, cudaStream_t st)
{
    if (s->d.nranks <= 1 || arrays.empty())
        return;
    WS_REQUIRE(s->ncclComm || s->extFn, WS_ESTATE, "multi-rank solver used before ws_comm_init");
    const size_t count = (size_t)h * s->plane;
    const int up = s->d.rank - 1, down = s->d.rank + 1;
    if (s->extFn) {
        // synchronous pairwise exchange; every rank talks to its upper neighbour first, so the chain cannot deadlock
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        WS_CUDA_CHECK(cudaStreamSynchronize(st));
        for (float *a : arrays) {
            float *origin = a + s->base - WS_PADX - (long long)(s->nzp > 1 ? WS_HALO : 0) * s->pitch;
            if (up >= 0)
                WS_REQUIRE(s->extFn(s->extUser, origin, origin - (long long)h * s->plane, count, up) == 0, WS_ECOMM, "external transport failed");
            if (down < s->d.nranks)
                WS_REQUIRE(s->extFn(s->extUser, origin + (long long)(s->nyl - h) * s->plane, origin + (long long)s->nyl * s->plane, count, down) == 0, WS_ECOMM,
                           "external transport failed");
        }
        return;
    }
    g_nccl.check(g_nccl.GroupStart(), "ncclGroupStart");
    for (float *a : arrays) {
        float *origin = a + s->base - WS_PADX - (long long)(s->nzp > 1 ? WS_HALO : 0) * s->pitch; // start of local plane 0
        if (up >= 0) {
            g_nccl.check(g_nccl.Send(origin, count, kNcclFloat, up, s->ncclComm, st), "ncclSend");
            g_nccl.check(g_nccl.Recv(origin - (long long)h * s->plane, count, kNcclFloat, up, s->ncclComm, st), "ncclRecv");
        }
        if (down < s->d.nranks) {
            g_nccl.check(g_nccl.Send(origin + (long long)(s->nyl - h) * s->plane, count, kNcclFloat, down, s->ncclComm, st), "ncclSend");
            g_nccl.check(g_nccl.Recv(origin + (long long)s->nyl * s->plane, count, kNcclFloat, down, s->ncclComm, st), "ncclRecv");
        }
    }
    g_nccl.check(g_nccl.GroupEnd(), "ncclGroupEnd");
}

#ifndef WS_EMULATE
// start of local plane 0 of an array (the ghost planes lie right before it and after plane nyl - 1)
float *planeZero(const ws_solver *s, float *a) { return a + s->base - WS_PADX - (long long)(s->nzp > 1 ? WS_HALO : 0) * s->pitch; }

// exchange `type` (0 = A: after the first half-step, 1 = B: after the step) of the wavefield slots `slots` over peer memory
void pushHalos(ws_solver *s, const std::vector<int> &slots, int type, cudaStream_t st)
{
    WS_REQUIRE((int)slots.size() <= WS_P2P_MAXF, WS_EINVAL, "too many fields in a halo exchange");
    WsPushArgs a{};
    const int h = s->h;
    a.nf = (int)slots.size();
    a.type = type;
    a.local = s->p2pLocal.p;
    a.count4 = (size_t)h * (size_t)s->plane / 4;
    const bool up = s->d.rank > 0, down = s->d.rank + 1 < s->d.nranks;
    a.peerFlag[0] = up ? s->peerFlags[0] : nullptr;
    a.peerFlag[1] = down ? s->peerFlags[1] : nullptr;
    for (int k = 0; k < a.nf; k++) {
        float *mine = planeZero(s, s->fld[slots[k]].p);
        const long long pos = s->ainfo.fldPos[slots[k]];
        WS_REQUIRE(pos >= 0, WS_ESTATE, "halo exchange over peer memory: the wavefield is not part of the arena");
        if (up) { // my planes [0, h) -> the upper neighbour's planes [nyl', nyl' + h)
            const long long peerTotal = s->plane * (long long)(s->peerNyl[0] + 2 * WS_HALO);
            a.src[0][k] = mine;
            a.dst[0][k] = planeZero(s, s->peerArena[0] + pos * peerTotal) + (long long)s->peerNyl[0] * s->plane;
        }
        if (down) { // my planes [nyl - h, nyl) -> the lower neighbour's planes [-h, 0)
            const long long peerTotal = s->plane * (long long)(s->peerNyl[1] + 2 * WS_HALO);
            a.src[1][k] = mine + (long long)(s->nyl - h) * s->plane;
            a.dst[1][k] = planeZero(s, s->peerArena[1] + pos * peerTotal) - (long long)h * s->plane;
        }
    }
    // enough thread blocks to keep the NVLink ports busy next to the interior kernel, few enough not to crowd it out
    const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>(32, (a.count4 + 2047) / 2048));
    kHaloPush<<<dim3(nb, 2 * a.nf), 256, 0, st>>>(a);
    s->launches++;
}

void waitHalos(ws_solver *s, int type, cudaStream_t st)
{
    kHaloWait<<<1, 1, 0, st>>>(s->p2pLocal.p, type, s->d.rank > 0 ? 1 : 0, s->d.rank + 1 < s->d.nranks ? 1 : 0);
    s->launches++;
}

// every rank publishes where its wavefield arena and its counters live; neighbours in the same process use the pointers directly
// (peer access enabled), neighbours in other processes open CUDA IPC handles.  Collective over the NCCL communicator; the exchange
// over peer memory is used only if EVERY rank could map its neighbours.
struct P2PInfo {
    int pid, device, nyl, pad;
    unsigned long long arena, flags;
    cudaIpcMemHandle_t arenaHandle, flagsHandle;
};
void p2pSetup(ws_solver *s)
{
    s->p2p = false;
    if (s->d.nranks <= 1 || !s->ncclComm || s->sparse)
        return;
    if (const char *e = getenv("WS_P2P"))
        if (atoi(e) == 0)
            return;
    float *arena = s->fldArena.p ? s->fldArena.p : s->arena.p;
    s->p2pLocal.alloc(P2P_NINTS);
    s->p2pLocal.zero(s->stream);
    P2PInfo mine{};
    mine.pid = (int)getpid();
    mine.device = s->d.device;
    mine.nyl = s->nyl;
    mine.arena = (unsigned long long)arena;
    mine.flags = (unsigned long long)s->p2pLocal.p;
    int okLocal = arena ? 1 : 0;
    if (okLocal && (cudaIpcGetMemHandle(&mine.arenaHandle, arena) != cudaSuccess || cudaIpcGetMemHandle(&mine.flagsHandle, s->p2pLocal.p) != cudaSuccess)) {
        cudaGetLastError();
        okLocal = 0;
    }
    const int n = s->d.nranks;
    DevBuf<char> sendb, recvb;
    sendb.alloc(sizeof(P2PInfo));
    recvb.alloc(sizeof(P2PInfo) * n);
    WS_CUDA_CHECK(cudaMemcpyAsync(sendb.p, &mine, sizeof(P2PInfo), cudaMemcpyHostToDevice, s->stream));
    g_nccl.check(g_nccl.AllGather(sendb.p, recvb.p, sizeof(P2PInfo), kNcclChar, s->ncclComm, s->stream), "ncclAllGather");
    std::vector<P2PInfo> all(n);
    WS_CUDA_CHECK(cudaMemcpyAsync(all.data(), recvb.p, sizeof(P2PInfo) * n, cudaMemcpyDeviceToHost, s->stream));
    WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    for (int d = 0; d < 2 && okLocal; d++) {
        const int peer = d == 0 ? s->d.rank - 1 : s->d.rank + 1;
        if (peer < 0 || peer >= n)
            continue;
        const P2PInfo &pi = all[peer];
        s->peerNyl[d] = pi.nyl;
        if (pi.arena == 0) {
            okLocal = 0;
        } else if (pi.pid == mine.pid) {
            int can = 0;
            if (pi.device != mine.device && (cudaDeviceCanAccessPeer(&can, mine.device, pi.device) != cudaSuccess || !can)) {
                okLocal = 0;
                continue;
            }
            if (pi.device != mine.device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    okLocal = 0;
                cudaGetLastError();
            }
            s->peerArena[d] = reinterpret_cast<float *>(pi.arena);
            s->peerFlags[d] = reinterpret_cast<int *>(pi.flags);
        } else {
            void *pa = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&pa, pi.arenaHandle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&pf, pi.flagsHandle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                if (pa)
                    cudaIpcCloseMemHandle(pa);
                okLocal = 0;
                continue;
            }
            s->peerArena[d] = static_cast<float *>(pa);
            s->peerFlags[d] = static_cast<int *>(pf);
            s->peerIpc[d] = true;
        }
    }
    // all or nothing: a rank that pushes needs a neighbour that waits for counters instead of an ncclRecv
    int bad = okLocal ? 0 : 1;
    WS_CUDA_CHECK(cudaMemcpyAsync(s->flag.p, &bad, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    g_nccl.check(g_nccl.AllReduce(s->flag.p, s->flag.p, 1, kNcclInt, kNcclMax, s->ncclComm, s->stream), "ncclAllReduce");
    WS_CUDA_CHECK(cudaMemcpyAsync(&bad, s->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    s->p2p = bad == 0;
}

// every rank has passed this point before any rank goes on (ws_reset in the peer-memory mode)
void rankBarrier(ws_solver *s)
{
    const int zero = 0;
    WS_CUDA_CHECK(cudaMemcpyAsync(s->flag.p, &zero, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    g_nccl.check(g_nccl.AllReduce(s->flag.p, s->flag.p, 1, kNcclInt, kNcclMax, s->ncclComm, s->stream), "ncclAllReduce");
    WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
}
#endif

void refreshParams(ws_solver *s)
{
    WsParams &P = s->P;
    P.nx = s->nx; P.nyl = s->nyl; P.nz = s->nz; P.gny = s->gny; P.gy0 = s->y0;
    P.pitch = s->pitch; P.nzp = s->nzp; P.plane = s->plane; P.base = s->base;
    P.dim = s->d.dim; P.eq = s->d.eq; P.q = s->q; P.h = s->h; P.L = s->L;
    P.free_surface = s->seismic ? s->d.free_surface : 0; // EM solvers ignore FreeSurface (ForwardSolver2Dtmem.cpp:31-33)
    P.damping = s->d.damping; P.W = s->W;
    P.ylo = 0; P.yhi = s->nyl;
    P.edge_policy = s->d.edge_policy;
    P.tab = s->tab.p;
    P.fldArena = s->fldArena.p;
    P.matArena = s->matArena.p;
    P.arenaStride = s->total;
    psiXLayout(s->nx, s->W, P.psiDX, P.psiPitchX);
    P.cxTab = s->cxTab.p;
    P.psiXArena = s->psiXArena.p;
    P.psiZArena = s->psiZArena.p;
    P.cax = s->cax.p; P.cbx = s->cbx.p; P.caxh = s->caxh.p; P.cbxh = s->cbxh.p;
    P.cay = s->cay.p; P.cby = s->cby.p; P.cayh = s->cayh.p; P.cbyh = s->cbyh.p;
    P.caz = s->caz.p; P.cbz = s->cbz.p; P.cazh = s->cazh.p; P.cbzh = s->cbzh.p;
    P.absCoeff = s->absCoeff.p;
    for (int k = 0; k < PSI_COUNT; k++) P.psi[k] = s->psi[k].p;
    for (int k = 0; k < F_COUNT; k++) P.fld[k] = s->fld[k].p;
    for (int k = 0; k < M_COUNT; k++) P.mat[k] = s->mat[k].p;
    P.sH = s->sH.p; P.sV = s->sV.p;
    for (int l = 0; l < 4; l++) { P.sRH[l] = s->sRH[l].p; P.sRV[l] = s->sRV[l].p; }
    P.DT = s->d.dt;
    P.fL = (float)s->L;
}

void refreshAcq(ws_solver *s)
{
    WsAcq &a = s->acq;
    a.nsrc = s->nsrc; a.nrec = s->nrec; a.nt = s->d.nt;
    a.srcType = s->srcType.p; a.srcOff = s->srcOff.p; a.srcSig = s->srcSig.p; a.srcStep = nullptr;
    a.recType = s->recType.p; a.recOff = s->recOff.p; a.seis = s->seis.p; a.recStep = nullptr;
    a.tdev = s->tdev.p;
}

void invalidateGraph(ws_solver *s)
{
    if (s->graphExec) {
        cudaGraphExecDestroy(s->graphExec);
        s->graphExec = nullptr;
        s->graphSteps = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// model preparation
// ---------------------------------------------------------------------------------------------------------------------
void launchPrep(ws_solver *s, int ylo, int yhi, dim3 &grid, dim3 &block) { s->gridFor(ylo, yhi, grid, block); s->launches++; }

void requireMat(ws_solver *s, int slot, const char *name)
{
    WS_REQUIRE(s->mat[slot].p && s->matGiven[slot], WS_ESTATE, std::string("model parameter '") + name + "' is not set");
}

void prepareSeismic(ws_solver *s)
{
    const ws_desc &d = s->d;
    const bool needP = d.eq == WS_EQ_ACOUSTIC || d.eq == WS_EQ_ELASTIC || d.eq == WS_EQ_VISCOELASTIC;
    const bool needS = d.eq != WS_EQ_ACOUSTIC;
    const bool d3 = d.dim == 3;
    dim3 grid, block;
    const int elo = -WS_HALO, ehi = s->nyl + WS_HALO; // extended range (needs ghost planes of the raw parameters)
    const wsprep::Geo gExt = s->geo(elo, ehi), gLoc = s->geo(0, s->nyl);
    // relaxed-modulus scaling sum, Viscoelastic.cpp:657-665
    float sum = 0;
    if (s->visco) {
        float w_ref = (float)(2.0 * M_PI * d.fc_cpml);
        for (int l = 0; l < s->L; l++) {
            float tauSigma = (float)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
            sum += (float)(w_ref * w_ref * tauSigma * tauSigma / (1.0 + w_ref * w_ref * tauSigma * tauSigma));
        }
    }
    if (needP && !s->matGiven[M_PW]) {
        requireMat(s, M_VP, "velocityP");
        requireMat(s, M_RHO, "density");
        if (s->visco)
            requireMat(s, M_TAUP, "tauP");
        float *out = matBuf(s, M_PW);
        launchPrep(s, elo, ehi, grid, block);
        WS_LAUNCH(wsprep::kModulus, grid, block, 0, s->stream, gExt, s->mat[M_VP].p, s->mat[M_RHO].p, s->visco ? s->mat[M_TAUP].p : nullptr, sum, out);
    }
    if (needS && !s->matGiven[M_MU]) {
        requireMat(s, M_VS, "velocityS");
        requireMat(s, M_RHO, "density");
        if (s->visco)
            requireMat(s, M_TAUS, "tauS");
        float *out = matBuf(s, M_MU);
        launchPrep(s, elo, ehi, grid, block);
        WS_LAUNCH(wsprep::kModulus, grid, block, 0, s->stream, gExt, s->mat[M_VS].p, s->mat[M_RHO].p, s->visco ? s->mat[M_TAUS].p : nullptr, sum, out);
    }
    if (needS) {
        // calcAveragedSWaveModulus clamps the modulus itself first (ModelparameterSeismic.cpp:424)
        launchPrep(s, elo, ehi, grid, block);
        WS_LAUNCH(wsprep::kClampLess, grid, block, 0, s->stream, gExt, matBuf(s, M_MU), 1.0f, 1.0f);
    }
    auto avg2 = [&](int in, int out, int axis, int mode) {
        float *o = matBuf(s, out);
        launchPrep(s, 0, s->nyl, grid, block);
        WS_LAUNCH(wsprep::kAvg2, grid, block, 0, s->stream, gLoc, s->mat[in].p, o, axis, mode);
    };
    auto avg4 = [&](int in, int out, int axA, int axB, int mode) {
        float *o = matBuf(s, out);
        launchPrep(s, 0, s->nyl, grid, block);
        WS_LAUNCH(wsprep::kAvg4, grid, block, 0, s->stream, gLoc, s->mat[in].p, o, axA, axB, mode);
    };
    if (d.eq == WS_EQ_SH || d.eq == WS_EQ_VISCOSH) {
        if (!s->matGiven[M_INVRHO]) {
            requireMat(s, M_RHO, "density");
            float *o = matBuf(s, M_INVRHO);
            launchPrep(s, 0, s->nyl, grid, block);
            WS_LAUNCH(wsprep::kInverse, grid, block, 0, s->stream, gLoc, s->mat[M_RHO].p, o);
        }
        if (!s->matGiven[M_MUXZ]) avg2(M_MU, M_MUXZ, 0, 2); // SH.cpp:424-430
        if (!s->matGiven[M_MUYZ]) avg2(M_MU, M_MUYZ, 1, 2);
        if (s->visco) {
            requireMat(s, M_TAUS, "tauS");
            if (!s->matGiven[M_TSXZ]) avg2(M_TAUS, M_TSXZ, 0, 0);
            if (!s->matGiven[M_TSYZ]) avg2(M_TAUS, M_TSYZ, 1, 0);
        }
        return;
    }
    if (!s->matGiven[M_RIX] || !s->matGiven[M_RIY] || (d3 && !s->matGiven[M_RIZ]))
        requireMat(s, M_RHO, "density");
    if (!s->matGiven[M_RIX]) avg2(M_RHO, M_RIX, 0, 1);
    if (!s->matGiven[M_RIY]) avg2(M_RHO, M_RIY, 1, 1);
    if (d3 && !s->matGiven[M_RIZ]) avg2(M_RHO, M_RIZ, 2, 1);
    if (d.eq == WS_EQ_ACOUSTIC)
        return;
    if (s->visco) {
        requireMat(s, M_TAUS, "tauS");
        requireMat(s, M_TAUP, "tauP");
    }
    if (!s->matGiven[M_MUXY]) avg4(M_MU, M_MUXY, 0, 1, 2);
    if (s->visco && !s->matGiven[M_TSXY]) avg4(M_TAUS, M_TSXY, 0, 1, 0);
    if (d3) {
        if (!s->matGiven[M_MUXZ]) avg4(M_MU, M_MUXZ, 0, 2, 2);
        if (!s->matGiven[M_MUYZ]) avg4(M_MU, M_MUYZ, 2, 1, 2);
        if (s->visco) {
            if (!s->matGiven[M_TSXZ]) avg4(M_TAUS, M_TSXZ, 0, 2, 0);
            if (!s->matGiven[M_TSYZ]) avg4(M_TAUS, M_TSYZ, 2, 1, 0);
        }
    }
}

void prepareEM(ws_solver *s)
{
    const ws_desc &d = s->d;
    const bool d3 = d.dim == 3;
    dim3 grid, block;
    const wsprep::Geo gLoc = s->geo(0, s->nyl);
    std::vector<float> relaxTime(s->L);
    for (int l = 0; l < s->L; l++)
        relaxTime[l] = (float)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
    const float DT = d.dt;
    wsprep::EmCoef c{};
    c.L = s->L;
    c.DT = DT;
    c.eps0 = (float)8.8541878176e-12; // Modelparameter.hpp:359-360
    {
        float sum = 0;
        for (int l = 0; l < s->L; l++)
            sum += (float)(1.0 / relaxTime[l]);
        if (s->L)
            sum /= s->L;
        c.sumInvRelax = sum;
    }
    for (int l = 0; l < s->L; l++) {
        float tempValue = (float)(1 / (1 + 0.5 * DT / relaxTime[l]));
        tempValue /= (s->L * relaxTime[l] * relaxTime[l]);
        c.cdScalar[l] = tempValue;
        s->P.Cc[l] = (float)((1 - 0.5 * DT / relaxTime[l]) / (1 + 0.5 * DT / relaxTime[l]));
    }
    requireMat(s, M_MUM, "magneticPermeability");
    requireMat(s, M_EPS, "dielectricPermittivity");
    requireMat(s, M_SIG, "electricConductivity");
    if (s->visco) {
        requireMat(s, M_TAUEPS, "tauDielectricPermittivity");
        requireMat(s, M_TAUSIG, "tauElectricConductivity");
    }
    auto avg2 = [&](const float *in, float *out, int axis, int mode) {
        launchPrep(s, 0, s->nyl, grid, block);
        WS_LAUNCH(wsprep::kAvg2, grid, block, 0, s->stream, gLoc, in, out, axis, mode);
    };
    auto avg4 = [&](const float *in, float *out, int axA, int axB, int mode) {
        launchPrep(s, 0, s->nyl, grid, block);
        WS_LAUNCH(wsprep::kAvg4, grid, block, 0, s->stream, gLoc, in, out, axA, axB, mode);
    };
    auto coef = [&](const float *eps, const float *sig, const float *te, const float *ts, int axis) {
        for (int l = 0; l < s->L; l++)
            c.cd[l] = matBuf(s, M_CD0 + 3 * l + axis);
        float *Ca = matBuf(s, M_CAX + axis), *Cb = matBuf(s, M_CBX + axis);
        launchPrep(s, 0, s->nyl, grid, block);
        WS_LAUNCH(wsprep::kEmCoefficients, grid, block, 0, s->stream, gLoc, eps, sig, te, ts, s->visco ? 1 : 0, c, Ca, Cb);
    };
    if (d.eq == WS_EQ_TMEM || d.eq == WS_EQ_VISCOTMEM) {
        if (!s->matGiven[M_MIXZ]) avg2(s->mat[M_MUM].p, matBuf(s, M_MIXZ), 0, 1); // ViscoTMEM.cpp:442-450
        if (!s->matGiven[M_MIYZ]) avg2(s->mat[M_MUM].p, matBuf(s, M_MIYZ), 1, 1);
        coef(s->mat[M_EPS].p, s->mat[M_SIG].p, s->mat[M_TAUEPS].p, s->mat[M_TAUSIG].p, RC_Z);
        return;
    }
    // EMEM.cpp:378-392, ViscoEMEM.cpp:409-423
    if (!s->matGiven[M_MIXY]) avg4(s->mat[M_MUM].p, matBuf(s, M_MIXY), 0, 1, 1);
    if (d3) {
        if (!s->matGiven[M_MIXZ]) avg4(s->mat[M_MUM].p, matBuf(s, M_MIXZ), 0, 2, 1);
        if (!s->matGiven[M_MIYZ]) avg4(s->mat[M_MUM].p, matBuf(s, M_MIYZ), 2, 1, 1);
    }
    DevBuf<float> e, g, te, ts;
    e.alloc((size_t)s->total); g.alloc((size_t)s->total);
    e.zero(s->stream); g.zero(s->stream);
    if (s->visco) { te.alloc((size_t)s->total); ts.alloc((size_t)s->total); te.zero(s->stream); ts.zero(s->stream); }
    const int nax = d3 ? 3 : 2;
    for (int a = 0; a < nax; a++) {
        avg2(s->mat[M_EPS].p, e.p, a, 0);
        avg2(s->mat[M_SIG].p, g.p, a, 0);
        if (s->visco) {
            avg2(s->mat[M_TAUEPS].p, te.p, a, 0);
            avg2(s->mat[M_TAUSIG].p, ts.p, a, 0);
        }
        coef(e.p, g.p, te.p, ts.p, a);
    }
    WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
}

void prepareBoundaries(ws_solver *s)
{
    const ws_desc &d = s->d;
    // derivative tables = Derivatives::init (FDTD3D.cpp:51-61)
    const bool fsTables = s->seismic && d.free_surface == 1;
    std::vector<float> tab = wstab::buildTables(s->q, d.edge_policy, fsTables, s->nx, s->gny, s->nz, d.dim, d.dh, d.dt);
    s->tab.upload(tab);
    {
        // interior rows as kernel parameters (constant bank) for the tiled kernels: forward taps are entries 1..q
        const int rows = 2 * s->h + 1, taps = s->q + 1;
        const float *xf = &tab[((size_t)OP_XF * rows + s->h) * taps], *yf = &tab[((size_t)(fsTables ? OP_YF_FS : OP_YF) * rows + s->h) * taps];
        for (int j = 0; j < WS_MAXQ; j++) {
            s->P.cw[j] = j < s->q ? xf[1 + j] : 0.0f;
            s->P.cwy[j] = j < s->q ? yf[1 + j] : 0.0f;
        }
    }
    if (d.damping == 2) {
        // CPML*.init (CPML3D.cpp:222-368): same 1-D profile on every axis (regular grid)
        wstab::CpmlAxis c = wstab::buildCpmlAxis(s->W, d.npower, d.fc_cpml, d.vmax_cpml, d.dt, d.dh);
        s->cax.upload(c.a); s->cbx.upload(c.b); s->caxh.upload(c.ah); s->cbxh.upload(c.bh);
        s->cay.upload(c.a); s->cby.upload(c.b); s->cayh.upload(c.ah); s->cbyh.upload(c.bh);
        s->caz.upload(c.a); s->cbz.upload(c.b); s->cazh.upload(c.ah); s->cbzh.upload(c.bh);
        {
            // coefficient rows in slab-row order for the tiled kernels: {a, b, a half, b half}[PX], zero on the padding
            int D, PX;
            psiXLayout(s->nx, s->W, D, PX);
            std::vector<float> t(4 * (size_t)PX, 0.0f);
            for (int x = 0; x < s->nx; x++) {
                const int k = wsCpmlIndex(x, s->nx, s->W);
                if (k < 0)
                    continue;
                const int kp = wsPsiXIndex(x, s->W, D);
                t[kp] = c.a[k]; t[PX + kp] = c.b[k]; t[2 * PX + kp] = c.ah[k]; t[3 * PX + kp] = c.bh[k];
            }
            s->cxTab.upload(t);
        }
        if (s->fldArena.p) {
            // 3D elastic: the memory variables of the x and of the z terms are slots of two arenas, in the order the
            // tiled kernels fetch them (ws_kernels_fast.cu: velocity half-step roles, then stress half-step groups)
            static const int ox[6] = {PSI_SXX_X, PSI_SXY_X, PSI_SXZ_X, PSI_VXX, PSI_VYX, PSI_VZX};
            static const int oz[6] = {PSI_SXZ_Z, PSI_SYZ_Z, PSI_SZZ_Z, PSI_VXZ, PSI_VYZ, PSI_VZZ};
            const size_t nxs = psiSize(s, 0), nzs = psiSize(s, 2);
            s->psiXArena.alloc(6 * nxs);
            s->psiXArena.zero(s->stream);
            s->psiZArena.alloc(6 * nzs);
            s->psiZArena.zero(s->stream);
            for (int k = 0; k < 6; k++) {
                s->psi[ox[k]].borrow(s->psiXArena.p + k * nxs, nxs);
                s->psi[oz[k]].borrow(s->psiZArena.p + k * nzs, nzs);
            }
        }
        for (auto &pa : psiFor(d)) {
            if (!s->psi[pa.first].p)
                s->psi[pa.first].alloc(psiSize(s, pa.second));
            s->psi[pa.first].zero(s->stream);
        }
    }
    if (d.damping == 1)
        s->absCoeff.upload(wstab::buildAbsCoeff(s->W, d.damping_coeff));
    // free surface scalings (only the rank that owns y = 0 evaluates them)
    if (s->seismic && d.free_surface == 1 && (d.eq == WS_EQ_ELASTIC || d.eq == WS_EQ_VISCOELASTIC)) {
        const size_t ns = (size_t)s->nx * s->nz;
        s->sH.alloc(ns); s->sV.alloc(ns);
        s->sH.zero(s->stream); s->sV.zero(s->stream);
        for (int l = 0; l < s->L; l++) {
            s->sRH[l].alloc(ns); s->sRV[l].alloc(ns);
            s->sRH[l].zero(s->stream); s->sRV[l].zero(s->stream);
        }
        if (s->y0 == 0) {
            dim3 block = s->nz > 1 ? dim3(64, 4, 1) : dim3(128, 1, 1);
            dim3 grid((s->nx + block.x - 1) / block.x, (s->nz + block.y - 1) / block.y, 1);
            s->flag.zero(s->stream);
            s->launches++;
            if (d.eq == WS_EQ_ELASTIC) {
                WS_LAUNCH(wsprep::kFreeSurfaceElastic, grid, block, 0, s->stream, s->geo(0, 1), s->mat[M_PW].p, s->mat[M_MU].p, s->sH.p, s->sV.p, s->flag.p);
            } else {
                wsprep::ViscoFS v{};
                v.L = s->L;
                v.fL = (float)s->L;
                for (int l = 0; l < s->L; l++) {
                    v.relaxTime[l] = (float)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
                    v.viscoCoeff2[l] = (float)(1.0 / (1.0 + d.dt / (2.0 * v.relaxTime[l])));
                    v.sRH[l] = s->sRH[l].p;
                    v.sRV[l] = s->sRV[l].p;
                }
                WS_LAUNCH(wsprep::kFreeSurfaceVisco, grid, block, 0, s->stream, s->geo(0, 1), s->mat[M_PW].p, s->mat[M_MU].p, s->mat[M_TAUP].p, s->mat[M_TAUS].p,
                                                                             s->sH.p, s->sV.p, v, s->flag.p);
            }
            int bad = 0;
            WS_CUDA_CHECK(cudaMemcpyAsync(&bad, s->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            WS_REQUIRE(!bad, WS_EINVAL, "S wave modulus can't be zero when using image method");
        }
    }
    if (s->visco && s->seismic) {
        // ForwardSolver3Dviscoelastic.cpp:55-66
        for (int l = 0; l < s->L; l++) {
            float relaxationTime = (float)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
            s->P.invRelaxTime[l] = (float)(1.0 / relaxationTime);
            s->P.viscoCoeff1[l] = (float)(1.0 - d.dt / (2.0 * relaxationTime));
            s->P.viscoCoeff2[l] = (float)(1.0 / (1.0 + d.dt / (2.0 * relaxationTime)));
        }
        s->P.DThalf = (float)(d.dt / 2.0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// time stepping
// ---------------------------------------------------------------------------------------------------------------------
void firstHalfFields(const ws_solver *s, int f[3])
{
    f[0] = f[1] = f[2] = -1;
    switch (s->d.eq) {
    case WS_EQ_ACOUSTIC:
    case WS_EQ_ELASTIC:
    case WS_EQ_VISCOELASTIC:
        f[0] = F_VX; f[1] = F_VY; if (s->d.dim == 3) f[2] = F_VZ;
        break;
    case WS_EQ_SH:
    case WS_EQ_VISCOSH: f[0] = F_VZ; break;
    case WS_EQ_TMEM:
    case WS_EQ_VISCOTMEM: f[0] = F_HX; f[1] = F_HY; break;
    default:
        f[0] = F_HZ; if (s->d.dim == 3) { f[1] = F_HX; f[2] = F_HY; }
        break;
    }
}

void launchPass(ws_solver *s, int pass, int ylo, int yhi)
{
    if (yhi <= ylo)
        return;
    WsParams P = s->P;
    P.ylo = ylo;
    P.yhi = yhi;
    if (s->useFast || (s->useFastA && pass == 0)) {
        const int n = wsLaunchFast(P, pass, s->stream);
        if (n > 0) {
            s->launches += n;
            return;
        }
    }
#ifndef WS_EMULATE
    if (s->useTile) {
        const int n = wsLaunchTile(P, pass, s->tileProg[pass], s->stream);
        if (n > 0) {
            s->launches += n;
            return;
        }
    }
    if (s->useTma) {
        const int n = wsLaunchTma(P, pass, s->tmaProg[pass], s->tmaNL[pass], s->stream);
        if (n > 0) {
            s->launches += n;
            return;
        }
    }
#endif
    if (s->useMarch) {
        const int n = wsLaunchMarch(P, pass, s->stream);
        if (n > 0) {
            s->launches += n;
            return;
        }
    }
    wsLaunchGeneral(P, s->exact, pass, s->stream);
    s->launches++;
}

void launchAcquisition(ws_solver *s, const float *srcStepDev, float *recStepDev)
{
    WsAcq a = s->acq;
    a.srcStep = srcStepDev;
    a.recStep = recStepDev;
#ifndef WS_EMULATE /* (the host emulation runs the threads of a block one after the other: no block barriers) */
    if (s->nsrc <= WS_ACQ_MAX_SRC && s->nrec <= WS_ACQ_MAX_REC && !(s->srcSequential && s->nsrc > 64)) {
        static const bool pdl = !(getenv("WS_PDL") && atoi(getenv("WS_PDL")) == 0);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1);
        cfg.blockDim = dim3(WS_ACQ_THREADS);
        cfg.stream = s->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl && s->useTile) ? 1 : 0; // (only next to kernels that wait themselves: the 2-D tile kernels)
        cudaLaunchKernelEx(&cfg, kAcquisition, s->P, a, s->srcSequential ? 1 : 0);
        s->launches++;
        return;
    }
#endif
    if (s->nsrc > 0) {
        if (s->srcSequential)
            WS_LAUNCH(kSources, 1, 32, 0, s->stream, s->P, a, 1);
        else
            WS_LAUNCH(kSources, (s->nsrc + 127) / 128, 128, 0, s->stream, s->P, a, 0);
        s->launches++;
    }
    if (s->nrec > 0) {
        WS_LAUNCH(kReceivers, (s->nrec + 127) / 128, 128, 0, s->stream, s->P, a);
        s->launches++;
    }
    WS_LAUNCH(kAdvance, 1, 1, 0, s->stream, s->tdev.p);
    s->launches++;
}

// Simulation.cpp:455-456 `*wavefields *= compensation` (all components, after the step incl. source injection and recording)
void launchStepScaling(ws_solver *s)
{
    if (!s->stepScale.p)
        return;
    WsFieldTable t{};
    for (int k = 0; k < F_COUNT; k++)
        if (s->fld[k].p)
            t.p[t.n++] = s->fld[k].p;
    const size_t first = s->sparse ? 0 : (size_t)WS_HALO * (size_t)s->plane, n = s->sparse ? (size_t)s->total : (size_t)s->nyl * (size_t)s->plane;
    WS_LAUNCH(kWfScaleVecAll, (unsigned)((n + 255) / 256), 256, 0, s->stream, t, s->stepScale.p, first, n);
    s->launches++;
}

// one reference time step = ForwardSolver::run(...), enqueued asynchronously
// ---------------------------------------------------------------------------------------------------------------------
// operator-given mode (ws_kernels_sparse.cuh)
// ---------------------------------------------------------------------------------------------------------------------
wssparse::Params sparseParams(ws_solver *s)
{
    wssparse::Params Q{};
    Q.n = s->nx;
    Q.dim = s->d.dim;
    for (int k = 0; k < wssparse::SP_NOPS; k++) {
        Q.col[k] = s->spCol[k].p;
        Q.val[k] = s->spVal[k].p;
        Q.taps[k] = s->spTaps[k];
    }
    for (int a = 0; a < 3; a++) {
        Q.cpK[a] = s->spCpK[a].p;
        Q.ca[a] = s->spCa[a].p; Q.cb[a] = s->spCb[a].p; Q.cah[a] = s->spCah[a].p; Q.cbh[a] = s->spCbh[a].p;
    }
    for (int k = 0; k < wssparse::SPSI_COUNT; k++)
        Q.psi[k] = s->spPsi[k].p;
    Q.vx = s->fld[F_VX].p; Q.vy = s->fld[F_VY].p; Q.vz = s->fld[F_VZ].p; Q.p = s->fld[F_P].p;
    Q.rix = s->mat[M_RIX].p; Q.riy = s->mat[M_RIY].p; Q.riz = s->mat[M_RIZ].p; Q.pw = s->mat[M_PW].p;
    return Q;
}

void sparseInterpolate(ws_solver *s, int which, float *field)
{
    ws_solver::Interp &I = s->spInterp[which];
    if (I.nrows == 0)
        return;
    const unsigned nb = (unsigned)((I.nrows + 255) / 256);
    if (s->exact)
        WS_LAUNCH(wssparse::kInterpGather<true>, nb, 256, 0, s->stream, I.nrows, I.taps, I.cols.p, I.vals.p, field, s->spTmp.p);
    else
        WS_LAUNCH(wssparse::kInterpGather<false>, nb, 256, 0, s->stream, I.nrows, I.taps, I.cols.p, I.vals.p, field, s->spTmp.p);
    WS_LAUNCH(wssparse::kInterpScatter, nb, 256, 0, s->stream, I.nrows, I.rows.p, s->spTmp.p, field);
    s->launches += 2;
}

// one time step on an irregular grid: ForwardSolver2Dacoustic.cpp:121-190 / ForwardSolver3Dacoustic.cpp:131-229 incl. the
// interpolation of the interface planes after every update
void enqueueSparseStep(ws_solver *s, const float *srcStepDev, float *recStepDev, cudaEvent_t *ev)
{
    const wssparse::Params Q = sparseParams(s);
    const unsigned nb = (unsigned)((Q.n + 255) / 256);
    if (ev)
        WS_CUDA_CHECK(cudaEventRecord(ev[0], s->stream));
    if (s->exact)
        WS_LAUNCH(wssparse::kVelAcoustic<true>, nb, 256, 0, s->stream, Q);
    else
        WS_LAUNCH(wssparse::kVelAcoustic<false>, nb, 256, 0, s->stream, Q);
    s->launches++;
    sparseInterpolate(s, 1, Q.vx);
    sparseInterpolate(s, 0, Q.vy);
    if (Q.dim == 3)
        sparseInterpolate(s, 2, Q.vz);
    if (ev) {
        WS_CUDA_CHECK(cudaEventRecord(ev[1], s->stream));
        WS_CUDA_CHECK(cudaEventRecord(ev[2], s->stream));
    }
    if (s->exact)
        WS_LAUNCH(wssparse::kPressure<true>, nb, 256, 0, s->stream, Q);
    else
        WS_LAUNCH(wssparse::kPressure<false>, nb, 256, 0, s->stream, Q);
    s->launches++;
    if (s->spAbsN > 0) { // DampingBoundary.apply(p, vX, vY[, vZ]) comes before the interpolation of p (ForwardSolver2Dacoustic.cpp:176-185)
        const unsigned nba = (unsigned)((s->spAbsN + 255) / 256);
        if (s->exact)
            WS_LAUNCH(wssparse::kAbsDamp<true>, nba, 256, 0, s->stream, s->spAbsN, s->spAbsIdx.p, s->spAbsVal.p, Q.p, Q.vx, Q.vy, Q.dim == 3 ? Q.vz : nullptr);
        else
            WS_LAUNCH(wssparse::kAbsDamp<false>, nba, 256, 0, s->stream, s->spAbsN, s->spAbsIdx.p, s->spAbsVal.p, Q.p, Q.vx, Q.vy, Q.dim == 3 ? Q.vz : nullptr);
        s->launches++;
    }
    sparseInterpolate(s, 0, Q.p);
    if (s->spSurfN > 0) {
        WS_LAUNCH(wssparse::kSurfaceZero, (unsigned)((s->spSurfN + 255) / 256), 256, 0, s->stream, s->spSurfN, s->spSurf.p, Q.p);
        s->launches++;
    }
    if (ev)
        WS_CUDA_CHECK(cudaEventRecord(ev[3], s->stream));
    launchAcquisition(s, srcStepDev, recStepDev);
    launchStepScaling(s);
}

// haloLanded: the halo exchange of the previous step is known to have completed (first step of a captured graph: the
// wait happened on the stream before the graph was launched)
void enqueueStep(ws_solver *s, const float *srcStepDev, float *recStepDev, cudaEvent_t *ev /* 4 events or null */, bool haloLanded = false)
{
    if (s->sparse) {
        enqueueSparseStep(s, srcStepDev, recStepDev, ev);
        return;
    }
    const bool multi = s->d.nranks > 1;
#ifdef WS_EMULATE
    const bool p2p = false;
    auto pushHalos = [](ws_solver *, const std::vector<int> &, int, cudaStream_t) {};
    auto waitHalos = [](ws_solver *, int, cudaStream_t) {};
#else
    const bool p2p = s->p2p;
#endif
    const int h = s->h, n = s->nyl;
    const cudaEvent_t evCompute = s->capturing ? s->evComputeG : s->evCompute, evComm = s->capturing ? s->evCommG : s->evComm;
    int fA[3];
    firstHalfFields(s, fA);
    auto gather = [&](const std::vector<int> &slots) {
        std::vector<float *> v;
        for (int k : slots)
            v.push_back(s->fld[k].p);
        return v;
    };
    if (ev)
        WS_CUDA_CHECK(cudaEventRecord(ev[0], s->stream));
    if (!multi) {
        launchPass(s, 0, 0, n);
    } else {
        // interior first (needs no ghost planes), then the edge slabs once the previous exchange has landed
        launchPass(s, 0, h, n - h);
        if (p2p)
            waitHalos(s, 1, s->stream); // the counters tell when exchange B of the step before has landed
        else if (!haloLanded)
            WS_CUDA_CHECK(cudaStreamWaitEvent(s->stream, evComm, 0));
        launchPass(s, 0, 0, std::min(h, n));
        launchPass(s, 0, std::max(n - h, h), n);
        WS_CUDA_CHECK(cudaEventRecord(evCompute, s->stream));
        WS_CUDA_CHECK(cudaStreamWaitEvent(s->commStream, evCompute, 0));
        if (p2p)
            pushHalos(s, s->exchA, 0, s->commStream);
        else
            exchangeHalos(s, gather(s->exchA), h, s->commStream);
        WS_CUDA_CHECK(cudaEventRecord(evComm, s->commStream));
    }
    if (ev) {
        WS_CUDA_CHECK(cudaEventRecord(ev[1], s->stream));
        WS_CUDA_CHECK(cudaEventRecord(ev[2], s->stream));
    }
    if (!multi) {
        launchPass(s, 1, 0, n);
    } else {
        launchPass(s, 1, h, n - h);
        if (p2p)
            waitHalos(s, 0, s->stream);
        else
            WS_CUDA_CHECK(cudaStreamWaitEvent(s->stream, evComm, 0));
        launchPass(s, 1, 0, std::min(h, n));
        launchPass(s, 1, std::max(n - h, h), n);
    }
    if (ev)
        WS_CUDA_CHECK(cudaEventRecord(ev[3], s->stream));
    if (s->d.damping == 1) {
        WsParams P = s->P;
        wsLaunchAbsFirstHalf(P, s->exact, fA[0], fA[1], fA[2], s->stream);
        s->launches++;
    }
    launchAcquisition(s, srcStepDev, recStepDev);
    launchStepScaling(s);
    if (multi) {
        WS_CUDA_CHECK(cudaEventRecord(evCompute, s->stream));
        WS_CUDA_CHECK(cudaStreamWaitEvent(s->commStream, evCompute, 0));
        if (p2p)
            pushHalos(s, s->exchB, 1, s->commStream);
        else
            exchangeHalos(s, gather(s->exchB), h, s->commStream);
        WS_CUDA_CHECK(cudaEventRecord(evComm, s->commStream));
    }
}

void setTime(ws_solver *s, int t)
{
    WS_REQUIRE(t >= 0 && t < s->d.nt, WS_EINVAL, "time step out of range");
    WS_CUDA_CHECK(cudaMemcpyAsync(s->tdev.p, &t, sizeof(int), cudaMemcpyHostToDevice, s->stream));
}

template <typename F>
int guarded(F fn)
{
    try {
        fn();
        return WS_OK;
    } catch (const WsError &e) {
        g_lastError = e.msg;
        return e.code;
    } catch (const std::exception &e) {
        g_lastError = e.what();
        return WS_EINVAL;
    }
}

} // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char *ws_last_error(void) { return g_lastError.c_str(); }
const char *ws_version(void) { return "wavesim-b200 0.1 (sm_100a)"; }
int ws_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

size_t ws_estimate_memory(const ws_desc *desc)
{
    if (!desc)
        return 0;
    try {
        validateDesc(*desc);
    } catch (...) {
        return 0;
    }
    int y0, nyl;
    slabRange(*desc, desc->rank, y0, nyl);
    const int nz = desc->dim == 2 ? 1 : desc->nz;
    const size_t pitch = ((size_t)desc->nx + WS_PADX + WS_HALO + 31) / 32 * 32;
    const size_t nzp = nz > 1 ? nz + 2 * WS_HALO : 1;
    const size_t total = pitch * nzp * (nyl + 2 * WS_HALO);
    const size_t nf = fieldsFor(*desc).f.size();
    size_t nm = 8;
    switch (desc->eq) {
    case WS_EQ_ACOUSTIC: nm = 2 + 1 + desc->dim; break;
    case WS_EQ_ELASTIC: nm = 3 + 2 + desc->dim + (desc->dim == 3 ? 3 : 1); break;
    case WS_EQ_VISCOELASTIC: nm = 5 + 2 + desc->dim + 2 * (desc->dim == 3 ? 3 : 1); break;
    case WS_EQ_SH: nm = 2 + 4; break;
    case WS_EQ_VISCOSH: nm = 3 + 6; break;
    default: nm = 5 + 3 + 2 * desc->dim + desc->n_relax * desc->dim; break;
    }
    size_t bytes = (nf + nm) * total * sizeof(float);
    if (desc->damping == 2) {
        const size_t W2 = 2 * (size_t)desc->boundary_width;
        for (auto &pa : psiFor(*desc)) {
            int D, PX;
            psiXLayout(desc->nx, desc->boundary_width, D, PX);
            const size_t n = pa.second == 0 ? (size_t)nyl * nz * PX : (pa.second == 1 ? W2 * nz * desc->nx : (size_t)nyl * W2 * desc->nx);
            bytes += n * sizeof(float);
        }
    }
    return bytes;
}

static int createImpl(const ws_desc *desc, long long sparseN, ws_solver **out)
{
    return guarded([&] {
        WS_REQUIRE(desc && out, WS_EINVAL, "null argument");
        ws_desc local = *desc;
        if (sparseN > 0) {
            // operator-given mode: the model vector is one row of n_points values; geometry-dependent keys are the caller's business
            WS_REQUIRE(sparseN < (1LL << 31), WS_EINVAL, "n_points exceeds int32 indices (scai::IndexType)");
            WS_REQUIRE(local.eq == WS_EQ_ACOUSTIC, WS_EINVAL, "operator-given mode (variable grid) is available for the acoustic solvers");
            WS_REQUIRE(local.nranks <= 1, WS_EINVAL, "operator-given mode runs on one GPU per shot");
            WS_REQUIRE(local.damping >= 0 && local.damping <= 2, WS_EINVAL, "DampingBoundary must be 0, 1 or 2");
            local.nx = (int32_t)sparseN;
            local.ny = 1;
            local.nz = 1;
            local.boundary_width = 0;
            if (local.fd_order < 2 || local.fd_order > WS_MAXQ || local.fd_order % 2)
                local.fd_order = 2; // per-layer orders live in the operators
            const int keepDamping = local.damping;
            local.damping = 0;
            validateDesc(local);
            local.damping = keepDamping;
        } else
            validateDesc(local);
        desc = &local;
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        WS_REQUIRE(e == cudaSuccess && ndev > 0, WS_ECUDA,
                   std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
        WS_REQUIRE(desc->device >= 0 && desc->device < ndev, WS_EINVAL, "device ordinal out of range");
        ws_solver *s = new ws_solver();
        try {
            s->d = *desc;
            if (s->d.dim == 2)
                s->d.nz = 1;
            setDevice(s);
            s->nx = s->d.nx; s->gny = s->d.ny; s->nz = s->d.nz;
            slabRange(s->d, s->d.rank, s->y0, s->nyl);
            s->q = s->d.fd_order; s->h = s->q / 2;
            s->seismic = s->d.eq <= WS_EQ_VISCOSH;
            s->visco = eqIsVisco(s->d.eq);
            s->L = s->visco ? s->d.n_relax : 0;
            s->W = s->d.damping ? s->d.boundary_width : 0;
            s->exact = s->d.exact_arith != 0;
            s->pitch = (s->nx + WS_PADX + WS_HALO + 31) / 32 * 32;
            s->nzp = s->nz > 1 ? s->nz + 2 * WS_HALO : 1;
            s->plane = (long long)s->pitch * s->nzp;
            s->total = s->plane * (s->nyl + 2 * WS_HALO);
            s->base = WS_PADX + (long long)(s->nzp > 1 ? WS_HALO : 0) * s->pitch + (long long)WS_HALO * s->plane;
            if (sparseN > 0) { // no pads: the arrays ARE the model vectors
                s->sparse = true;
                s->pitch = s->nx;
                s->plane = s->nx;
                s->total = s->nx;
                s->base = 0;
            }
            WS_CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
            WS_CUDA_CHECK(cudaStreamCreateWithFlags(&s->commStream, cudaStreamNonBlocking));
            WS_CUDA_CHECK(cudaEventCreateWithFlags(&s->evCompute, cudaEventDisableTiming));
            WS_CUDA_CHECK(cudaEventCreateWithFlags(&s->evComm, cudaEventDisableTiming));
            WS_CUDA_CHECK(cudaEventCreateWithFlags(&s->evComputeG, cudaEventDisableTiming));
            WS_CUDA_CHECK(cudaEventCreateWithFlags(&s->evCommG, cudaEventDisableTiming));
            WS_CUDA_CHECK(cudaEventRecord(s->evComm, s->commStream));
            if ((s->d.eq == WS_EQ_ELASTIC || s->d.eq == WS_EQ_VISCOELASTIC) && s->d.dim == 3) {
                // arena order of the tiled kernels (ws_kernels_fast.cu): neighbours are fetched by one TMA box.  The
                // viscoelastic solver shares the velocity half-step with the elastic one, so its arrays start the same way
                // and the memory variables / relaxation parameters follow.
                static const int fo[9] = {F_VX, F_VY, F_VZ, F_SXX, F_SXY, F_SYY, F_SYZ, F_SZZ, F_SXZ};
                static const int mo[13] = {M_RIX, M_RIY, M_RIZ, M_PW, M_MU, M_MUXY, M_MUXZ, M_MUYZ, M_TAUP, M_TAUS, M_TSXY, M_TSXZ, M_TSYZ};
                const int nf = 9 + 6 * s->L, nm = s->visco ? 13 : 8;
                s->fldArena.alloc((size_t)s->total * nf);
                s->fldArena.zero(s->stream);
                s->matArena.alloc((size_t)s->total * nm);
                s->matArena.zero(s->stream);
                for (int k = 0; k < F_COUNT; k++) { s->ainfo.fldPos[k] = -1; s->ainfo.fldArena[k] = 0; }
                for (int k = 0; k < M_COUNT; k++) { s->ainfo.matPos[k] = -1; s->ainfo.matArena[k] = 1; }
                for (int k = 0; k < nf; k++) {
                    const int slot = k < 9 ? fo[k] : F_R0 + (k - 9); // memory variables: [mechanism][xx yy zz xy xz yz]
                    s->fld[slot].borrow(s->fldArena.p + (size_t)k * s->total, (size_t)s->total);
                    s->ainfo.fldPos[slot] = (short)k;
                }
                for (int k = 0; k < nm; k++) {
                    s->mat[mo[k]].borrow(s->matArena.p + (size_t)k * s->total, (size_t)s->total);
                    s->ainfo.matPos[mo[k]] = (short)k;
                }
                for (auto &kv : fieldsFor(s->d).f)
                    s->fldSlot[kv.first] = kv.second;
                s->ainfo.base[0] = s->fldArena.p; s->ainfo.base[1] = s->matArena.p;
                s->ainfo.count[0] = nf; s->ainfo.count[1] = nm;
                s->ainfo.stride = s->total;
            } else {
                // one arena: the wavefields and the model parameters the half-steps read, in the order of the operand table
                // (ws_kernels_march.cuh spec), second half-step first: arrays that one TMA box can fetch together are neighbours
                std::vector<int> fo, mo;
                auto addU = [](std::vector<int> &v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
                for (int pass = 1; pass >= 0; pass--) {
                    const wsmarch::Lists S = wsmarch::spec(s->d.eq, s->d.dim, pass);
                    for (int k = 0; k < S.nf; k++) addU(fo, S.f[k]);
                    for (int k = 0; k < S.nt; k++) addU(fo, S.t[k]);
                    for (int k = 0; k < S.nq; k++) addU(fo, S.qf[k]);
                    for (int k = 0; k < S.nm; k++) addU(mo, S.m[k]);
                }
                {
                    const wsmarch::Lists S = wsmarch::spec(s->d.eq, s->d.dim, 1);
                    for (int l = 0; l < s->L; l++)
                        for (int k = 0; k < S.nr; k++) addU(fo, F_R0 + 6 * l + S.r[k]);
                    for (int l = 0; l < s->L; l++)
                        for (int k = 0; k < S.nc; k++) addU(mo, M_CD0 + 3 * l + S.c[k]);
                }
                for (auto &kv : fieldsFor(s->d).f)
                    addU(fo, kv.second);
                const size_t nf = fo.size(), nm = mo.size();
                s->arena.alloc((size_t)s->total * (nf + nm));
                s->arena.zero(s->stream);
                for (int k = 0; k < F_COUNT; k++) { s->ainfo.fldPos[k] = -1; s->ainfo.fldArena[k] = 0; }
                for (int k = 0; k < M_COUNT; k++) { s->ainfo.matPos[k] = -1; s->ainfo.matArena[k] = 0; }
                for (size_t k = 0; k < nf; k++) {
                    s->fld[fo[k]].borrow(s->arena.p + k * (size_t)s->total, (size_t)s->total);
                    s->ainfo.fldPos[fo[k]] = (short)k;
                }
                for (size_t k = 0; k < nm; k++) {
                    s->mat[mo[k]].borrow(s->arena.p + (nf + k) * (size_t)s->total, (size_t)s->total);
                    s->ainfo.matPos[mo[k]] = (short)(nf + k);
                }
                s->ainfo.base[0] = s->arena.p;
                s->ainfo.count[0] = (int)(nf + nm);
                s->ainfo.stride = s->total;
                for (auto &kv : fieldsFor(s->d).f)
                    s->fldSlot[kv.first] = kv.second;
            }
            for (int k = 0; k < PSI_COUNT; k++)
                s->psiAxis[k] = -1;
            for (auto &pa : psiFor(s->d))
                s->psiAxis[pa.first] = pa.second;
            s->tdev.alloc(1);
            s->tdev.zero(s->stream);
            s->flag.alloc(1);
            // ghost exchange lists: the fields differentiated along y by the NEXT half-step
            const bool d3 = s->d.dim == 3;
            switch (s->d.eq) {
            case WS_EQ_ACOUSTIC: s->exchA = {F_VY}; s->exchB = {F_P}; break;
            case WS_EQ_ELASTIC:
            case WS_EQ_VISCOELASTIC:
                s->exchA = d3 ? std::vector<int>{F_VX, F_VY, F_VZ} : std::vector<int>{F_VX, F_VY};
                s->exchB = d3 ? std::vector<int>{F_SXY, F_SYY, F_SYZ} : std::vector<int>{F_SXY, F_SYY};
                break;
            case WS_EQ_SH:
            case WS_EQ_VISCOSH: s->exchA = {F_VZ}; s->exchB = {F_SYZ}; break;
            case WS_EQ_TMEM:
            case WS_EQ_VISCOTMEM: s->exchA = {F_HX}; s->exchB = {F_EZ}; break;
            default:
                s->exchA = d3 ? std::vector<int>{F_HZ, F_HX} : std::vector<int>{F_HZ};
                s->exchB = d3 ? std::vector<int>{F_EX, F_EZ} : std::vector<int>{F_EX};
                break;
            }
            WS_CUDA_CHECK(cudaDeviceSynchronize());
            refreshParams(s);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

int ws_create(const ws_desc *desc, ws_solver **out) { return createImpl(desc, 0, out); }
int ws_create_sparse(const ws_desc *desc, int64_t n_points, ws_solver **out)
{
    if (n_points <= 0) {
        g_lastError = "n_points must be positive";
        return WS_EINVAL;
    }
    return createImpl(desc, n_points, out);
}

void ws_destroy(ws_solver *s)
{
    if (!s)
        return;
    cudaSetDevice(s->d.device);
    cudaDeviceSynchronize();
    delete s;
}

int ws_local_range(const ws_solver *s, int32_t *y0, int32_t *nyl)
{
    return guarded([&] {
        WS_REQUIRE(s && y0 && nyl, WS_EINVAL, "null argument");
        *y0 = s->y0;
        *nyl = s->nyl;
    });
}

int ws_set_material(ws_solver *s, const char *name, const float *host, size_t n)
{
    return guarded([&] {
        WS_REQUIRE(s && name && host, WS_EINVAL, "null argument");
        setDevice(s);
        WS_REQUIRE(n == (size_t)s->nx * s->gny * s->nz, WS_EINVAL, "model vector must hold NX*NY*NZ values");
        const int slot = matSlotOf(name);
        uploadGlobal(s, host, matBuf(s, slot));
        s->matGiven[slot] = true;
        s->prepared = false;
    });
}

int ws_set_material_device(ws_solver *s, const char *name, const float *dev, size_t n_local)
{
    return guarded([&] {
        WS_REQUIRE(s && name && dev, WS_EINVAL, "null argument");
        setDevice(s);
        WS_REQUIRE(n_local == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "device model slab must hold NX*NYlocal*NZ values");
        const int slot = matSlotOf(name);
        float *dst = matBuf(s, slot);
        WS_CUDA_CHECK(cudaDeviceSynchronize()); // the producer may have used another stream
        packPlanes(s, dev, dst, 0, s->nyl);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        if (s->d.nranks > 1) { // ghost planes of raw parameters are needed by the averaging passes
            exchangeHalos(s, {dst}, WS_HALO, s->stream);
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        }
        s->matGiven[slot] = true;
        s->prepared = false;
    });
}

int ws_get_material(ws_solver *s, const char *name, float *host, size_t n)
{
    return guarded([&] {
        WS_REQUIRE(s && name && host, WS_EINVAL, "null argument");
        setDevice(s);
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        const int slot = matSlotOf(name);
        WS_REQUIRE(s->mat[slot].p, WS_ESTATE, std::string("model parameter '") + name + "' is not available");
        downloadLocal(s, s->mat[slot].p, host);
    });
}

int ws_prepare(ws_solver *s)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        setDevice(s);
        invalidateGraph(s);
        if (s->sparse) {
            // the caller supplies the operators and the prepareForModelling products of its irregular grid
            const int nop = s->d.dim == 3 ? 6 : 4;
            static const int need2[4] = {wssparse::SP_XF, wssparse::SP_XB, wssparse::SP_YF, wssparse::SP_YB};
            for (int k = 0; k < 4; k++)
                WS_REQUIRE(s->spCol[need2[k]].p, WS_ESTATE, "operator-given mode: Dxf, Dxb, Dyf and Dyb must be set (ws_set_operator)");
            if (nop == 6)
                WS_REQUIRE(s->spCol[wssparse::SP_ZF].p && s->spCol[wssparse::SP_ZB].p, WS_ESTATE, "operator-given mode: Dzf and Dzb must be set (ws_set_operator)");
            requireMat(s, M_PW, "pWaveModulus");
            requireMat(s, M_RIX, "inverseDensityAverageX");
            requireMat(s, M_RIY, "inverseDensityAverageY");
            if (s->d.dim == 3)
                requireMat(s, M_RIZ, "inverseDensityAverageZ");
            refreshParams(s);
            s->useFast = s->useFastA = s->useTma = s->useMarch = s->useTile = false;
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            s->prepared = true;
            return;
        }
        if (s->seismic)
            prepareSeismic(s);
        else
            prepareEM(s);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        prepareBoundaries(s);
        refreshParams(s);
        s->useFast = s->d.kernel_variant == 0 && wsFastSupported(s->P, s->exact);
        s->useFastA = !s->useFast && s->d.kernel_variant == 0 && wsFastSupported(s->P, s->exact, 0) && !(getenv("WS_NO_FAST_A") && atoi(getenv("WS_NO_FAST_A")));
        if (s->fastMaps) {
            wsFastRelease(s->fastMaps);
            s->fastMaps = nullptr;
        }
        if (s->useFast || s->useFastA)
            s->fastMaps = wsFastPrepare(s->P, s->nyl + 2 * WS_HALO);
        // kernel_variant: 0 = best available (3-D elastic TMA kernels, else TMA marching kernels in 3-D and the tile kernels in 2-D),
        // 4 = 2-D tile kernels,
        // 1 = per-point kernels, 2 = cp.async marching kernels, 3 = TMA marching kernels
#ifndef WS_EMULATE
        if (s->tmaMaps) {
            wsTmaRelease(s->tmaMaps);
            s->tmaMaps = nullptr;
        }
        s->P.tmaMaps = nullptr;
        // (2-D grids do not take the TMA MARCHING kernels by default: their strips stage 512 bytes per array and plane, and the
        // TMA ring then holds fewer resident warps per SM than the cp.async kernels do; measured 30 against 41 Gpt/s on 2-D elastic
        // 4096^2, profiles/r02_tma_sweep.txt.  They run on the 2-D tile kernels below.)
        s->useTma = ((!s->useFast && s->d.kernel_variant == 0 && s->d.dim == 3) || s->d.kernel_variant == 3) && wsTmaSupported(s->P, s->ainfo, s->exact);
        if (s->useTma) {
            s->useFast = false;
            s->tmaMaps = wsTmaPrepare(s->P, s->ainfo, s->nyl + 2 * WS_HALO, s->tmaProg, s->tmaNL);
        }
#endif
#ifndef WS_EMULATE
        if (s->tileMaps) {
            wsTileRelease(s->tileMaps);
            s->tileMaps = nullptr;
        }
        s->P.tileMaps = nullptr;
        s->useTile = !s->useFast && !s->useTma && s->d.dim == 2 && (s->d.kernel_variant == 0 || s->d.kernel_variant == 4) && wsTileSupported(s->P, s->ainfo, s->exact);
        if (s->useTile)
            s->tileMaps = wsTilePrepare(s->P, s->ainfo, s->nyl + 2 * WS_HALO, s->tileProg);
#endif
        s->useMarch = !s->useFast && !s->useTma && !s->useTile && (s->d.kernel_variant == 0 || s->d.kernel_variant >= 2) && wsMarchSupported(s->P, s->exact);
        if (s->useMarch)
            wsMarchPrepare(s->P);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        WS_CUDA_CHECK(cudaGetLastError());
        s->prepared = true;
    });
}

} // extern "C"
template <typename IdxT>
static int setSourcesImpl(ws_solver *s, int32_t n, const int32_t *type, const IdxT *idx1d, const float *signals)
{
    return guarded([&] {
        WS_REQUIRE(s && n >= 0 && (n == 0 || (type && idx1d && signals)), WS_EINVAL, "invalid source arguments");
        setDevice(s);
        invalidateGraph(s);
        std::vector<int> types(type, type + n);
        std::vector<long long> off(n);
        const long long N = (long long)s->nx * s->gny * s->nz;
        std::map<std::pair<int, long long>, int> seen;
        bool unique = true;
        for (int k = 0; k < n; k++) {
            WS_REQUIRE(idx1d[k] >= 0 && idx1d[k] < N, WS_EINVAL, "source coordinate outside the grid");
            WS_REQUIRE(type[k] >= 1 && type[k] <= 4, WS_EINVAL, "unknown source type");
            if (s->d.eq == WS_EQ_SH || s->d.eq == WS_EQ_VISCOSH)
                WS_REQUIRE(type[k] == WS_TYPE_VZ, WS_EINVAL, "Pressure, VX and VY sources can not be implemented in SH modeling");
            if (s->seismic && type[k] == WS_TYPE_VZ)
                WS_REQUIRE(s->fld[F_VZ].p, WS_EINVAL, "no VZ wavefield in this modelling");
            if (!s->seismic) {
                const int slot = type[k] == WS_TYPE_EZ ? F_EZ : (type[k] == WS_TYPE_EX ? F_EX : (type[k] == WS_TYPE_EY ? F_EY : F_HZ));
                WS_REQUIRE(s->fld[slot].p, WS_EINVAL, "source type has no wavefield in this modelling");
            }
            const long long pl = (long long)s->nx * s->nz;
            const int y = (int)(idx1d[k] / pl), r = (int)(idx1d[k] % pl), z = r / s->nx, x = r % s->nx;
            off[k] = (y >= s->y0 && y < s->y0 + s->nyl) ? s->offsetOf(x, y - s->y0, z) : -1;
            if (++seen[{type[k], (long long)idx1d[k]}] > 1)
                unique = false;
        }
        // a P source touches Sxx/Syy/Szz: distinct from V targets, so uniqueness per (type, index) is sufficient
        s->nsrc = n;
        s->srcSequential = !unique;
        s->srcType.upload(types);
        s->srcOff.upload(off);
        std::vector<float> sig(signals, signals + (size_t)n * s->d.nt);
        s->srcSig.upload(sig);
        s->srcStep.alloc(std::max(1, n));
        if (s->pinSrc)
            cudaFreeHost(s->pinSrc);
        WS_CUDA_CHECK(cudaMallocHost(&s->pinSrc, std::max(1, n) * sizeof(float)));
        refreshAcq(s);
    });
}

template <typename IdxT>
static int setReceiversImpl(ws_solver *s, int32_t n, const int32_t *type, const IdxT *idx1d)
{
    return guarded([&] {
        WS_REQUIRE(s && n >= 0 && (n == 0 || (type && idx1d)), WS_EINVAL, "invalid receiver arguments");
        setDevice(s);
        invalidateGraph(s);
        std::vector<int> types(type, type + n);
        std::vector<long long> off(n);
        s->recOwned.assign(n, 0);
        const long long N = (long long)s->nx * s->gny * s->nz;
        for (int k = 0; k < n; k++) {
            WS_REQUIRE(idx1d[k] >= 0 && idx1d[k] < N, WS_EINVAL, "receiver coordinate outside the grid");
            WS_REQUIRE(type[k] >= 1 && type[k] <= 4, WS_EINVAL, "unknown receiver type");
            if (s->d.eq == WS_EQ_SH || s->d.eq == WS_EQ_VISCOSH)
                WS_REQUIRE(type[k] == WS_TYPE_VZ, WS_EINVAL, "Pressure, VX and VY receivers can not be implemented in SH modeling");
            if (s->seismic && type[k] == WS_TYPE_VZ)
                WS_REQUIRE(s->fld[F_VZ].p, WS_EINVAL, "no VZ wavefield in this modelling");
            if (!s->seismic) {
                const int slot = type[k] == WS_TYPE_EZ ? F_EZ : (type[k] == WS_TYPE_EX ? F_EX : (type[k] == WS_TYPE_EY ? F_EY : F_HZ));
                WS_REQUIRE(s->fld[slot].p, WS_EINVAL, "receiver type has no wavefield in this modelling");
            }
            const long long pl = (long long)s->nx * s->nz;
            const int y = (int)(idx1d[k] / pl), r = (int)(idx1d[k] % pl), z = r / s->nx, x = r % s->nx;
            const bool mine = y >= s->y0 && y < s->y0 + s->nyl;
            off[k] = mine ? s->offsetOf(x, y - s->y0, z) : -1;
            s->recOwned[k] = mine ? 1 : 0;
        }
        s->nrec = n;
        s->recType.upload(types);
        s->recOff.upload(off);
        s->seis.alloc(std::max<size_t>(1, (size_t)n * s->d.nt));
        s->seis.zero(s->stream);
        s->recStep.alloc(std::max(1, n));
        s->recStep.zero(s->stream);
        if (s->pinRec)
            cudaFreeHost(s->pinRec);
        WS_CUDA_CHECK(cudaMallocHost(&s->pinRec, std::max(1, n) * sizeof(float)));
        WS_CUDA_CHECK(cudaDeviceSynchronize());
        refreshAcq(s);
    });
}

extern "C" {
int ws_set_sources(ws_solver *s, int32_t n, const int32_t *type, const int32_t *idx1d, const float *signals)
{
    if (s && (long long)s->nx * s->gny * s->nz >= (1LL << 31)) {
        g_lastError = "grid exceeds int32 linear indices (scai::IndexType): use ws_set_sources64";
        return WS_EINVAL;
    }
    return setSourcesImpl<int32_t>(s, n, type, idx1d, signals);
}
int ws_set_sources64(ws_solver *s, int32_t n, const int32_t *type, const int64_t *idx1d, const float *signals)
{
    return setSourcesImpl<int64_t>(s, n, type, idx1d, signals);
}
int ws_set_receivers(ws_solver *s, int32_t n, const int32_t *type, const int32_t *idx1d)
{
    if (s && (long long)s->nx * s->gny * s->nz >= (1LL << 31)) {
        g_lastError = "grid exceeds int32 linear indices (scai::IndexType): use ws_set_receivers64";
        return WS_EINVAL;
    }
    return setReceiversImpl<int32_t>(s, n, type, idx1d);
}
int ws_set_receivers64(ws_solver *s, int32_t n, const int32_t *type, const int64_t *idx1d)
{
    return setReceiversImpl<int64_t>(s, n, type, idx1d);
}

int ws_reset(ws_solver *s)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        setDevice(s);
        // the last halo exchange of the previous shot may still be landing in the ghost planes (its ncclRecv completes when
        // the neighbour sends): the memsets below must come after it
        WS_CUDA_CHECK(cudaStreamSynchronize(s->commStream));
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
#ifndef WS_EMULATE
        // peer-memory exchange: the NEIGHBOURS write this rank's ghost planes, so every rank must have finished its last step
        // before anybody clears (and everybody must have cleared before anybody pushes again: second barrier below)
        if (s->p2p)
            rankBarrier(s);
#endif
        for (int k = 0; k < F_COUNT; k++)
            s->fld[k].zero(s->stream); // Wavefields::resetWavefields (Wavefields3Delastic.cpp:111-122)
        for (int k = 0; k < PSI_COUNT; k++)
            s->psi[k].zero(s->stream); // ForwardSolver::resetCPML (CPML3D.cpp:6-26)
        for (int k = 0; k < wssparse::SPSI_COUNT; k++)
            s->spPsi[k].zero(s->stream);
        s->seis.zero(s->stream);
        s->tdev.zero(s->stream);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
#ifndef WS_EMULATE
        if (s->p2p)
            rankBarrier(s);
#endif
    });
}

int ws_step(ws_solver *s, int32_t t)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        WS_REQUIRE(s->prepared, WS_ESTATE, "ws_prepare must be called before time stepping");
        setDevice(s);
        setTime(s, t);
        enqueueStep(s, nullptr, nullptr, nullptr);
        WS_CUDA_CHECK(cudaGetLastError());
    });
}

int ws_step_host(ws_solver *s, int32_t t, const float *src_samples, float *rec_samples)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        WS_REQUIRE(s->prepared, WS_ESTATE, "ws_prepare must be called before time stepping");
        setDevice(s);
        setTime(s, t);
        const float *srcDev = nullptr;
        if (src_samples && s->nsrc > 0) {
            std::memcpy(s->pinSrc, src_samples, s->nsrc * sizeof(float));
            WS_CUDA_CHECK(cudaMemcpyAsync(s->srcStep.p, s->pinSrc, s->nsrc * sizeof(float), cudaMemcpyHostToDevice, s->stream));
            srcDev = s->srcStep.p;
        }
        enqueueStep(s, srcDev, s->recStep.p, nullptr);
        if (rec_samples && s->nrec > 0) {
            WS_CUDA_CHECK(cudaMemcpyAsync(s->pinRec, s->recStep.p, s->nrec * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            std::memcpy(rec_samples, s->pinRec, s->nrec * sizeof(float));
        } else
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        WS_CUDA_CHECK(cudaGetLastError());
    });
}

int ws_set_timing(ws_solver *s, int enable)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        s->timing = enable != 0;
    });
}

int ws_run(ws_solver *s, int32_t t0, int32_t t1)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        WS_REQUIRE(s->prepared, WS_ESTATE, "ws_prepare must be called before time stepping");
        WS_REQUIRE(t0 >= 0 && t1 <= s->d.nt && t0 <= t1, WS_EINVAL, "time range out of bounds");
        setDevice(s);
        if (t0 == t1)
            return;
        setTime(s, t0);
        const int nsteps = t1 - t0;
        if (s->timing) {
            // direct launches with CUDA events around both half-step kernels of every step
            const size_t need = (size_t)4 * nsteps + 2;
            while (s->evPool.size() < need) {
                cudaEvent_t e;
                WS_CUDA_CHECK(cudaEventCreate(&e));
                s->evPool.push_back(e);
            }
            WS_CUDA_CHECK(cudaEventRecord(s->evPool[need - 2], s->stream));
            for (int k = 0; k < nsteps; k++)
                enqueueStep(s, nullptr, nullptr, &s->evPool[(size_t)4 * k]);
            WS_CUDA_CHECK(cudaEventRecord(s->evPool[need - 1], s->stream));
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            double a = 0, b = 0;
            for (int k = 0; k < nsteps; k++) {
                float ms = 0;
                WS_CUDA_CHECK(cudaEventElapsedTime(&ms, s->evPool[4 * k], s->evPool[4 * k + 1]));
                a += ms;
                WS_CUDA_CHECK(cudaEventElapsedTime(&ms, s->evPool[4 * k + 2], s->evPool[4 * k + 3]));
                b += ms;
            }
            float tot = 0;
            WS_CUDA_CHECK(cudaEventElapsedTime(&tot, s->evPool[need - 2], s->evPool[need - 1]));
            s->msA = (float)(a / nsteps);
            s->msB = (float)(b / nsteps);
            s->msStep = tot / nsteps;
            return;
        }
        const bool multi = s->d.nranks > 1;
#ifdef WS_EMULATE
        const bool canGraph = false;
#else
        // multi-rank: the NCCL send / recv of the halo exchange are captured too (fork to the communication stream and
        // join at the end of the graph); not with a bring-your-own transport, which is host-synchronous
        static const bool multiGraph = !(getenv("WS_MULTI_GRAPH") && atoi(getenv("WS_MULTI_GRAPH")) == 0);
        const bool canGraph = !multi || (multiGraph && s->ncclComm && !s->extFn);
#endif
        if (!canGraph || nsteps < 4 || (getenv("WS_NO_GRAPH") && atoi(getenv("WS_NO_GRAPH")) != 0)) {
            for (int k = 0; k < nsteps; k++)
                enqueueStep(s, nullptr, nullptr, nullptr);
            WS_CUDA_CHECK(cudaGetLastError());
            return;
        }
        // CUDA graph of G consecutive steps (the time index lives in device memory, so the graph is step-invariant)
        // (steps per graph: the launch of a graph costs a few microseconds on the device, which 8 steps of a 0.25 ms 2-D step do not
        // amortise as well as 8 steps of a 29 ms 3-D step; developer switch WS_GRAPH_STEPS)
        static const int graphStepsEnv = getenv("WS_GRAPH_STEPS") ? atoi(getenv("WS_GRAPH_STEPS")) : 0;
        const int G = graphStepsEnv >= 2 && graphStepsEnv <= 256 ? graphStepsEnv : 8;
        if (!s->graphExec) {
            cudaGraph_t graph;
            if (multi) // nothing of an earlier exchange may be pending when the capture forks the communication stream
                WS_CUDA_CHECK(cudaStreamSynchronize(s->commStream));
            WS_CUDA_CHECK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            const uint64_t before = s->launches;
            s->capturing = true;
            try {
                for (int k = 0; k < G; k++)
                    enqueueStep(s, nullptr, nullptr, nullptr, multi && k == 0);
                if (multi) // join: the last exchange of the graph belongs to it
                    WS_CUDA_CHECK(cudaStreamWaitEvent(s->stream, s->evCommG, 0));
            } catch (...) {
                s->capturing = false;
                cudaStreamEndCapture(s->stream, &graph);
                throw;
            }
            s->capturing = false;
            s->graphLaunchesPerStep = (s->launches - before) / G;
            s->launches = before;
            WS_CUDA_CHECK(cudaStreamEndCapture(s->stream, &graph));
            WS_CUDA_CHECK(cudaGraphInstantiate(&s->graphExec, graph, 0));
            WS_CUDA_CHECK(cudaGraphDestroy(graph));
            s->graphSteps = G;
        }
        int done = 0;
        const uint64_t perStep = s->graphLaunchesPerStep;
        while (nsteps - done >= G) {
            if (multi) // the exchange of the step before the graph (the graph's first step does not wait inside)
                WS_CUDA_CHECK(cudaStreamWaitEvent(s->stream, s->evComm, 0));
            WS_CUDA_CHECK(cudaGraphLaunch(s->graphExec, s->stream));
            s->launches += perStep * G;
            done += G;
        }
        for (; done < nsteps; done++)
            enqueueStep(s, nullptr, nullptr, nullptr);
        WS_CUDA_CHECK(cudaGetLastError());
    });
}

int ws_sync(ws_solver *s)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        setDevice(s);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        WS_CUDA_CHECK(cudaStreamSynchronize(s->commStream));
        WS_CUDA_CHECK(cudaGetLastError());
        if (s->p2p) {
            int err = 0;
            WS_CUDA_CHECK(cudaMemcpy(&err, s->p2pLocal.p + P2P_ERROR, sizeof(int), cudaMemcpyDeviceToHost));
            WS_REQUIRE(err == 0, WS_ECOMM, "halo exchange over peer memory: a neighbour did not deliver its planes within 20 s");
        }
    });
}

int ws_get_seismogram(ws_solver *s, float *host)
{
    return guarded([&] {
        WS_REQUIRE(s && host, WS_EINVAL, "null argument");
        setDevice(s);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        if (s->nrec == 0)
            return;
        std::vector<float> tmp((size_t)s->nrec * s->d.nt);
        WS_CUDA_CHECK(cudaMemcpy(tmp.data(), s->seis.p, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int r = 0; r < s->nrec; r++)
            if (s->recOwned[r])
                std::memcpy(host + (size_t)r * s->d.nt, tmp.data() + (size_t)r * s->d.nt, s->d.nt * sizeof(float));
    });
}

int ws_get_wavefield(ws_solver *s, const char *comp, float *host, size_t n)
{
    return guarded([&] {
        WS_REQUIRE(s && comp && host, WS_EINVAL, "null argument");
        setDevice(s);
        const std::string name(comp);
        if (name == "CURL" || name == "DIV") {
            // snapType 3 (Wavefields3Delastic.cpp:86-91): derived from the particle velocities with the plain operators
            // elastic / viscoelastic, 2-D TMEz / viscoTMEz, 3-D EM / viscoEM (the wavefield classes that implement getCurl / getDiv)
            const int eq = s->d.eq;
            const bool ok = eq == WS_EQ_ELASTIC || eq == WS_EQ_VISCOELASTIC || ((eq == WS_EQ_TMEM || eq == WS_EQ_VISCOTMEM) && s->d.dim == 2) ||
                            ((eq == WS_EQ_EMEM || eq == WS_EQ_VISCOEMEM) && s->d.dim == 3);
            WS_REQUIRE(ok, WS_EINVAL, "There is no curl or div of wavefield in this modelling");
            WS_REQUIRE(s->seismic || (s->mat[M_EPS].p && s->mat[M_MUM].p && s->mat[M_SIG].p), WS_ESTATE, "the EM model parameters must be set before the curl / div snapshot");
            WS_REQUIRE(s->prepared, WS_ESTATE, "ws_prepare must precede the curl / div snapshot");
            WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
            DevBuf<float> tmp;
            tmp.alloc((size_t)s->total);
            tmp.zero(s->stream);
            WsParams P = s->P;
            P.ylo = 0;
            P.yhi = s->nyl;
            if (s->d.nranks > 1) {
                // the y derivatives read ghost planes: those of the last exchange predate the ABS damping and the source
                // injection of the step, so they are refreshed (collective: every rank takes the snapshot)
                int fA[3];
                firstHalfFields(s, fA);
                std::vector<float *> v;
                for (int k = 0; k < 3; k++)
                    if (fA[k] >= 0)
                        v.push_back(s->fld[fA[k]].p);
                WS_CUDA_CHECK(cudaStreamSynchronize(s->commStream));
                exchangeHalos(s, v, s->h, s->stream);
                WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            }
            wsLaunchDivCurl(P, tmp.p, name == "DIV" ? 1 : 0, s->stream);
            s->launches++;
            downloadLocal(s, tmp.p, host);
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
            return;
        }
        auto it = s->fldSlot.find(comp);
        WS_REQUIRE(it != s->fldSlot.end(), WS_EINVAL, std::string("wavefield '") + comp + "' does not exist in this modelling");
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        downloadLocal(s, s->fld[it->second].p, host);
    });
}

int ws_set_wavefield(ws_solver *s, const char *comp, const float *host, size_t n)
{
    return guarded([&] {
        WS_REQUIRE(s && comp && host, WS_EINVAL, "null argument");
        setDevice(s);
        auto it = s->fldSlot.find(comp);
        WS_REQUIRE(it != s->fldSlot.end(), WS_EINVAL, std::string("wavefield '") + comp + "' does not exist in this modelling");
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        ensureScratch(s, n);
        WS_CUDA_CHECK(cudaMemcpyAsync(s->scratch.p, host, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        packPlanes(s, s->scratch.p, s->fld[it->second].p, 0, s->nyl);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        if (s->d.nranks > 1) {
            exchangeHalos(s, {s->fld[it->second].p}, WS_HALO, s->stream);
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        }
    });
}

int ws_is_finite(ws_solver *s, int32_t *flag)
{
    return guarded([&] {
        WS_REQUIRE(s && flag, WS_EINVAL, "null argument");
        setDevice(s);
        s->flag.zero(s->stream);
        dim3 grid, block;
        s->gridFor(0, s->nyl, grid, block);
        for (int k = 0; k < F_COUNT; k++)
            if (s->fld[k].p) {
                WS_LAUNCH(wsprep::kIsFinite, grid, block, 0, s->stream, s->geo(0, s->nyl), s->fld[k].p, s->flag.p);
                s->launches++;
            }
        int bad = 0;
        WS_CUDA_CHECK(cudaMemcpyAsync(&bad, s->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        if (s->nrec > 0) { // SeismogramHandler::isFinite
            std::vector<float> tmp((size_t)s->nrec * s->d.nt);
            WS_CUDA_CHECK(cudaMemcpy(tmp.data(), s->seis.p, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (float v : tmp)
                if (!std::isfinite(v)) {
                    bad = 1;
                    break;
                }
        }
        if (s->d.nranks > 1 && !s->ncclComm) {
            // bring-your-own transport: the flag travels along the chain of slabs, nranks - 1 sweeps make it global
            WS_REQUIRE(s->extFn, WS_ESTATE, "multi-rank solver used before ws_comm_init");
            DevBuf<float> fl;
            fl.alloc(2);
            float mine = bad ? 1.0f : 0.0f;
            const int up = s->d.rank - 1, down = s->d.rank + 1;
            for (int sweep = 0; sweep < s->d.nranks - 1; sweep++) {
                float got[2] = {0.0f, 0.0f};
                WS_CUDA_CHECK(cudaMemcpy(fl.p, &mine, sizeof(float), cudaMemcpyHostToDevice));
                if (up >= 0) {
                    WS_REQUIRE(s->extFn(s->extUser, fl.p, fl.p + 1, 1, up) == 0, WS_ECOMM, "external transport failed");
                    WS_CUDA_CHECK(cudaMemcpy(&got[0], fl.p + 1, sizeof(float), cudaMemcpyDeviceToHost));
                }
                if (down < s->d.nranks) {
                    WS_REQUIRE(s->extFn(s->extUser, fl.p, fl.p + 1, 1, down) == 0, WS_ECOMM, "external transport failed");
                    WS_CUDA_CHECK(cudaMemcpy(&got[1], fl.p + 1, sizeof(float), cudaMemcpyDeviceToHost));
                }
                mine = std::max(mine, std::max(got[0], got[1]));
            }
            bad = mine != 0.0f;
        }
        if (s->d.nranks > 1 && s->ncclComm) { // commShot->all(...)
            WS_CUDA_CHECK(cudaMemcpyAsync(s->flag.p, &bad, sizeof(int), cudaMemcpyHostToDevice, s->stream));
            g_nccl.check(g_nccl.AllReduce(s->flag.p, s->flag.p, 1, kNcclInt, kNcclMax, s->ncclComm, s->stream), "ncclAllReduce");
            WS_CUDA_CHECK(cudaMemcpyAsync(&bad, s->flag.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        }
        *flag = bad ? 0 : 1;
    });
}

// --- operator-given mode ---------------------------------------------------------------------------------------------
namespace {
// row-major host ELL (n x taps) -> column-major device arrays
void uploadEll(long long n, int taps, const int32_t *cols, const float *vals, DevBuf<int> &dc, DevBuf<float> &dv)
{
    std::vector<int> c((size_t)n * taps);
    std::vector<float> v((size_t)n * taps);
    for (long long i = 0; i < n; i++)
        for (int k = 0; k < taps; k++) {
            c[(size_t)k * n + i] = cols[(size_t)i * taps + k];
            v[(size_t)k * n + i] = vals[(size_t)i * taps + k];
        }
    dc.upload(c);
    dv.upload(v);
}
} // namespace

int ws_set_operator(ws_solver *s, const char *name, int32_t max_taps, const int32_t *cols, const float *vals)
{
    return guarded([&] {
        WS_REQUIRE(s && name && cols && vals, WS_EINVAL, "null argument");
        WS_REQUIRE(s->sparse, WS_ESTATE, "ws_set_operator needs a solver created with ws_create_sparse");
        WS_REQUIRE(max_taps >= 1 && max_taps <= 64, WS_EINVAL, "max_taps out of range");
        setDevice(s);
        invalidateGraph(s);
        static const char *names[wssparse::SP_NOPS] = {"Dxf", "Dxb", "Dyf", "Dyb", "Dzf", "Dzb"};
        int op = -1;
        for (int k = 0; k < wssparse::SP_NOPS; k++)
            if (std::strcmp(name, names[k]) == 0)
                op = k;
        WS_REQUIRE(op >= 0, WS_EINVAL, std::string("unknown operator '") + name + "' (Dxf Dxb Dyf Dyb Dzf Dzb)");
        const long long n = s->nx;
        for (long long i = 0; i < n * max_taps; i++)
            WS_REQUIRE(cols[i] >= -1 && cols[i] < n, WS_EINVAL, std::string("operator '") + name + "': column index out of range");
        uploadEll(n, max_taps, cols, vals, s->spCol[op], s->spVal[op]);
        s->spTaps[op] = max_taps;
        s->prepared = false;
    });
}

int ws_set_interpolation(ws_solver *s, const char *name, int64_t n_rows, const int32_t *rows, int32_t max_taps, const int32_t *cols, const float *vals)
{
    return guarded([&] {
        WS_REQUIRE(s && name && (n_rows == 0 || (rows && cols && vals)), WS_EINVAL, "null argument");
        WS_REQUIRE(s->sparse, WS_ESTATE, "ws_set_interpolation needs a solver created with ws_create_sparse");
        WS_REQUIRE(n_rows >= 0 && max_taps >= 1 && max_taps <= 64, WS_EINVAL, "invalid interpolation size");
        setDevice(s);
        invalidateGraph(s);
        static const char *names[3] = {"InterpolationFull", "InterpolationStaggeredX", "InterpolationStaggeredZ"};
        int w = -1;
        for (int k = 0; k < 3; k++)
            if (std::strcmp(name, names[k]) == 0)
                w = k;
        WS_REQUIRE(w >= 0, WS_EINVAL, std::string("unknown interpolation '") + name + "'");
        ws_solver::Interp &I = s->spInterp[w];
        I.nrows = n_rows;
        I.taps = max_taps;
        if (n_rows == 0)
            return;
        for (int64_t r = 0; r < n_rows; r++)
            WS_REQUIRE(rows[r] >= 0 && rows[r] < s->nx, WS_EINVAL, "interpolation row out of range");
        for (int64_t i = 0; i < n_rows * max_taps; i++)
            WS_REQUIRE(cols[i] >= -1 && cols[i] < s->nx, WS_EINVAL, "interpolation column out of range");
        I.rows.upload(std::vector<int>(rows, rows + n_rows));
        uploadEll(n_rows, max_taps, cols, vals, I.cols, I.vals);
        if ((long long)s->spTmp.n < n_rows)
            s->spTmp.alloc((size_t)n_rows);
    });
}

int ws_set_cpml_profile(ws_solver *s, int32_t axis, int64_t n, const int32_t *idx, const float *a, const float *b, const float *a_half, const float *b_half)
{
    return guarded([&] {
        WS_REQUIRE(s && (n == 0 || (idx && a && b && a_half && b_half)), WS_EINVAL, "null argument");
        WS_REQUIRE(s->sparse, WS_ESTATE, "ws_set_cpml_profile needs a solver created with ws_create_sparse");
        WS_REQUIRE(axis >= 0 && axis < 3 && n >= 0, WS_EINVAL, "invalid axis");
        setDevice(s);
        invalidateGraph(s);
        std::vector<int> k((size_t)s->nx, -1);
        for (int64_t e = 0; e < n; e++) {
            WS_REQUIRE(idx[e] >= 0 && idx[e] < s->nx && k[idx[e]] < 0, WS_EINVAL, "CPML profile: index out of range or listed twice");
            k[idx[e]] = (int)e;
        }
        s->spCpK[axis].upload(k);
        s->spCa[axis].upload(std::vector<float>(a, a + n));
        s->spCb[axis].upload(std::vector<float>(b, b + n));
        s->spCah[axis].upload(std::vector<float>(a_half, a_half + n));
        s->spCbh[axis].upload(std::vector<float>(b_half, b_half + n));
        // memory variables of the two terms of this axis: d/d(axis) of p, and of the velocity component along the axis
        for (int slot : {wssparse::SPSI_P_X + axis, wssparse::SPSI_VXX + axis}) {
            s->spPsi[slot].alloc((size_t)std::max<int64_t>(1, n));
            s->spPsi[slot].zero(s->stream);
        }
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    });
}

int ws_set_abs_profile(ws_solver *s, int64_t n, const int32_t *idx, const float *damping)
{
    return guarded([&] {
        WS_REQUIRE(s && (n == 0 || (idx && damping)), WS_EINVAL, "null argument");
        WS_REQUIRE(s->sparse, WS_ESTATE, "ws_set_abs_profile needs a solver created with ws_create_sparse");
        WS_REQUIRE(s->d.damping == 1 || n == 0, WS_ESTATE, "ws_set_abs_profile: the solver was not created with DampingBoundary = 1");
        setDevice(s);
        invalidateGraph(s);
        for (int64_t e = 0; e < n; e++)
            WS_REQUIRE(idx[e] >= 0 && idx[e] < s->nx, WS_EINVAL, "damping index out of range");
        s->spAbsN = n;
        if (n > 0) {
            s->spAbsIdx.upload(std::vector<int>(idx, idx + n));
            s->spAbsVal.upload(std::vector<float>(damping, damping + n));
        }
    });
}

int ws_set_surface(ws_solver *s, int64_t n, const int32_t *idx)
{
    return guarded([&] {
        WS_REQUIRE(s && (n == 0 || idx), WS_EINVAL, "null argument");
        WS_REQUIRE(s->sparse, WS_ESTATE, "ws_set_surface needs a solver created with ws_create_sparse");
        setDevice(s);
        invalidateGraph(s);
        for (int64_t e = 0; e < n; e++)
            WS_REQUIRE(idx[e] >= 0 && idx[e] < s->nx, WS_EINVAL, "surface index out of range");
        s->spSurfN = n;
        if (n > 0)
            s->spSurf.upload(std::vector<int>(idx, idx + n));
    });
}


} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// wavefield objects and their operators (Wavefields/Wavefields.hpp:62-80)
// ---------------------------------------------------------------------------------------------------------------------
struct ws_wavefields {
    ws_solver *owner = nullptr;
    DevBuf<float> f[F_COUNT]; // the slots the owner has, each `total` floats in the owner's layout
};

namespace {
// this rank's slab of a grid vector (dense, nyl*nz*nx values) -> a zero-padded array in the layout of the wavefields
void uploadLocalPadded(ws_solver *s, const float *host, size_t n, DevBuf<float> &out)
{
    out.alloc((size_t)s->total);
    out.zero(s->stream);
    if (s->sparse) {
        WS_CUDA_CHECK(cudaMemcpyAsync(out.p, host, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        return;
    }
    ensureScratch(s, n);
    WS_CUDA_CHECK(cudaMemcpyAsync(s->scratch.p, host, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    packPlanes(s, s->scratch.p, out.p, 0, s->nyl);
}
float *wfSlot(ws_solver *s, ws_wavefields *w, int k) { return w ? w->f[k].p : s->fld[k].p; }
const float *wfSlot(ws_solver *s, const ws_wavefields *w, int k) { return w ? w->f[k].p : s->fld[k].p; }
void wfCheck(ws_solver *s, const ws_wavefields *a, const ws_wavefields *b)
{
    WS_REQUIRE(s, WS_EINVAL, "null argument");
    WS_REQUIRE((!a || a->owner == s) && (!b || b->owner == s), WS_EINVAL, "the wavefield object belongs to another solver");
}
int wfBinary(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src, int op)
{
    return guarded([&] {
        wfCheck(s, dst, src);
        WS_REQUIRE(dst != src, WS_EINVAL, "source and destination are the same wavefield object");
        setDevice(s);
        const size_t n = (size_t)s->total;
        for (int k = 0; k < F_COUNT; k++)
            if (s->fld[k].p) {
                WS_LAUNCH(kWfBinary, (unsigned)((n + 255) / 256), 256, 0, s->stream, wfSlot(s, dst, k), wfSlot(s, src, k), n, op);
                s->launches++;
            }
        WS_CUDA_CHECK(cudaGetLastError());
    });
}
} // namespace

extern "C" {

int ws_wavefields_create(ws_solver *s, ws_wavefields **out)
{
    return guarded([&] {
        WS_REQUIRE(s && out, WS_EINVAL, "null argument");
        setDevice(s);
        std::unique_ptr<ws_wavefields> w(new ws_wavefields);
        w->owner = s;
        for (int k = 0; k < F_COUNT; k++)
            if (s->fld[k].p) {
                w->f[k].alloc((size_t)s->total);
                w->f[k].zero(s->stream);
            }
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        *out = w.release();
    });
}
void ws_wavefields_destroy(ws_wavefields *w)
{
    if (!w)
        return;
    if (w->owner) {
        setDevice(w->owner);
        cudaStreamSynchronize(w->owner->stream);
    }
    delete w;
}
int ws_wavefields_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src) { return wfBinary(s, dst, src, 0); }
int ws_wavefields_plus_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src) { return wfBinary(s, dst, src, 1); }
int ws_wavefields_minus_assign(ws_solver *s, ws_wavefields *dst, const ws_wavefields *src) { return wfBinary(s, dst, src, 2); }
int ws_wavefields_times_assign(ws_solver *s, ws_wavefields *dst, float rhs)
{
    return guarded([&] {
        wfCheck(s, dst, nullptr);
        setDevice(s);
        const size_t n = (size_t)s->total;
        for (int k = 0; k < F_COUNT; k++)
            if (s->fld[k].p) {
                WS_LAUNCH(kWfScale, (unsigned)((n + 255) / 256), 256, 0, s->stream, wfSlot(s, dst, k), n, rhs);
                s->launches++;
            }
        WS_CUDA_CHECK(cudaGetLastError());
    });
}
int ws_wavefields_times_assign_vector(ws_solver *s, ws_wavefields *dst, const float *host, size_t n)
{
    return guarded([&] {
        wfCheck(s, dst, nullptr);
        WS_REQUIRE(host, WS_EINVAL, "null argument");
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        setDevice(s);
        DevBuf<float> vec;
        uploadLocalPadded(s, host, n, vec);
        // own planes only: the pads of the vector are zero, and the ghost planes of live wavefields belong to the neighbours
        const size_t first = s->sparse ? 0 : (size_t)WS_HALO * (size_t)s->plane, cnt = s->sparse ? (size_t)s->total : (size_t)s->nyl * (size_t)s->plane;
        for (int k = 0; k < F_COUNT; k++)
            if (s->fld[k].p) {
                WS_LAUNCH(kWfScaleVec, (unsigned)((cnt + 255) / 256), 256, 0, s->stream, wfSlot(s, dst, k) + first, vec.p + first, cnt);
                s->launches++;
            }
        WS_CUDA_CHECK(cudaGetLastError());
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream)); // `vec` is released on return
    });
}
int ws_wavefields_get(ws_solver *s, const ws_wavefields *w, const char *comp, float *host, size_t n)
{
    if (!w)
        return ws_get_wavefield(s, comp, host, n);
    return guarded([&] {
        wfCheck(s, w, nullptr);
        WS_REQUIRE(comp && host, WS_EINVAL, "null argument");
        setDevice(s);
        auto it = s->fldSlot.find(comp);
        WS_REQUIRE(it != s->fldSlot.end(), WS_EINVAL, std::string("wavefield '") + comp + "' does not exist in this modelling");
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        downloadLocal(s, w->f[it->second].p, host);
    });
}
int ws_set_step_scaling(ws_solver *s, const float *host, size_t n)
{
    return guarded([&] {
        WS_REQUIRE(s, WS_EINVAL, "null argument");
        setDevice(s);
        invalidateGraph(s);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
        if (!host) {
            s->stepScale.release();
            return;
        }
        WS_REQUIRE(n == (size_t)s->nx * s->nyl * s->nz, WS_EINVAL, "size mismatch");
        uploadLocalPadded(s, host, n, s->stepScale);
        WS_CUDA_CHECK(cudaStreamSynchronize(s->stream));
    });
}

} // extern "C"


extern "C" {

int ws_comm_unique_id(void *id128)
{
    return guarded([&] {
        WS_REQUIRE(id128, WS_EINVAL, "null argument");
        g_nccl.load();
        NcclApi::UniqueId id;
        g_nccl.check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id128, &id, 128);
    });
}

int ws_comm_init(ws_solver *s, const void *id128)
{
    return guarded([&] {
        WS_REQUIRE(s && id128, WS_EINVAL, "null argument");
        setDevice(s);
        g_nccl.load();
        NcclApi::UniqueId id;
        std::memcpy(&id, id128, 128);
        g_nccl.check(g_nccl.CommInitRank(&s->ncclComm, s->d.nranks, id, s->d.rank), "ncclCommInitRank");
#ifndef WS_EMULATE
        p2pSetup(s); // halo planes over peer memory where every neighbour can be mapped (NCCL stays for the collectives)
#endif
    });
}

int ws_halo_transport(const ws_solver *s)
{
    if (!s || s->d.nranks <= 1)
        return 0;
    return s->p2p ? 3 : (s->extFn ? 2 : (s->ncclComm ? 1 : 0));
}

int ws_comm_init_external(ws_solver *s, ws_sendrecv_fn fn, void *user)
{
    return guarded([&] {
        WS_REQUIRE(s && fn, WS_EINVAL, "null argument");
        s->extFn = fn;
        s->extUser = user;
    });
}

uint64_t ws_launch_count(const ws_solver *s) { return s ? s->launches : 0; }

int ws_last_timing(ws_solver *s, int which, float *ms)
{
    return guarded([&] {
        WS_REQUIRE(s && ms, WS_EINVAL, "null argument");
        *ms = which == 0 ? s->msA : (which == 1 ? s->msB : s->msStep);
    });
}

void *ws_stream(ws_solver *s) { return s ? (void *)s->stream : nullptr; }

int ws_uses_fast_kernels(const ws_solver *s) { return s && s->useFast ? 1 : 0; }
int ws_kernel_path(const ws_solver *s) { return !s ? -1 : (s->useFast ? 2 : (s->useTma ? 3 : (s->useTile ? 4 : (s->useMarch ? 1 : 0)))); }

} // extern "C"

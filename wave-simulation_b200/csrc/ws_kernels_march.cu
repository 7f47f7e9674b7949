// ws_kernels_march.cu — instantiation + launch dispatch of the marching kernels (ws_kernels_march.cuh).
// Built once per FD order (-DWS_MARCH_Q=<q>: the kernels of that order) and once without (the dispatcher), so that the
// orders compile in parallel; the host emulation build of the test suite (-DWS_MARCH_ALL) takes everything in one unit.
#include "../../include/wavesim.h"
#include "ws_kernels_march.cuh"
#include "ws_launch.hpp"

#include <cstdlib>

#ifdef WS_EMULATE
#define WS_LAUNCH_COOP(kern, grid, block, smem, stream, ...) wsemu::launchCoop(kern, dim3(grid), dim3(block), smem, __VA_ARGS__)
#else
#define WS_LAUNCH_COOP(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

namespace {

template <int EQ, int DIM, int Q, int PASS, int NST, int NL> void launchK(const WsParams &P, size_t smem, cudaStream_t st)
{
    using G = wsmarch::Geo<DIM, Q, NL>;
    auto k = wsmarch::kMarch<EQ, DIM, Q, PASS, NST, NL>;
#ifndef WS_EMULATE
    if (wsOptInSmem(reinterpret_cast<const void *>(k), 227 * 1024 - 1024) != cudaSuccess)
        return; // the error stays pending: ws_step / ws_run report it through cudaGetLastError
#endif
    const int ny = P.yhi - P.ylo;
    const dim3 grid((P.nx + G::TX - 1) / G::TX, (P.nz + G::TZ - 1) / G::TZ, (ny + P.marchChunk - 1) / P.marchChunk);
    const int nthr = G::NTHR;
    WS_LAUNCH_COOP(k, grid, nthr, smem, st, P);
}

template <int EQ, int DIM, int Q, int PASS, int NL> void launchP(const WsParams &P, cudaStream_t st)
{
    const int Ls = P.marchStageR ? P.L : 0;
    const size_t stage = sizeof(float) * wsmarch::stageFloats<EQ, DIM, Q, PASS, NL>(Ls);
    // ring depth 2 (the plane after the one being computed is in flight): more resident thread blocks hide the latencies
    // of the compute phase better than a deeper ring does (measured on all BASELINE configs; 3 is a developer switch)
    int nst = 2;
    if (P.marchStages == 2 || P.marchStages == 3)
        nst = P.marchStages;
#ifndef WS_EMULATE /* the host emulation of the test suite instantiates the default depth only */
    if (nst == 3) {
        launchK<EQ, DIM, Q, PASS, 3, NL>(P, 3 * stage, st);
        return;
    }
#endif
    (void)nst;
    launchK<EQ, DIM, Q, PASS, 2, NL>(P, 2 * stage, st);
}

// one x point per thread where the operands of a plane are so many that 4 points per thread would leave one or two
// thread blocks of 4 warps per SM (the 3-D elastic / viscoelastic half-steps); 4 points per thread otherwise
template <int EQ, int DIM, int Q> void launchT(const WsParams &P, int pass, cudaStream_t st)
{
    // (measured: 3-D viscoelastic 768^3 velocity half-step 7.9 ms with 1 point against 9.1 ms with 4, stress half-step
    // 23.8 against 40.3 ms; 3-D elastic 1024^3 17.7 / 21.0 against 19.8 / 22.1 ms)
    constexpr bool heavy = DIM == 3 && (EQ == WS_EQ_ELASTIC || EQ == WS_EQ_VISCOELASTIC || EQ == WS_EQ_VISCOEMEM);
    if constexpr (heavy) {
        const bool one = P.marchLanes == 1 || (P.marchLanes == 0 && (EQ != WS_EQ_VISCOEMEM || (pass == 1 && P.L > 0 && P.marchStageR)));
        if (one) {
            if (pass == 0)
                launchP<EQ, DIM, Q, 0, 1>(P, st);
            else
                launchP<EQ, DIM, Q, 1, 1>(P, st);
            return;
        }
    }
    if (pass == 0)
        launchP<EQ, DIM, Q, 0, 4>(P, st);
    else
        launchP<EQ, DIM, Q, 1, 4>(P, st);
}

template <int EQ, int Q> void launchD(const WsParams &P, int pass, cudaStream_t st)
{
    if (P.dim == 3)
        launchT<EQ, 3, Q>(P, pass, st);
    else
        launchT<EQ, 2, Q>(P, pass, st);
}

template <int Q> void launchQ(const WsParams &P, int pass, cudaStream_t st)
{
    switch (P.eq) {
    case WS_EQ_ACOUSTIC: launchD<WS_EQ_ACOUSTIC, Q>(P, pass, st); break;
    case WS_EQ_ELASTIC: launchD<WS_EQ_ELASTIC, Q>(P, pass, st); break;
    case WS_EQ_VISCOELASTIC: launchD<WS_EQ_VISCOELASTIC, Q>(P, pass, st); break;
    case WS_EQ_SH: launchT<WS_EQ_SH, 2, Q>(P, pass, st); break;
    case WS_EQ_VISCOSH: launchT<WS_EQ_VISCOSH, 2, Q>(P, pass, st); break;
    case WS_EQ_TMEM: launchT<WS_EQ_TMEM, 2, Q>(P, pass, st); break;
    case WS_EQ_VISCOTMEM: launchT<WS_EQ_VISCOTMEM, 2, Q>(P, pass, st); break;
    case WS_EQ_EMEM: launchD<WS_EQ_EMEM, Q>(P, pass, st); break;
    case WS_EQ_VISCOEMEM: launchD<WS_EQ_VISCOEMEM, Q>(P, pass, st); break;
    default: break;
    }
}

} // namespace

#define WS_MARCH_NAME2(q) wsLaunchMarchQ##q
#define WS_MARCH_NAME(q) WS_MARCH_NAME2(q)

#ifdef WS_MARCH_Q
void WS_MARCH_NAME(WS_MARCH_Q)(const WsParams &P, int pass, cudaStream_t st) { launchQ<WS_MARCH_Q>(P, pass, st); }
#else

#ifdef WS_MARCH_ALL
void wsLaunchMarchQ2(const WsParams &P, int pass, cudaStream_t st) { launchQ<2>(P, pass, st); }
void wsLaunchMarchQ4(const WsParams &P, int pass, cudaStream_t st) { launchQ<4>(P, pass, st); }
void wsLaunchMarchQ6(const WsParams &P, int pass, cudaStream_t st) { launchQ<6>(P, pass, st); }
void wsLaunchMarchQ8(const WsParams &P, int pass, cudaStream_t st) { launchQ<8>(P, pass, st); }
void wsLaunchMarchQ10(const WsParams &P, int pass, cudaStream_t st) { launchQ<10>(P, pass, st); }
void wsLaunchMarchQ12(const WsParams &P, int pass, cudaStream_t st) { launchQ<12>(P, pass, st); }
#else
void wsLaunchMarchQ2(const WsParams &P, int pass, cudaStream_t st);
void wsLaunchMarchQ4(const WsParams &P, int pass, cudaStream_t st);
void wsLaunchMarchQ6(const WsParams &P, int pass, cudaStream_t st);
void wsLaunchMarchQ8(const WsParams &P, int pass, cudaStream_t st);
void wsLaunchMarchQ10(const WsParams &P, int pass, cudaStream_t st);
void wsLaunchMarchQ12(const WsParams &P, int pass, cudaStream_t st);
#endif

// every equation type and order; FMA arithmetic only (the exact-arithmetic parity mode stays on the per-point kernels)
bool wsMarchSupported(const WsParams &P, bool exact)
{
    if (exact)
        return false;
    if (P.q < 2 || P.q > 12 || (P.q & 1))
        return false;
    return P.dim == 2 || P.dim == 3;
}

// planes per thread block.  Short chunks win (measured, profiles/r01_march_chunk_sweep.txt): the thread blocks of one
// y range march in step, so the halo rows they share are still in L2 when the neighbour asks for them (with one chunk
// per column the marches drift apart and the halos come from HBM again: 3-D acoustic 1024^3 58.7 -> 77.1 Gpt/s at 64
// planes), and small 2-D grids only fill the 148 SMs when they are cut into ~3000 thread blocks.  The price, q feed-only
// planes at the start of every chunk, bounds the chunk from below.
void wsMarchPrepare(WsParams &P)
{
    const int TX = P.dim == 3 ? 64 : WS_MARCH_TX2D, TZ = P.dim == 3 ? 8 : 1;
    P.marchLanes = getenv("WS_MARCH_LANES") ? atoi(getenv("WS_MARCH_LANES")) : 0;
    const long long tiles = (long long)((P.nx + TX - 1) / TX) * ((P.nz + TZ - 1) / TZ);
    const long long want = 148LL * 20;
    long long nchunks = (want + tiles - 1) / tiles;
    if (nchunks < 1)
        nchunks = 1;
    int chunk = (int)((P.nyl + nchunks - 1) / nchunks);
    chunk = chunk < 16 ? 16 : (chunk > 64 ? 64 : chunk);
    if (const char *e = getenv("WS_MARCH_CHUNK"))
        chunk = atoi(e) > 0 ? atoi(e) : chunk;
    P.marchChunk = chunk;
    P.marchStages = getenv("WS_MARCH_STAGES") ? atoi(getenv("WS_MARCH_STAGES")) : 0;
    P.marchDebug = getenv("WS_MARCH_DEBUG") ? atoi(getenv("WS_MARCH_DEBUG")) : 0;
    P.marchStageR = getenv("WS_MARCH_STAGE_R") ? atoi(getenv("WS_MARCH_STAGE_R")) : 1;
}

int wsLaunchMarch(const WsParams &P, int pass, cudaStream_t st)
{
    if (P.yhi <= P.ylo || P.marchChunk <= 0)
        return 0;
    switch (P.q) {
    case 2: wsLaunchMarchQ2(P, pass, st); break;
    case 4: wsLaunchMarchQ4(P, pass, st); break;
    case 6: wsLaunchMarchQ6(P, pass, st); break;
    case 8: wsLaunchMarchQ8(P, pass, st); break;
    case 10: wsLaunchMarchQ10(P, pass, st); break;
    case 12: wsLaunchMarchQ12(P, pass, st); break;
    default: return 0;
    }
    return 1;
}
#endif

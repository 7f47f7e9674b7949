// ws_launch.hpp — launch entry points shared between translation units
#pragma once
#include "ws_common.cuh"

// general kernels (ws_kernels_general.cu)
void wsGeneralGrid(const WsParams &P, dim3 &grid, dim3 &block);
void wsLaunchGeneral(const WsParams &P, bool exact, int pass, cudaStream_t st);
void wsLaunchAbsFirstHalf(const WsParams &P, bool exact, int f0, int f1, int f2, cudaStream_t st);
// snapType 3: curl (which = 0) / div (which = 1) energy measure of the velocity field into the padded array `out`
void wsLaunchDivCurl(const WsParams &P, float *out, int which, cudaStream_t st);

// tiled fast kernels (ws_kernels_fast.cu); return false when the configuration is not covered
bool wsFastSupported(const WsParams &P, bool exact, int pass = -1);
// builds the TMA tensor maps for this solver's arrays into a device buffer (returned, owned by the caller; freed with
// wsFastRelease) and fills P.fastMaps / P.fastChunk
void *wsFastPrepare(WsParams &P, int nyp);
void wsFastRelease(void *maps);
int wsLaunchFast(const WsParams &P, int pass, cudaStream_t st); // number of kernels launched (0 = not served)

// marching kernels (ws_kernels_march.cu): every equation type, compile-time FD order, FMA arithmetic
bool wsMarchSupported(const WsParams &P, bool exact);
void wsMarchPrepare(WsParams &P); // fills P.marchChunk
int wsLaunchMarch(const WsParams &P, int pass, cudaStream_t st); // number of kernels launched (0 = not served)

// TMA marching kernels (ws_kernels_tma.cu): every equation type, compile-time FD order, FMA arithmetic; the operands are
// fetched by tensor maps over the solver's arenas (allocations that hold several padded arrays at a constant stride)
struct WsArenaInfo {
    const float *base[2] = {nullptr, nullptr}; // arena origins
    int count[2] = {0, 0};                     // arrays per arena
    long long stride = 0;                      // floats between two arrays of an arena
    signed char fldArena[F_COUNT], matArena[M_COUNT]; // arena of a wavefield / model slot
    short fldPos[F_COUNT], matPos[M_COUNT];           // position inside it, -1 = not in an arena
};
namespace wstma { struct TmaProg; }
bool wsTmaSupported(const WsParams &P, const WsArenaInfo &A, bool exact);
void *wsTmaPrepare(WsParams &P, const WsArenaInfo &A, int nyp, wstma::TmaProg prog[2], int nl[2]);
void wsTmaRelease(void *maps);
int wsLaunchTma(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);

// 2-D tile kernels (ws_kernels_tile2d.cu): every 2-D equation type, compile-time FD order, FMA arithmetic; one thread block per
// 128 x 8 / 16 tile, all operands of the tile fetched by a handful of TMA boxes
namespace wstile { struct TileProg; }
bool wsTileSupported(const WsParams &P, const WsArenaInfo &A, bool exact);
void *wsTilePrepare(WsParams &P, const WsArenaInfo &A, int nyp, wstile::TileProg prog[2]);
void wsTileRelease(void *maps);
int wsLaunchTile(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);

#ifndef WS_EMULATE
#include <mutex>
#include <set>
#include <utility>
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, device): one process may drive several GPUs
// (host/Simulation.cpp runs one thread per shot domain), so the opt-in is remembered per device, under a lock.
inline cudaError_t wsOptInSmem(const void *kernel, int bytes)
{
    static std::mutex m;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    std::lock_guard<std::mutex> lock(m);
    if (done.count({kernel, dev}))
        return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess)
        done.insert({kernel, dev});
    return e;
}
#endif

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
bool wsFastSupported(const WsParams &P, bool exact);
// builds the TMA tensor maps for this solver's arrays into a device buffer (returned, owned by the caller; freed with
// wsFastRelease) and fills P.fastMaps / P.fastChunk
void *wsFastPrepare(WsParams &P, int nyp);
void wsFastRelease(void *maps);
int wsLaunchFast(const WsParams &P, int pass, cudaStream_t st); // number of kernels launched (0 = not served)

// marching kernels (ws_kernels_march.cu): every equation type, compile-time FD order, FMA arithmetic
bool wsMarchSupported(const WsParams &P, bool exact);
void wsMarchPrepare(WsParams &P); // fills P.marchChunk
int wsLaunchMarch(const WsParams &P, int pass, cudaStream_t st); // number of kernels launched (0 = not served)

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

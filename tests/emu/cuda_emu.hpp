// cuda_emu.hpp — TEST INFRASTRUCTURE ONLY.
//
// Minimal host-side stand-in for the CUDA runtime API and the device intrinsics used by wave-simulation_b200/csrc,
// so that the CPU-only test suite (`pytest -m "not gpu"`, no GPU in the build container) can compile the general
// kernels, the model-preparation kernels and the whole C ABI with -DWS_EMULATE into tests/emu/libwavesim_emu.so and
// check their LOGIC against the oracle before GPU time is spent.  Kernels without shared memory run as sequential loops
// over the launch grid (launch); kernels that use shared memory and __syncthreads run one thread block at a time with
// one OS thread per CUDA thread and a barrier (launchCoop).  This library is never loaded by the product, by
// bench.py or by the `-m gpu` tests: the product has no CPU path and fails loudly without its CUDA library.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 {
    unsigned x, y, z;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
namespace wsemu {
inline thread_local uint3 g_blockIdx, g_threadIdx;
inline thread_local dim3 g_blockDim, g_gridDim;
template <typename K, typename... Args>
void launch(K kern, dim3 grid, dim3 block, Args... args)
{
    g_gridDim = grid;
    g_blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = {bx, by, bz};
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            g_threadIdx = {tx, ty, tz};
                            kern(args...);
                        }
            }
}
// cooperative kernels: the threads of a block are real threads, __syncthreads is a barrier, shared memory is g_smem
inline void *g_smem = nullptr;
inline std::barrier<> *g_barrier = nullptr;
template <typename K, typename... Args>
void launchCoop(K kern, dim3 grid, dim3 block, size_t smemBytes, Args... args)
{
    const unsigned nthr = block.x * block.y * block.z;
    std::vector<unsigned char> smem(smemBytes + 16);
    std::barrier<> bar(nthr), blockBar(nthr);
    g_smem = smem.data();
    g_barrier = &bar;
    // one OS thread per CUDA thread, reused for every block of the grid (blocks run one after the other)
    std::vector<std::thread> pool;
    pool.reserve(nthr);
    for (unsigned t = 0; t < nthr; t++)
        pool.emplace_back([=, &blockBar]() {
            g_gridDim = grid;
            g_blockDim = block;
            g_threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        g_blockIdx = {bx, by, bz};
                        kern(args...);
                        blockBar.arrive_and_wait(); // shared memory is reused by the next block
                    }
        });
    for (auto &th : pool)
        th.join();
    g_smem = nullptr;
    g_barrier = nullptr;
}
} // namespace wsemu
inline void __syncthreads() { wsemu::g_barrier->arrive_and_wait(); }
#define blockIdx wsemu::g_blockIdx
#define threadIdx wsemu::g_threadIdx
#define blockDim wsemu::g_blockDim
#define gridDim wsemu::g_gridDim

// device intrinsics (compile with -ffp-contract=off so that the *_rn forms are single roundings)
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline int atomicExch(int *p, int v)
{
    int o = *p;
    *p = v;
    return o;
}
using std::isinf;
using std::isnan;
using std::max;
using std::min;

// runtime API subset on host memory
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraph_t;
typedef void *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1 };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)std::malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }

// ws_fast_common.cuh — building blocks shared by the tiled TMA kernels (ws_kernels_fast.cu: 3-D elastic,
// ws_kernels_fast_acoustic.cu: 3-D acoustic): tile geometry, mbarrier / TMA wrappers, 128-bit shared-memory stencil
// helpers, the CPML update on staged memory variables, consumer bookkeeping, the developer trace and the host-side
// tensor-map builders.  Everything lives in an unnamed namespace (one copy per translation unit).
#pragma once
#include "../../include/wavesim.h"
#include "ws_launch.hpp"

#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

// planes per trip of the unrolled march: the y queue holds Q + UNR - 1 planes and is shifted by UNR once per trip.  Round 1 (128
// registers): 1 was fastest for both half-steps (full rotation, UNR = Q, needs no moves but its code overflows the instruction cache).
// With the register trade (152 registers) the stress half-step gains from 4 planes per trip (15.44 -> 15.23 ms at 1024^3), the
// velocity half-step still loses (11.36 -> 11.66 ms): one constant per half-step (profiles/r02_northstar_notes.txt).
#ifndef WS_UNRV
#define WS_UNRV 1
#endif
#ifndef WS_UNRS
#define WS_UNRS 4
#endif
constexpr int TX = 64, TZ = 8;
constexpr int WS_TRACE_MAX = 1 << 16; // thread blocks per half-step covered by the developer trace
constexpr int NSTMAX = 4; // deepest stage ring of any tiled kernel
constexpr int NGROUPS = 3;
constexpr int UNRV = WS_UNRV, UNRS = WS_UNRS, UNRMAX = UNRV > UNRS ? UNRV : UNRS;

// Register budget of the warp-specialised kernels.  The register file is handed out per SM sub-partition (16384 registers each):
// with 12 consumer warps + 1 producer warp one partition holds 4 warps, which caps EVERY thread at 128 registers — the consumers sit
// at that cap (the CPML-layer variant of the stress half-step spills).  With a whole producer warpgroup (4 warps, three of them idle)
// the warpgroups can trade registers (setmaxnreg): the producer side keeps 40, the three consumer warpgroups take 152 each
// (384 x 152 + 128 x 40 = 63488 <= 65536).  WS_FAST_MAXNREG=0 builds the 13-warp form of round 1.
#ifndef WS_FAST_MAXNREG
#define WS_FAST_MAXNREG 1
#endif
#if WS_FAST_MAXNREG
constexpr int WS_FAST_PRODUCER_THREADS = 128;
__device__ __forceinline__ void wsConsumerRegs() { asm volatile("setmaxnreg.inc.sync.aligned.u32 152;" ::: "memory"); }
__device__ __forceinline__ void wsProducerRegs() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory"); }
#else
constexpr int WS_FAST_PRODUCER_THREADS = 32;
__device__ __forceinline__ void wsConsumerRegs() {}
__device__ __forceinline__ void wsProducerRegs() {}
#endif

template <int Q> struct Cfg {
    static constexpr int H = Q / 2;
    static constexpr int HX = (H <= 4) ? 4 : 8; // x halo rounded to a float4
    static constexpr int TXH = TX + 2 * HX;
    static constexpr int TZH = TZ + 2 * H;
    static constexpr int LXN = TX / 4;
    static constexpr int NTG = LXN * TZ; // threads per consumer group
    static constexpr int WPG = NTG / 32; // warps per group
    // tile sizes (floats); all are multiples of 32 floats = 128 bytes
    static constexpr int N_P = TX * TZ, N_X = TXH * TZ, N_Z = TX * TZH, N_XZ = TXH * TZH;
    static_assert(NTG % 32 == 0, "consumer groups must be whole warps");
    static_assert(N_P % 32 == 0 && N_X % 32 == 0 && N_Z % 32 == 0 && N_XZ % 32 == 0, "TMA destinations must stay 128-byte aligned");
    static constexpr int QL = Q + UNRMAX - 1; // physical length of the y queue (the half-step with the shorter trips leaves the tail unused)
    static_assert(Q % UNRV == 0 && NSTMAX % UNRV == 0 && Q % UNRS == 0 && NSTMAX % UNRS == 0, "trips must tile the queue prologue and the stage ring");
};

__device__ __forceinline__ uint32_t smemU32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmaLoad4D(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// same with an L2 eviction-priority hint: operands that no thread block reads again before the grid has been swept
// (own-point wavefields and model parameters) are fetched evict-first, so that L2 keeps the halo rows shared with the
// neighbouring tiles and the planes that re-enter as stencil tiles
__device__ __forceinline__ void tmaLoad4DHint(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
                 "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t policyEvictFirst()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policyEvictLast()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 ld4(const float *p)
{
    const float4 t = *reinterpret_cast<const float4 *>(p);
    F4 r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ F4 ldg4(const float *p)
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    F4 r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ void st4(float *p, const F4 &a) { *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
// streaming store: the written plane is not read again before the whole grid has been swept
__device__ __forceinline__ void st4cs(float *p, const F4 &a) { __stcs(reinterpret_cast<float4 *>(p), make_float4(a.v[0], a.v[1], a.v[2], a.v[3])); }
__device__ __forceinline__ F4 zero4()
{
    F4 r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f;
    return r;
}

using A = Ar<false>;

// x derivative of 4 consecutive points from a shared-memory row; `row` points at the tile column of x0 - HX
template <int Q, bool FWD> __device__ __forceinline__ F4 dX(const float *row, const float *__restrict__ c)
{
    constexpr int H = Q / 2, HX = Cfg<Q>::HX, NV = (2 * HX + 4) / 4;
    float w[NV * 4];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const F4 t = ld4(row + 4 * k);
        w[4 * k] = t.v[0]; w[4 * k + 1] = t.v[1]; w[4 * k + 2] = t.v[2]; w[4 * k + 3] = t.v[3];
    }
    F4 r;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < Q; j++)
            acc = A::madd(c[j], w[HX + p + (FWD ? j - H + 1 : j - H)], acc);
        r.v[p] = acc;
    }
    return r;
}
// z derivative: `col` points at (row of z - H, column of x0) of a tile with z halo; rows are LD floats apart
template <int Q, bool FWD, int LD> __device__ __forceinline__ F4 dZ(const float *col, const float *__restrict__ c)
{
    F4 r = zero4();
#pragma unroll
    for (int j = 0; j < Q; j++) {
        const F4 t = ld4(col + (FWD ? j + 1 : j) * LD);
#pragma unroll
        for (int p = 0; p < 4; p++)
            r.v[p] = A::madd(c[j], t.v[p], r.v[p]);
    }
    return r;
}
// y derivative from the register queue; tap k of the R-th plane of a trip lives in q[k + R]
template <int Q, int R> __device__ __forceinline__ F4 dY(const F4 (&q)[Cfg<Q>::QL], const float *__restrict__ w)
{
    F4 r = zero4();
#pragma unroll
    for (int j = 0; j < Q; j++)
#pragma unroll
        for (int p = 0; p < 4; p++)
            r.v[p] = A::madd(w[j], q[j + R].v[p], r.v[p]);
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// CPML (CPML.cpp:84-95 applyCPML): psi = b psi + a d ; d = d + psi.  The memory variables live in compact slabs
// (x: [ly][z][PX], y: [2W][z][x], z: [ly][2W][x]).
// Row layout of the x slabs (wsPsiXIndex, ws_common.cuh): the low-side entries sit at k' = x, the high-side entries at
// k' = x - D with D a multiple of 4, and the entries between the two sides are padding.  A thread's 4 points therefore
// map to ONE aligned float4 of the row, whose entries are either its own layer points or padding nobody else touches:
// the x term is a branch-free vector update with coefficient rows that are zero on the padding (a = b = 0 leaves the
// derivative unchanged and stores psi = 0).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PXMAX = 80; // floats per x-slab row the coefficient table in shared memory can hold (W <= 36)
struct CpT { // per thread, constant over the march
    int kxv;  // k' of the thread's first point, or -1 if none of its 4 points lies in an x layer
    int oXA;  // offset of this role's a' row in the shared coefficient table (b' row follows PX later)
    int kz;
    float za, zb;
};
template <bool CPML> __device__ __forceinline__ void cpSetup(const WsParams &P, CpT &t, bool active, int x0, int z, bool halfX, bool halfZ)
{
    t.kxv = -1;
    t.oXA = 0;
    t.kz = -1;
    t.za = t.zb = 0.0f;
    if (!CPML || !active)
        return;
    const int W = P.W;
    if (x0 < W || x0 + 3 >= P.nx - W)
        t.kxv = x0 < W ? x0 : x0 - P.psiDX;
    t.oXA = (halfX ? 2 : 0) * P.psiPitchX;
    t.kz = wsCpmlIndex(z, P.nz, W);
    if (t.kz >= 0) {
        t.za = __ldg((halfZ ? P.cazh : P.caz) + t.kz);
        t.zb = __ldg((halfZ ? P.cbzh : P.cbz) + t.kz);
    }
}
// one x term: sps = staged slab row of this thread's (ly, z) in shared memory, gps = the same row in the slab, tab = shared
// coefficient table {a', b', a'half, b'half}[PX]
__device__ __forceinline__ void cpApplyXv(const CpT &t, const float *sps, float *gps, const float *tab, int PX, F4 &d)
{
    const F4 old = ld4(sps + t.kxv), a = ld4(tab + t.oXA + t.kxv), b = ld4(tab + t.oXA + PX + t.kxv);
    F4 nw;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float v = A::mul(old.v[p], b.v[p]);
        v = A::add(v, A::mul(a.v[p], d.v[p]));
        nw.v[p] = v;
        d.v[p] = A::add(d.v[p], v);
    }
    st4(gps + t.kxv, nw);
}
__device__ __forceinline__ void cpApply4(float *ps, const F4 &old, float a, float b, F4 &d)
{
    F4 nw;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float v = A::mul(old.v[p], b);
        v = A::add(v, A::mul(a, d.v[p]));
        nw.v[p] = v;
        d.v[p] = A::add(d.v[p], v);
    }
    st4(ps, nw);
}

// per-plane (run-time) y quantities of the generic step
template <int Q> struct YDyn {
    float w[Q]; // y weights of this plane
    int ky;     // y-CPML slab index or -1
    float ya, yb;
};
// y weights of the first half-step for global plane gy: with a free surface every row comes from the image-method
// operators (their interior rows are scaled (c/DH)*DT, not c*(DT/DH): Derivatives.cpp:407-425 vs FDTD3D.cpp:211-216)
template <int Q, bool FWD> __device__ __forceinline__ void loadYWeightsVel(const WsParams &P, int gy, float (&w)[Q])
{
    constexpr int H = Q / 2;
    const int op = P.free_surface == 1 ? (FWD ? OP_YF_FS : OP_YB_FS) : (FWD ? OP_YF : OP_YB);
    const int row = (P.free_surface == 1 && gy < H) ? max(gy, 0) : H;
    const float *t = P.tab + ((size_t)op * (2 * H + 1) + row) * (Q + 1) + (FWD ? 1 : 0);
#pragma unroll
    for (int j = 0; j < Q; j++)
        w[j] = __ldg(t + j);
}
__device__ __forceinline__ int yCpmlIndex(const WsParams &P, int gy)
{
    int ky = wsCpmlIndex(gy, P.gny, P.W);
    if (P.free_surface != 0 && gy < P.W)
        ky = -1; // no CPML in the top layer below a free surface (CPML3D.cpp:320-328)
    return ky;
}
// true if planes [gy0, gy1] need the generic step (image-method rows or a y-CPML layer)
template <bool CPML> __device__ __forceinline__ bool needsGeneric(const WsParams &P, int gy0, int gy1, int H, bool velocity)
{
    if (velocity && P.free_surface == 1 && gy0 < H)
        return true;
    if (P.edge_policy == 1 && (gy0 < H || gy1 >= P.gny - H)) // order-reducing edges: the rows next to the grid edge have their own weights
        return true;
    if (CPML) {
        if (P.free_surface == 0 && gy0 < P.W)
            return true;
        if (gy1 >= P.gny - P.W)
            return true;
    }
    return false;
}

struct Bars {
    uint64_t full[NSTMAX], empty[NSTMAX];
    uint64_t xfull, xfree; // stress half-step: derivative exchange between the roles
};

// consumer bookkeeping shared by both half-steps
struct Thr {
    int lx, lz, x0, z, lane;
    bool active;
    uint32_t barFull, barEmpty; // shared addresses of full[0] / empty[0]
    uint32_t barXFull, barXFree;
    int stride;                 // floats per stage (operands + staged CPML memory variables)
    const float *cxTab;         // shared coefficient table of the x layers
    int oPX, oPZ, oPZ2;         // offsets inside a stage of this thread's staged memory variables (x row, z term, 2nd z term)
};
// developer trace (env WS_FAST_TRACE=file): per thread block {start ns, end ns, SM id} of the last launch of each half-step
__device__ __forceinline__ unsigned long long globalTimer()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void traceStart(const WsParams &P, int pass)
{
    if (P.fastTrace && threadIdx.x == 0) {
        const size_t b = (size_t)pass * WS_TRACE_MAX + (P.fastTileBase + blockIdx.x + (size_t)P.fastNTiles * blockIdx.z);
        if (b < (size_t)(pass + 1) * WS_TRACE_MAX) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.fastTrace[3 * b] = globalTimer();
            P.fastTrace[3 * b + 1] = 0;
            P.fastTrace[3 * b + 2] = smid;
        }
    }
}
__device__ __forceinline__ void traceEnd(const WsParams &P, int pass, int lane)
{
    if (P.fastTrace && lane == 0) {
        const size_t b = (size_t)pass * WS_TRACE_MAX + (P.fastTileBase + blockIdx.x + (size_t)P.fastNTiles * blockIdx.z);
        if (b < (size_t)(pass + 1) * WS_TRACE_MAX)
            atomicMax(P.fastTrace + 3 * b + 1, globalTimer());
    }
}
__device__ __forceinline__ void consumerRelease(const Thr &t, int stage)
{
    __syncwarp();
    if (t.lane == 0)
        mbarArrive(t.barEmpty + 8u * stage);
}

// x derivative with 9-tap weights (see above); `row` points at the tile column of x0 - HX
template <int Q> __device__ __forceinline__ F4 dX9(const float *row, const float *__restrict__ c)
{
    constexpr int H = Q / 2, HX = Cfg<Q>::HX, NV = (2 * HX + 4) / 4;
    float w[NV * 4];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const F4 t = ld4(row + 4 * k);
        w[4 * k] = t.v[0]; w[4 * k + 1] = t.v[1]; w[4 * k + 2] = t.v[2]; w[4 * k + 3] = t.v[3];
    }
    F4 r;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j <= Q; j++)
            acc = A::madd(c[j], w[HX + p + j - H], acc);
        r.v[p] = acc;
    }
    return r;
}

// Order-reducing edges (edge_policy 1, Derivatives.cpp:159-175): the columns / rows within q/2 of a grid face have their own
// weights (table rows, ws_tables.hpp).  They lie inside the x / z CPML layers, i.e. in layer tiles only; a thread that owns
// such a column or row takes these forms, with the accumulation order of the per-point kernels (all q + 1 window entries).
template <int Q> __device__ __forceinline__ F4 dX9t(const float *row, const float *__restrict__ tab, int op, int x0, int nx)
{
    constexpr int H = Q / 2, HX = Cfg<Q>::HX, NV = (2 * HX + 4) / 4;
    float w[NV * 4];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const F4 t = ld4(row + 4 * k);
        w[4 * k] = t.v[0]; w[4 * k + 1] = t.v[1]; w[4 * k + 2] = t.v[2]; w[4 * k + 3] = t.v[3];
    }
    F4 r;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const float *__restrict__ c = tab + ((size_t)op * (2 * H + 1) + wsRowClass(x0 + p, nx, H)) * (Q + 1);
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j <= Q; j++)
            acc = A::madd(__ldg(c + j), w[HX + p + j - H], acc);
        r.v[p] = acc;
    }
    return r;
}
// `col` points at (row of z - H, column of x0); c = the q + 1 weights of the row's class
template <int Q, int LD> __device__ __forceinline__ F4 dZ9t(const float *col, const float *__restrict__ c)
{
    F4 r = zero4();
#pragma unroll
    for (int j = 0; j <= Q; j++) {
        const F4 t = ld4(col + j * LD);
        const float wj = __ldg(c + j);
#pragma unroll
        for (int p = 0; p < 4; p++)
            r.v[p] = A::madd(wj, t.v[p], r.v[p]);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeFn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || !p)
            throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D map over an arena: (x, z, y, array); box = boxX x boxZ x 1 plane x boxF arrays
CUtensorMap makeMap(const float *arena, int pitch, int nzp, int nyp, long long arrayStride, int nArrays, int boxX, int boxZ, int boxF)
{
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)nzp, (cuuint64_t)nyp, (cuuint64_t)nArrays};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * nzp * 4, (cuuint64_t)arrayStride * 4};
    const cuuint32_t box[4] = {(cuuint32_t)boxX, (cuuint32_t)boxZ, 1, (cuuint32_t)boxF};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encodeFn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(arena), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

constexpr int kMaxDynSmem = 227 * 1024 - 2048; // 227 KB per block minus the static barriers and coefficient table
// dense 4-D map (d0 fastest, arrays outermost); box = b0 x b1 x 1 x bA
CUtensorMap makeMapG(const float *arena, int d0, int d1, int d2, int nArrays, int b0, int b1, int bA)
{
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2, (cuuint64_t)nArrays};
    const cuuint64_t strides[3] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4, (cuuint64_t)d0 * d1 * d2 * 4};
    const cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, 1, (cuuint32_t)bA};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encodeFn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(arena), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled (memory variables) failed with code " + std::to_string((int)r));
    return m;
}

// planes per thread block.  Chunks of ~64 planes are fastest (measured at 1024^3: 36.7 Gpt/s against 35.5 with 3 chunks
// per column, profiles/r01_march_chunk_sweep.txt): the thread blocks of one y range march in step, so the halo rows
// they share are still in L2 when the neighbouring tile asks for them; shorter chunks pay too many of the q-1
// feed-only planes and pipeline fills every chunk starts with.  With thousands of thread blocks per launch the length
// of the last wave no longer matters.
inline int wsPickChunk(int nTiles, int nyl, int q)
{
    (void)nTiles;
    (void)q;
    if (nyl < 96)
        return nyl;
    const int k = (nyl + 63) / 64;
    return (nyl + k - 1) / k;
}

} // namespace

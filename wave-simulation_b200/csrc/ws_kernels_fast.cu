// ws_kernels_fast.cu — tiled sm_100a kernels for the 3-D elastic hot path (FD3Delastic::run,
// ForwardSolver/ForwardSolver3Delastic.cpp:120-414): one fused kernel per half-step.
//
// Structure (2.5-D blocking):
//   * a thread block owns a TX x TZ tile of the x-z plane and marches along y (the slowest axis);
//   * every thread owns 4 consecutive x points (128-bit accesses) of one z row;
//   * the 3 fields that are differentiated along y in this half-step live in REGISTER QUEUES (Q planes deep);
//   * the tiles needed for the x and z stencils (with their halos) are staged in SHARED MEMORY by TMA
//     (cp.async.bulk.tensor.3d + mbarrier, 3-stage ring), prefetched two planes ahead of the compute;
//   * operands touched only at the own point (updated fields, model parameters, queue feed) are plain coalesced
//     128-bit global accesses issued at the top of the iteration, consumed at its end;
//   * off-grid taps read the zero pads of the HBM layout (StencilMatrix "drop off-grid taps", edge_policy 0);
//   * image-method free surface: per-plane y weights for the first q/2 planes + surface correction in the stress kernel;
//   * CPML: memory variables in compact boundary slabs, touched only by tiles / planes / rows inside the layers.
// The arithmetic sequence is the one of the general kernels (ws_kernels_general.cuh) in FMA mode, so both produce
// bit-identical results.
#include "../../include/wavesim.h"
#include "ws_launch.hpp"

#include <cuda.h>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int TX = 64, TZ = 16, NTHREADS = 256, NSTAGE = 3;

template <int Q> struct Cfg {
    static constexpr int H = Q / 2;
    static constexpr int HX = (H <= 4) ? 4 : 8; // x halo rounded to a float4
    static constexpr int TXH = TX + 2 * HX;
    static constexpr int TZH = TZ + 2 * H;
};
constexpr uint32_t al128(uint32_t v) { return (v + 127u) & ~127u; }

// tensor-map slots (one CUtensorMap per array and box shape)
enum {
    TM_SXX_X = 0, TM_SXY_X, TM_SXZ_XZ, TM_SYZ_Z, TM_SZZ_Z, // velocity half-step
    TM_VX_XZ, TM_VY_XZ, TM_VZ_XZ,                           // stress half-step
    TM_COUNT
};

__device__ __forceinline__ uint32_t smemU32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smemU32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmaLoad3D(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smemU32(dst)),
                 "l"((unsigned long long)map), "r"(smemU32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
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
__device__ __forceinline__ F4 zero4()
{
    F4 r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0f;
    return r;
}

using A = Ar<false>;

// x derivative of 4 consecutive points from a shared-memory row; `row` points at the tile column of x0 - HX
template <int Q, bool FWD> __device__ __forceinline__ F4 dX(const float *row, const float (&c)[Q])
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
// z derivative: `col` points at (row of z - H, column of x0) of a tile with z halo; rows are `ld` floats apart
template <int Q, bool FWD> __device__ __forceinline__ F4 dZ(const float *col, int ld, const float (&c)[Q])
{
    F4 r = zero4();
#pragma unroll
    for (int j = 0; j < Q; j++) {
        const F4 t = ld4(col + (FWD ? j + 1 : j) * ld);
#pragma unroll
        for (int p = 0; p < 4; p++)
            r.v[p] = A::madd(c[j], t.v[p], r.v[p]);
    }
    return r;
}
// y derivative from a register queue (q[k] = plane of the k-th tap), weights w[k]
template <int Q> __device__ __forceinline__ F4 dY(const F4 (&q)[Q], const float (&w)[Q])
{
    F4 r = zero4();
#pragma unroll
    for (int j = 0; j < Q; j++)
#pragma unroll
        for (int p = 0; p < 4; p++)
            r.v[p] = A::madd(w[j], q[j].v[p], r.v[p]);
    return r;
}
// CPML.cpp:84-95 on one value
__device__ __forceinline__ float cpml1(float d, float *ps, float a, float b)
{
    float v = A::mul(*ps, b);
    const float t = A::mul(a, d);
    v = A::add(v, t);
    *ps = v;
    return A::add(d, v);
}

struct Cp { // per-thread CPML bookkeeping
    int kx[4];
    int kz, ky;
    long long pxBase, pzBase; // psi offsets without the ly-dependent part
};

template <bool CPML> __device__ __forceinline__ void cpX(const WsParams &P, const Cp &cp, F4 &d, int slot, bool half, long long lyTerm)
{
    if (!CPML)
        return;
    const float *ca = half ? P.caxh : P.cax, *cb = half ? P.cbxh : P.cbx;
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (cp.kx[p] >= 0)
            d.v[p] = cpml1(d.v[p], P.psi[slot] + lyTerm + cp.pxBase + cp.kx[p], __ldg(ca + cp.kx[p]), __ldg(cb + cp.kx[p]));
}
template <bool CPML> __device__ __forceinline__ void cpY(const WsParams &P, const Cp &cp, F4 &d, int slot, bool half, long long pyOff)
{
    if (!CPML || cp.ky < 0)
        return;
    const float a = __ldg((half ? P.cayh : P.cay) + cp.ky), b = __ldg((half ? P.cbyh : P.cby) + cp.ky);
    float *ps = P.psi[slot] + pyOff;
    F4 old = ld4(ps), nw;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float v = A::mul(old.v[p], b);
        const float t = A::mul(a, d.v[p]);
        v = A::add(v, t);
        nw.v[p] = v;
        d.v[p] = A::add(d.v[p], v);
    }
    st4(ps, nw);
}
template <bool CPML> __device__ __forceinline__ void cpZ(const WsParams &P, const Cp &cp, F4 &d, int slot, bool half, long long pzOff)
{
    if (!CPML || cp.kz < 0)
        return;
    const float a = __ldg((half ? P.cazh : P.caz) + cp.kz), b = __ldg((half ? P.cbzh : P.cbz) + cp.kz);
    float *ps = P.psi[slot] + pzOff;
    F4 old = ld4(ps), nw;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float v = A::mul(old.v[p], b);
        const float t = A::mul(a, d.v[p]);
        v = A::add(v, t);
        nw.v[p] = v;
        d.v[p] = A::add(d.v[p], v);
    }
    st4(ps, nw);
}

template <int Q> struct SmemA { // velocity half-step stage layout (bytes)
    using C = Cfg<Q>;
    static constexpr uint32_t SZ_X = al128(TZ * C::TXH * 4), SZ_XZ = al128(C::TZH * C::TXH * 4), SZ_Z = al128(C::TZH * TX * 4);
    static constexpr uint32_t OFF_SXX = 0, OFF_SXY = SZ_X, OFF_SXZ = 2 * SZ_X, OFF_SYZ = 2 * SZ_X + SZ_XZ, OFF_SZZ = OFF_SYZ + SZ_Z;
    static constexpr uint32_t STAGE = OFF_SZZ + SZ_Z;
    static constexpr uint32_t TXBYTES = 2 * TZ * C::TXH * 4 + C::TZH * C::TXH * 4 + 2 * C::TZH * TX * 4;
};
template <int Q> struct SmemB { // stress half-step stage layout
    using C = Cfg<Q>;
    static constexpr uint32_t SZ_XZ = al128(C::TZH * C::TXH * 4);
    static constexpr uint32_t OFF_VX = 0, OFF_VY = SZ_XZ, OFF_VZ = 2 * SZ_XZ, STAGE = 3 * SZ_XZ;
    static constexpr uint32_t TXBYTES = 3 * C::TZH * C::TXH * 4;
};

// ---------------------------------------------------------------------------------------------------------------------
// velocity half-step (ForwardSolver3Delastic.cpp:181-277)
// ---------------------------------------------------------------------------------------------------------------------
template <int Q, bool CPML> __global__ void __launch_bounds__(NTHREADS, 1) kFastVel(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = SmemA<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) uint64_t bars[NSTAGE];
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TX, tz0 = blockIdx.y * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    const int lx = tid & 15, lz = tid >> 4;
    const int x0 = tx0 + 4 * lx, z = tz0 + lz;
    const bool active = (x0 < P.nx) && (z < P.nz);
    const int HZP = (P.nzp > 1) ? WS_HALO : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++)
            mbarInit(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int ly, int stage) {
        unsigned char *st = smraw + stage * S::STAGE;
        uint64_t *bar = &bars[stage];
        mbarExpectTx(bar, S::TXBYTES);
        const int cy = WS_HALO + ly, cx = WS_PADX + tx0, cz = HZP + tz0;
        tmaLoad3D(st + S::OFF_SXX, &maps[TM_SXX_X], bar, cx - HX, cz, cy);
        tmaLoad3D(st + S::OFF_SXY, &maps[TM_SXY_X], bar, cx - HX, cz, cy);
        tmaLoad3D(st + S::OFF_SXZ, &maps[TM_SXZ_XZ], bar, cx - HX, cz - H, cy);
        tmaLoad3D(st + S::OFF_SYZ, &maps[TM_SYZ_Z], bar, cx, cz - H, cy);
        tmaLoad3D(st + S::OFF_SZZ, &maps[TM_SZZ_Z], bar, cx, cz - H, cy);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE - 1; s++)
            if (yc0 + s < yc1)
                issue(yc0 + s, s);
    }

    // interior weights (same on every axis; policy 0): forward taps are table indices 1..Q of the interior row
    float c[Q];
    {
        const float *w = P.tab + ((size_t)OP_XF * (2 * H + 1) + H) * (Q + 1);
#pragma unroll
        for (int j = 0; j < Q; j++)
            c[j] = __ldg(w + 1 + j);
    }
    const long long rowOff = P.base + x0 + (long long)z * P.pitch;
    const float *gsxy = P.fld[F_SXY] + rowOff, *gsyy = P.fld[F_SYY] + rowOff, *gsyz = P.fld[F_SYZ] + rowOff;
    float *gvx = P.fld[F_VX] + rowOff, *gvy = P.fld[F_VY] + rowOff, *gvz = P.fld[F_VZ] + rowOff;
    const float *grx = P.mat[M_RIX] + rowOff, *gry = P.mat[M_RIY] + rowOff, *grz = P.mat[M_RIZ] + rowOff;

    Cp cp;
    cp.kz = cp.ky = -1;
    cp.pxBase = cp.pzBase = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
        cp.kx[p] = -1;
    bool anyX = false;
    if (CPML && active) {
        const int W = P.W;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            cp.kx[p] = wsCpmlIndex(x0 + p, P.nx, W);
            anyX |= cp.kx[p] >= 0;
        }
        cp.kz = wsCpmlIndex(z, P.nz, W);
        cp.pxBase = (long long)z * (2 * W);
        cp.pzBase = (long long)cp.kz * P.nx + x0;
    }

    // register queues: qxy/qyz[k] = plane y-H+k (backward window), qyy[k] = plane y-H+1+k (forward window)
    F4 qxy[Q], qyy[Q], qyz[Q];
#pragma unroll
    for (int k = 0; k < Q - 1; k++) {
        if (active) {
            qxy[k] = ldg4(gsxy + (long long)(yc0 - H + k) * P.plane);
            qyz[k] = ldg4(gsyz + (long long)(yc0 - H + k) * P.plane);
            qyy[k] = ldg4(gsyy + (long long)(yc0 - H + 1 + k) * P.plane);
        } else {
            qxy[k] = zero4(); qyz[k] = zero4(); qyy[k] = zero4();
        }
    }

    int stage = 0;
    uint32_t parity = 0;
    for (int ly = yc0; ly < yc1; ly++) {
        const int gy = P.gy0 + ly;
        const long long po = (long long)ly * P.plane;
        // own-point operands of this plane, issued before the barrier wait
        F4 vx, vy, vz, rx, ry, rz;
        if (active) {
            qxy[Q - 1] = ldg4(gsxy + (long long)(ly + H - 1) * P.plane);
            qyz[Q - 1] = ldg4(gsyz + (long long)(ly + H - 1) * P.plane);
            qyy[Q - 1] = ldg4(gsyy + (long long)(ly + H) * P.plane);
            vx = ld4(gvx + po); vy = ld4(gvy + po); vz = ld4(gvz + po);
            rx = ldg4(grx + po); ry = ldg4(gry + po); rz = ldg4(grz + po);
        } else {
            qxy[Q - 1] = zero4(); qyz[Q - 1] = zero4(); qyy[Q - 1] = zero4();
            vx = vy = vz = rx = ry = rz = zero4();
        }
        // per-plane y weights: image method on the first H planes below the free surface, interior weights elsewhere
        float wyb[Q], wyf[Q];
        if (P.free_surface == 1 && gy < H) {
            const float *tb = P.tab + ((size_t)OP_YB_FS * (2 * H + 1) + gy) * (Q + 1);
            const float *tf = P.tab + ((size_t)OP_YF_FS * (2 * H + 1) + gy) * (Q + 1);
#pragma unroll
            for (int j = 0; j < Q; j++) {
                wyb[j] = __ldg(tb + j);
                wyf[j] = __ldg(tf + 1 + j);
            }
        } else {
#pragma unroll
            for (int j = 0; j < Q; j++) {
                wyb[j] = c[j];
                wyf[j] = c[j];
            }
        }
        if (CPML) {
            cp.ky = wsCpmlIndex(gy, P.gny, P.W);
            if (P.free_surface != 0 && gy < P.W)
                cp.ky = -1;
        }
        const long long pxLy = (long long)ly * P.nz * (2 * P.W);
        const long long pyOff = ((long long)cp.ky * P.nz + z) * P.nx + x0;
        const long long pzOff = (long long)ly * (2 * P.W) * P.nx + cp.pzBase;

        mbarWait(&bars[stage], parity);
        const unsigned char *st = smraw + stage * S::STAGE;
        const float *tSxx = reinterpret_cast<const float *>(st + S::OFF_SXX) + lz * TXH + 4 * lx;
        const float *tSxy = reinterpret_cast<const float *>(st + S::OFF_SXY) + lz * TXH + 4 * lx;
        const float *tSxzX = reinterpret_cast<const float *>(st + S::OFF_SXZ) + (lz + H) * TXH + 4 * lx;
        const float *tSxzZ = reinterpret_cast<const float *>(st + S::OFF_SXZ) + lz * TXH + 4 * lx + HX;
        const float *tSyz = reinterpret_cast<const float *>(st + S::OFF_SYZ) + lz * TX + 4 * lx;
        const float *tSzz = reinterpret_cast<const float *>(st + S::OFF_SZZ) + lz * TX + 4 * lx;

        // ---- vx ----
        F4 u = dX<Q, true>(tSxx, c);
        if (anyX) cpX<CPML>(P, cp, u, PSI_SXX_X, true, pxLy);
        F4 w = dY<Q>(qxy, wyb);
        cpY<CPML>(P, cp, w, PSI_SXY_Y, false, pyOff);
#pragma unroll
        for (int p = 0; p < 4; p++) u.v[p] = A::add(u.v[p], w.v[p]);
        w = dZ<Q, false>(tSxzZ, TXH, c);
        cpZ<CPML>(P, cp, w, PSI_SXZ_Z, false, pzOff);
#pragma unroll
        for (int p = 0; p < 4; p++) {
            u.v[p] = A::add(u.v[p], w.v[p]);
            u.v[p] = A::mul(u.v[p], rx.v[p]);
            vx.v[p] = A::add(vx.v[p], u.v[p]);
        }
        // ---- vy ----
        u = dX<Q, false>(tSxy, c);
        if (anyX) cpX<CPML>(P, cp, u, PSI_SXY_X, false, pxLy);
        w = dY<Q>(qyy, wyf);
        cpY<CPML>(P, cp, w, PSI_SYY_Y, true, pyOff);
#pragma unroll
        for (int p = 0; p < 4; p++) u.v[p] = A::add(u.v[p], w.v[p]);
        w = dZ<Q, false>(tSyz, TX, c);
        cpZ<CPML>(P, cp, w, PSI_SYZ_Z, false, pzOff);
#pragma unroll
        for (int p = 0; p < 4; p++) {
            u.v[p] = A::add(u.v[p], w.v[p]);
            u.v[p] = A::mul(u.v[p], ry.v[p]);
            vy.v[p] = A::add(vy.v[p], u.v[p]);
        }
        // ---- vz ----
        u = dX<Q, false>(tSxzX, c);
        if (anyX) cpX<CPML>(P, cp, u, PSI_SXZ_X, false, pxLy);
        w = dY<Q>(qyz, wyb);
        cpY<CPML>(P, cp, w, PSI_SYZ_Y, false, pyOff);
#pragma unroll
        for (int p = 0; p < 4; p++) u.v[p] = A::add(u.v[p], w.v[p]);
        w = dZ<Q, true>(tSzz, TX, c);
        cpZ<CPML>(P, cp, w, PSI_SZZ_Z, true, pzOff);
#pragma unroll
        for (int p = 0; p < 4; p++) {
            u.v[p] = A::add(u.v[p], w.v[p]);
            u.v[p] = A::mul(u.v[p], rz.v[p]);
            vz.v[p] = A::add(vz.v[p], u.v[p]);
        }
        if (active) {
            st4(gvx + po, vx); st4(gvy + po, vy); st4(gvz + po, vz);
        }
#pragma unroll
        for (int k = 0; k < Q - 1; k++) {
            qxy[k] = qxy[k + 1]; qyy[k] = qyy[k + 1]; qyz[k] = qyz[k + 1];
        }
        __syncthreads(); // everybody is done with this stage
        if (tid == 0 && ly + NSTAGE - 1 < yc1)
            issue(ly + NSTAGE - 1, (stage + NSTAGE - 1) % NSTAGE);
        stage++;
        if (stage == NSTAGE) {
            stage = 0;
            parity ^= 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stress half-step (ForwardSolver3Delastic.cpp:288-404) incl. free-surface correction
// ---------------------------------------------------------------------------------------------------------------------
template <int Q, bool CPML> __global__ void __launch_bounds__(NTHREADS, 1) kFastStress(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = SmemB<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) uint64_t bars[NSTAGE];
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TX, tz0 = blockIdx.y * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    const int lx = tid & 15, lz = tid >> 4;
    const int x0 = tx0 + 4 * lx, z = tz0 + lz;
    const bool active = (x0 < P.nx) && (z < P.nz);
    const int HZP = (P.nzp > 1) ? WS_HALO : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++)
            mbarInit(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int ly, int stage) {
        unsigned char *st = smraw + stage * S::STAGE;
        uint64_t *bar = &bars[stage];
        mbarExpectTx(bar, S::TXBYTES);
        const int cy = WS_HALO + ly, cx = WS_PADX + tx0 - HX, cz = HZP + tz0 - H;
        tmaLoad3D(st + S::OFF_VX, &maps[TM_VX_XZ], bar, cx, cz, cy);
        tmaLoad3D(st + S::OFF_VY, &maps[TM_VY_XZ], bar, cx, cz, cy);
        tmaLoad3D(st + S::OFF_VZ, &maps[TM_VZ_XZ], bar, cx, cz, cy);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE - 1; s++)
            if (yc0 + s < yc1)
                issue(yc0 + s, s);
    }
    float c[Q];
    {
        const float *w = P.tab + ((size_t)OP_XF * (2 * H + 1) + H) * (Q + 1);
#pragma unroll
        for (int j = 0; j < Q; j++)
            c[j] = __ldg(w + 1 + j);
    }
    const long long rowOff = P.base + x0 + (long long)z * P.pitch;
    const float *gvx = P.fld[F_VX] + rowOff, *gvy = P.fld[F_VY] + rowOff, *gvz = P.fld[F_VZ] + rowOff;

    Cp cp;
    cp.kz = cp.ky = -1;
    cp.pxBase = cp.pzBase = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
        cp.kx[p] = -1;
    bool anyX = false;
    if (CPML && active) {
        const int W = P.W;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            cp.kx[p] = wsCpmlIndex(x0 + p, P.nx, W);
            anyX |= cp.kx[p] >= 0;
        }
        cp.kz = wsCpmlIndex(z, P.nz, W);
        cp.pxBase = (long long)z * (2 * W);
        cp.pzBase = (long long)cp.kz * P.nx + x0;
    }
    // queues: qvx/qvz[k] = plane y-H+1+k (forward window), qvy[k] = plane y-H+k (backward window)
    F4 qvx[Q], qvy[Q], qvz[Q];
#pragma unroll
    for (int k = 0; k < Q - 1; k++) {
        if (active) {
            qvx[k] = ldg4(gvx + (long long)(yc0 - H + 1 + k) * P.plane);
            qvz[k] = ldg4(gvz + (long long)(yc0 - H + 1 + k) * P.plane);
            qvy[k] = ldg4(gvy + (long long)(yc0 - H + k) * P.plane);
        } else {
            qvx[k] = zero4(); qvy[k] = zero4(); qvz[k] = zero4();
        }
    }
    int stage = 0;
    uint32_t parity = 0;
    for (int ly = yc0; ly < yc1; ly++) {
        const int gy = P.gy0 + ly;
        const long long o = rowOff + (long long)ly * P.plane;
        F4 sxx, syy, szz, sxy, sxz, syz, pi, mu, mxy, mxz, myz;
        if (active) {
            qvx[Q - 1] = ldg4(gvx + (long long)(ly + H) * P.plane);
            qvz[Q - 1] = ldg4(gvz + (long long)(ly + H) * P.plane);
            qvy[Q - 1] = ldg4(gvy + (long long)(ly + H - 1) * P.plane);
            sxx = ld4(P.fld[F_SXX] + o); syy = ld4(P.fld[F_SYY] + o); szz = ld4(P.fld[F_SZZ] + o);
            sxy = ld4(P.fld[F_SXY] + o); sxz = ld4(P.fld[F_SXZ] + o); syz = ld4(P.fld[F_SYZ] + o);
            pi = ldg4(P.mat[M_PW] + o); mu = ldg4(P.mat[M_MU] + o);
            mxy = ldg4(P.mat[M_MUXY] + o); mxz = ldg4(P.mat[M_MUXZ] + o); myz = ldg4(P.mat[M_MUYZ] + o);
        } else {
            qvx[Q - 1] = zero4(); qvy[Q - 1] = zero4(); qvz[Q - 1] = zero4();
            sxx = syy = szz = sxy = sxz = syz = pi = mu = mxy = mxz = myz = zero4();
        }
        if (CPML) {
            cp.ky = wsCpmlIndex(gy, P.gny, P.W);
            if (P.free_surface != 0 && gy < P.W)
                cp.ky = -1;
        }
        const long long pxLy = (long long)ly * P.nz * (2 * P.W);
        const long long pyOff = ((long long)cp.ky * P.nz + z) * P.nx + x0;
        const long long pzOff = (long long)ly * (2 * P.W) * P.nx + cp.pzBase;

        mbarWait(&bars[stage], parity);
        const unsigned char *st = smraw + stage * S::STAGE;
        const float *bvx = reinterpret_cast<const float *>(st + S::OFF_VX), *bvy = reinterpret_cast<const float *>(st + S::OFF_VY),
                    *bvz = reinterpret_cast<const float *>(st + S::OFF_VZ);
        const int oX = (lz + H) * TXH + 4 * lx;     // x stencil: own row, column of x0 - HX
        const int oZ = lz * TXH + 4 * lx + HX;      // z stencil: row z - H, own column

        // normal strain rates; the y derivative of vy is the plain operator even below a free surface (:289)
        F4 vxx = dX<Q, false>(bvx + oX, c);
        F4 vyy = dY<Q>(qvy, c);
        F4 vzz = dZ<Q, false>(bvz + oZ, TXH, c);
        if (anyX) cpX<CPML>(P, cp, vxx, PSI_VXX, false, pxLy);
        cpY<CPML>(P, cp, vyy, PSI_VYY, false, pyOff);
        cpZ<CPML>(P, cp, vzz, PSI_VZZ, false, pzOff);
#pragma unroll
        for (int p = 0; p < 4; p++) {
            float u = A::add(vxx.v[p], vyy.v[p]);
            u = A::add(u, vzz.v[p]);
            u = A::mul(u, pi.v[p]);
            sxx.v[p] = A::add(sxx.v[p], u);
            syy.v[p] = A::add(syy.v[p], u);
            szz.v[p] = A::add(szz.v[p], u);
            u = A::mul(A::add(vyy.v[p], vzz.v[p]), mu.v[p]);
            sxx.v[p] = A::msub(2.0f, u, sxx.v[p]);
            u = A::mul(A::add(vxx.v[p], vzz.v[p]), mu.v[p]);
            syy.v[p] = A::msub(2.0f, u, syy.v[p]);
            u = A::mul(A::add(vxx.v[p], vyy.v[p]), mu.v[p]);
            szz.v[p] = A::msub(2.0f, u, szz.v[p]);
        }
        // shear stresses
        {
            F4 u = dY<Q>(qvx, c);
            cpY<CPML>(P, cp, u, PSI_VXY, true, pyOff);
            F4 w = dX<Q, true>(bvy + oX, c);
            if (anyX) cpX<CPML>(P, cp, w, PSI_VYX, true, pxLy);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float t = A::add(u.v[p], w.v[p]);
                sxy.v[p] = A::add(sxy.v[p], A::mul(t, mxy.v[p]));
            }
            u = dZ<Q, true>(bvx + oZ, TXH, c);
            cpZ<CPML>(P, cp, u, PSI_VXZ, true, pzOff);
            w = dX<Q, true>(bvz + oX, c);
            if (anyX) cpX<CPML>(P, cp, w, PSI_VZX, true, pxLy);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float t = A::add(u.v[p], w.v[p]);
                sxz.v[p] = A::add(sxz.v[p], A::mul(t, mxz.v[p]));
            }
            u = dZ<Q, true>(bvy + oZ, TXH, c);
            cpZ<CPML>(P, cp, u, PSI_VYZ, true, pzOff);
            w = dY<Q>(qvz, c);
            cpY<CPML>(P, cp, w, PSI_VZY, true, pyOff);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float t = A::add(u.v[p], w.v[p]);
                syz.v[p] = A::add(syz.v[p], A::mul(t, myz.v[p]));
            }
        }
        if (P.free_surface == 1 && gy == 0 && active) {
            // FreeSurface3Delastic.cpp:15-47, FreeSurface.cpp:13-20
            const F4 sH = ldg4(P.sH + (long long)z * P.nx + x0), sV = ldg4(P.sV + (long long)z * P.nx + x0);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float hor = A::add(vxx.v[p], vzz.v[p]);
                float t = A::mul(sH.v[p], hor);
                sxx.v[p] = A::add(sxx.v[p], t);
                szz.v[p] = A::add(szz.v[p], t);
                t = A::mul(sV.v[p], vyy.v[p]);
                sxx.v[p] = A::sub(sxx.v[p], t);
                szz.v[p] = A::sub(szz.v[p], t);
                syy.v[p] = A::mul(syy.v[p], 0.0f);
            }
        }
        if (active) {
            st4(P.fld[F_SXX] + o, sxx); st4(P.fld[F_SYY] + o, syy); st4(P.fld[F_SZZ] + o, szz);
            st4(P.fld[F_SXY] + o, sxy); st4(P.fld[F_SXZ] + o, sxz); st4(P.fld[F_SYZ] + o, syz);
        }
#pragma unroll
        for (int k = 0; k < Q - 1; k++) {
            qvx[k] = qvx[k + 1]; qvy[k] = qvy[k + 1]; qvz[k] = qvz[k + 1];
        }
        __syncthreads();
        if (tid == 0 && ly + NSTAGE - 1 < yc1)
            issue(ly + NSTAGE - 1, (stage + NSTAGE - 1) % NSTAGE);
        stage++;
        if (stage == NSTAGE) {
            stage = 0;
            parity ^= 1;
        }
    }
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

CUtensorMap makeMap(const float *base, int pitch, int nzp, int nyp, int boxX, int boxZ)
{
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)nzp, (cuuint64_t)nyp};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * nzp * 4};
    const cuuint32_t box[3] = {(cuuint32_t)boxX, (cuuint32_t)boxZ, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encodeFn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

template <int Q> void setAttrs()
{
    static bool done = false;
    if (done)
        return;
    cudaFuncSetAttribute(kFastVel<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * SmemA<Q>::STAGE);
    cudaFuncSetAttribute(kFastVel<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * SmemA<Q>::STAGE);
    cudaFuncSetAttribute(kFastStress<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * SmemB<Q>::STAGE);
    cudaFuncSetAttribute(kFastStress<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * SmemB<Q>::STAGE);
    done = true;
}

template <int Q> void launchQ(const WsParams &P, int pass, cudaStream_t st)
{
    setAttrs<Q>();
    const int ny = P.yhi - P.ylo;
    dim3 grid((P.nx + TX - 1) / TX, (P.nz + TZ - 1) / TZ, (ny + P.fastChunk - 1) / P.fastChunk);
    const bool cpml = P.damping == 2;
    if (pass == 0) {
        const size_t sm = NSTAGE * SmemA<Q>::STAGE;
        if (cpml)
            kFastVel<Q, true><<<grid, NTHREADS, sm, st>>>(P);
        else
            kFastVel<Q, false><<<grid, NTHREADS, sm, st>>>(P);
    } else {
        const size_t sm = NSTAGE * SmemB<Q>::STAGE;
        if (cpml)
            kFastStress<Q, true><<<grid, NTHREADS, sm, st>>>(P);
        else
            kFastStress<Q, false><<<grid, NTHREADS, sm, st>>>(P);
    }
}

} // namespace

bool wsFastSupported(const WsParams &P, bool exact)
{
    if (exact || P.eq != WS_EQ_ELASTIC || P.dim != 3)
        return false;
    if (P.edge_policy != 0 || P.damping == 1)
        return false;
    if (P.q != 8 && P.q != 4)
        return false;
    if (P.nx % 4 != 0)
        return false;
    return true;
}

void *wsFastPrepare(WsParams &P, int nyp)
{
    const int H = P.h, HX = H <= 4 ? 4 : 8;
    const int TXH = TX + 2 * HX, TZH = TZ + 2 * H;
    std::vector<CUtensorMap> maps(TM_COUNT);
    auto mk = [&](int slot, const float *base, int bx, int bz) { maps[slot] = makeMap(base, P.pitch, P.nzp, nyp, bx, bz); };
    mk(TM_SXX_X, P.fld[F_SXX], TXH, TZ);
    mk(TM_SXY_X, P.fld[F_SXY], TXH, TZ);
    mk(TM_SXZ_XZ, P.fld[F_SXZ], TXH, TZH);
    mk(TM_SYZ_Z, P.fld[F_SYZ], TX, TZH);
    mk(TM_SZZ_Z, P.fld[F_SZZ], TX, TZH);
    mk(TM_VX_XZ, P.fld[F_VX], TXH, TZH);
    mk(TM_VY_XZ, P.fld[F_VY], TXH, TZH);
    mk(TM_VZ_XZ, P.fld[F_VZ], TXH, TZH);
    void *dev = nullptr;
    if (cudaMalloc(&dev, sizeof(CUtensorMap) * TM_COUNT) != cudaSuccess)
        throw std::runtime_error("cudaMalloc for tensor maps failed");
    cudaMemcpy(dev, maps.data(), sizeof(CUtensorMap) * TM_COUNT, cudaMemcpyHostToDevice);
    P.fastMaps = dev;
    // planes per block: enough blocks to fill 148 SMs a few times over, long enough marches to amortise the prologue
    const int tiles = ((P.nx + TX - 1) / TX) * ((P.nz + TZ - 1) / TZ);
    int chunks = (4 * 148 + tiles - 1) / tiles;
    if (chunks < 1)
        chunks = 1;
    int chunk = (P.nyl + chunks - 1) / chunks;
    if (chunk < 16)
        chunk = 16;
    P.fastChunk = chunk;
    return dev;
}

void wsFastRelease(void *maps)
{
    if (maps)
        cudaFree(maps);
}

bool wsLaunchFast(const WsParams &P, int pass, cudaStream_t st)
{
    if (!P.fastMaps || P.yhi <= P.ylo)
        return false;
    if (P.q == 8)
        launchQ<8>(P, pass, st);
    else if (P.q == 4)
        launchQ<4>(P, pass, st);
    else
        return false;
    return true;
}

// ws_kernels_fast.cu — warp-specialised sm_100a kernels for the 3-D elastic hot path (FD3Delastic::run,
// ForwardSolver/ForwardSolver3Delastic.cpp:120-414): one fused kernel per half-step.
//
// Structure (2.5-D blocking, producer / consumer pipeline):
//   * a thread block owns a TX x TZ tile of the x-z plane and marches along y (the slowest axis);
//   * ONE PRODUCER WARP streams, per plane, every operand of that plane into a ring of NST shared-memory stages with TMA
//     (cp.async.bulk.tensor.3d, completion on a "full" mbarrier per stage): halo tiles for the x / z stencils, the plane
//     that enters the y-derivative window, and the own-point operands (updated fields, model parameters).  Consumer
//     threads never wait on DRAM: up to NST-1 planes (~100 KB per SM) are in flight while one is being computed;
//   * CONSUMER GROUPS split the work by OUTPUT COMPONENT (velocity half-step: vx | vy | vz; stress half-step:
//     sxx,syy,szz | sxy | sxz | syz).  Every thread owns 4 consecutive x points (128-bit shared / global accesses) and
//     keeps only the ONE field it differentiates along y in a register queue (Q planes deep), so the kernels run with
//     ~100 registers and 13-17 warps per SM instead of 230 registers and 8 warps;
//   * a stage is handed back to the producer through an "empty" mbarrier (one arrival per consumer warp);
//   * off-grid taps read the zero pads of the HBM layout (StencilMatrix "drop off-grid taps", edge_policy 0);
//   * image-method free surface: per-plane y weights (the reference's DyfFreeSurface / DybFreeSurface rows) + surface
//     correction in the stress kernel; CPML: memory variables in compact boundary slabs, loaded before the stage wait.
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

constexpr int TX = 64, TZ = 8, NST = 4;
constexpr int NG_VEL = 3, NG_STR = 3;

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
};

// tensor-map slots (one CUtensorMap per array and box shape)
enum {
    TM_SXX_X = 0, TM_SXY_X, TM_SXZ_XZ, TM_SYZ_Z, TM_SZZ_Z, TM_SXY_P, TM_SYY_P, TM_SYZ_P, // velocity half-step
    TM_VX_P, TM_VY_P, TM_VZ_P, TM_RIX_P, TM_RIY_P, TM_RIZ_P,
    TM_VX_XZ, TM_VY_XZ, TM_VZ_XZ,                                                         // stress half-step
    TM_SXX_P, TM_SZZ_P, TM_SXZ_P, TM_PW_P, TM_MU_P, TM_MUXY_P, TM_MUXZ_P, TM_MUYZ_P,
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
__device__ __forceinline__ void mbarArrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemU32(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "DONE_%=:\n"
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
__device__ __forceinline__ void tmaPrefetchDesc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)map) : "memory");
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

// ---------------------------------------------------------------------------------------------------------------------
// CPML (CPML.cpp:84-95 applyCPML): psi = b psi + a d ; d = d + psi.  The memory variables live in compact slabs
// (x: [ly][z][2W], y: [2W][z][x], z: [ly][2W][x]); their loads are issued before the stage wait.
// ---------------------------------------------------------------------------------------------------------------------
struct CpT { // per-thread, constant over the march
    bool active, anyX;
    int kx[4], kz;
    long long pxBase, pzBase;
    float xa[4], xb[4], za, zb;
};
struct CpI { // per iteration
    int ky;
    float ya, yb;
    long long pxOff, pyOff, pzOff;
    float px[4];
    F4 py, pz, pz2;
};

template <bool CPML> __device__ __forceinline__ void cpSetup(const WsParams &P, CpT &t, bool active, int x0, int z, bool halfX, bool halfZ)
{
    t.active = active;
    t.anyX = false;
    t.kz = -1;
    t.pxBase = t.pzBase = 0;
    t.za = t.zb = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        t.kx[p] = -1;
        t.xa[p] = t.xb[p] = 0.0f;
    }
    if (!CPML || !active)
        return;
    const int W = P.W;
    const float *ca = halfX ? P.caxh : P.cax, *cb = halfX ? P.cbxh : P.cbx;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        t.kx[p] = wsCpmlIndex(x0 + p, P.nx, W);
        if (t.kx[p] >= 0) {
            t.anyX = true;
            t.xa[p] = __ldg(ca + t.kx[p]);
            t.xb[p] = __ldg(cb + t.kx[p]);
        }
    }
    t.kz = wsCpmlIndex(z, P.nz, W);
    if (t.kz >= 0) {
        t.za = __ldg((halfZ ? P.cazh : P.caz) + t.kz);
        t.zb = __ldg((halfZ ? P.cbzh : P.cbz) + t.kz);
    }
    t.pxBase = (long long)z * (2 * W);
    t.pzBase = (long long)t.kz * P.nx + x0;
}
// issue the loads of this plane's memory variables (x slot sx, y slot sy, z slot sz)
template <bool CPML, int sx, int sy, int sz, int sz2 = -1> __device__ __forceinline__ void cpLoad(const WsParams &P, const CpT &t, CpI &it, int ly, int gy, int x0, int z, bool halfY)
{
    it.ky = -1;
    if (!CPML || !t.active)
        return;
    const int W = P.W;
    it.ky = wsCpmlIndex(gy, P.gny, W);
    if (P.free_surface != 0 && gy < W)
        it.ky = -1; // no CPML in the top layer below a free surface (CPML3D.cpp:320-328)
    it.pxOff = (long long)ly * P.nz * (2 * W) + t.pxBase;
    it.pzOff = (long long)ly * (2 * W) * P.nx + t.pzBase;
    it.pyOff = ((long long)it.ky * P.nz + z) * P.nx + x0;
    if (sx >= 0 && t.anyX) {
#pragma unroll
        for (int p = 0; p < 4; p++)
            if (t.kx[p] >= 0)
                it.px[p] = P.psi[sx >= 0 ? sx : 0][it.pxOff + t.kx[p]];
    }
    if (sy >= 0 && it.ky >= 0) {
        it.ya = __ldg((halfY ? P.cayh : P.cay) + it.ky);
        it.yb = __ldg((halfY ? P.cbyh : P.cby) + it.ky);
        it.py = ld4(P.psi[sy >= 0 ? sy : 0] + it.pyOff);
    }
    if (sz >= 0 && t.kz >= 0)
        it.pz = ld4(P.psi[sz >= 0 ? sz : 0] + it.pzOff);
    if (sz2 >= 0 && t.kz >= 0)
        it.pz2 = ld4(P.psi[sz2 >= 0 ? sz2 : 0] + it.pzOff);
}
template <bool CPML> __device__ __forceinline__ void cpApplyX(const WsParams &P, const CpT &t, const CpI &it, F4 &d, int slot)
{
    if (!CPML || !t.anyX)
        return;
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (t.kx[p] >= 0) {
            float v = A::mul(it.px[p], t.xb[p]);
            v = A::add(v, A::mul(t.xa[p], d.v[p]));
            P.psi[slot][it.pxOff + t.kx[p]] = v;
            d.v[p] = A::add(d.v[p], v);
        }
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
template <bool CPML> __device__ __forceinline__ void cpApplyY(const WsParams &P, const CpT &t, const CpI &it, F4 &d, int slot)
{
    if (!CPML || !t.active || it.ky < 0)
        return;
    cpApply4(P.psi[slot] + it.pyOff, it.py, it.ya, it.yb, d);
}
template <bool CPML> __device__ __forceinline__ void cpApplyZ(const WsParams &P, const CpT &t, const CpI &it, F4 &d, int slot)
{
    if (!CPML || t.kz < 0)
        return;
    cpApply4(P.psi[slot] + it.pzOff, it.pz, t.za, t.zb, d);
}
template <bool CPML> __device__ __forceinline__ void cpApplyZ2(const WsParams &P, const CpT &t, const CpI &it, F4 &d, int slot)
{
    if (!CPML || t.kz < 0)
        return;
    cpApply4(P.psi[slot] + it.pzOff, it.pz2, t.za, t.zb, d);
}
// interior weights (same on every axis; policy 0): forward taps are table indices 1..Q of the interior row
template <int Q> __device__ __forceinline__ void loadInterior(const WsParams &P, float (&c)[Q])
{
    constexpr int H = Q / 2;
    const float *w = P.tab + ((size_t)OP_XF * (2 * H + 1) + H) * (Q + 1);
#pragma unroll
    for (int j = 0; j < Q; j++)
        c[j] = __ldg(w + 1 + j);
}
// y weights of the velocity half-step for global plane gy: with a free surface every row comes from the image-method
// operators (their interior rows are scaled (c/DH)*DT, not c*(DT/DH): Derivatives.cpp:407-425 vs FDTD3D.cpp:211-216)
template <int Q, bool FWD> __device__ __forceinline__ void loadYWeights(const WsParams &P, int gy, float (&w)[Q])
{
    constexpr int H = Q / 2;
    const int op = P.free_surface == 1 ? (FWD ? OP_YF_FS : OP_YB_FS) : (FWD ? OP_YF : OP_YB);
    const int row = (P.free_surface == 1 && gy < H) ? max(gy, 0) : H;
    const float *t = P.tab + ((size_t)op * (2 * H + 1) + row) * (Q + 1) + (FWD ? 1 : 0);
#pragma unroll
    for (int j = 0; j < Q; j++)
        w[j] = __ldg(t + j);
}

// ---------------------------------------------------------------------------------------------------------------------
// shared-memory stage layouts (offsets in floats)
// ---------------------------------------------------------------------------------------------------------------------
template <int Q> struct StageV { // velocity half-step
    using C = Cfg<Q>;
    static constexpr int SXX = 0, SXY = SXX + C::N_X, SXZ = SXY + C::N_X, SYZ = SXZ + C::N_XZ, SZZ = SYZ + C::N_Z;
    static constexpr int FEED = SZZ + C::N_Z;      // 3 plain tiles: Sxy(y+H-1), Syy(y+H), Syz(y+H-1)
    static constexpr int OWNV = FEED + 3 * C::N_P; // vx vy vz
    static constexpr int OWNR = OWNV + 3 * C::N_P; // rix riy riz
    static constexpr int SIZE = OWNR + 3 * C::N_P;
    static constexpr uint32_t BYTES_FEED = 3u * C::N_P * 4u;
    static constexpr uint32_t BYTES_FULL = (uint32_t)SIZE * 4u;
};
template <int Q> struct StageS { // stress half-step
    using C = Cfg<Q>;
    static constexpr int TV = 0;                    // 3 XZ tiles: vx vy vz
    static constexpr int FEED = TV + 3 * C::N_XZ;   // 3 plain tiles: vx(y+H), vy(y+H-1), vz(y+H)
    static constexpr int OWNS = FEED + 3 * C::N_P;  // sxx syy szz sxy sxz syz
    static constexpr int OWNM = OWNS + 6 * C::N_P;  // pi mu muxy muxz muyz
    static constexpr int SIZE = OWNM + 5 * C::N_P;
    static constexpr uint32_t BYTES_FEED = 3u * C::N_P * 4u;
    static constexpr uint32_t BYTES_FULL = (uint32_t)SIZE * 4u;
};

struct Bars {
    uint64_t full[NST], empty[NST];
};

// ---------------------------------------------------------------------------------------------------------------------
// velocity half-step (ForwardSolver3Delastic.cpp:181-277)
//   group 0: vx += rix * (Dxf Sxx + Dyb* Sxy + Dzb Sxz)      group 1: vy += riy * (Dxb Sxy + Dyf* Syy + Dzb Syz)
//   group 2: vz += riz * (Dxb Sxz + Dyb* Syz + Dzf Szz)
// ---------------------------------------------------------------------------------------------------------------------
template <int Q, bool CPML, int G>
__device__ __forceinline__ void velConsumer(const WsParams &P, const float *sm, Bars *bars, int tg, int tx0, int tz0, int yc0, int yc1)
{
    using C = Cfg<Q>;
    using S = StageV<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH;
    constexpr bool YFWD = (G == 1);
    const int lx = tg % C::LXN, lz = tg / C::LXN;
    const int x0 = tx0 + 4 * lx, z = tz0 + lz;
    const bool active = (x0 < P.nx) && (z < P.nz);
    const int lane = threadIdx.x & 31;

    float c[Q], wy[Q];
    loadInterior<Q>(P, c);
    loadYWeights<Q, YFWD>(P, H, wy);
    const long long rowOff = P.base + x0 + (long long)z * P.pitch;
    float *gv = P.fld[F_VX + G] + rowOff;

    constexpr int sx = (G == 0) ? PSI_SXX_X : (G == 1 ? PSI_SXY_X : PSI_SXZ_X);
    constexpr int sy = (G == 0) ? PSI_SXY_Y : (G == 1 ? PSI_SYY_Y : PSI_SYZ_Y);
    constexpr int sz = (G == 0) ? PSI_SXZ_Z : (G == 1 ? PSI_SYZ_Z : PSI_SZZ_Z);
    CpT cpt;
    cpSetup<CPML>(P, cpt, active, x0, z, /*halfX*/ G == 0, /*halfZ*/ G == 2);

    // shared-memory offsets of this thread inside a stage
    const int oX = (G == 0 ? S::SXX : (G == 1 ? S::SXY : S::SXZ + H * TXH)) + lz * TXH + 4 * lx; // column of x0 - HX
    const int oZ = (G == 0 ? S::SXZ + lz * TXH + 4 * lx + HX : (G == 1 ? S::SYZ : S::SZZ) + lz * TX + 4 * lx);
    constexpr int ldZ = (G == 0) ? TXH : TX;
    const int oP = lz * TX + 4 * lx;

    F4 q[Q];
#pragma unroll
    for (int k = 0; k < Q; k++)
        q[k] = zero4();

    const int nIter = (Q - 1) + (yc1 - yc0);
    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < nIter; it++) {
        const int ly = yc0 - (Q - 1) + it;
        const int gy = P.gy0 + ly;
        const bool comp = it >= Q - 1;
        CpI cpi;
        cpi.ky = -1;
        if (comp) {
            cpLoad<CPML, sx, sy, sz>(P, cpt, cpi, ly, gy, x0, z, /*halfY*/ G == 1);
            if (P.free_surface == 1 && gy <= H)
                loadYWeights<Q, YFWD>(P, gy, wy);
        }
        mbarWait(&bars->full[stage], parity);
        const float *st = sm + stage * S::SIZE;
        q[Q - 1] = ld4(st + S::FEED + G * C::N_P + oP);
        if (comp) {
            F4 u = dX<Q, G == 0>(st + oX, c);
            cpApplyX<CPML>(P, cpt, cpi, u, sx);
            F4 w = dY<Q>(q, wy);
            cpApplyY<CPML>(P, cpt, cpi, w, sy);
#pragma unroll
            for (int p = 0; p < 4; p++)
                u.v[p] = A::add(u.v[p], w.v[p]);
            w = dZ<Q, G == 2>(st + oZ, ldZ, c);
            cpApplyZ<CPML>(P, cpt, cpi, w, sz);
            F4 v = ld4(st + S::OWNV + G * C::N_P + oP);
            const F4 r = ld4(st + S::OWNR + G * C::N_P + oP);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                u.v[p] = A::add(u.v[p], w.v[p]);
                u.v[p] = A::mul(u.v[p], r.v[p]);
                v.v[p] = A::add(v.v[p], u.v[p]);
            }
            if (active)
                st4cs(gv + (long long)ly * P.plane, v);
        }
        __syncwarp();
        if (lane == 0)
            mbarArrive(&bars->empty[stage]);
#pragma unroll
        for (int k = 0; k < Q - 1; k++)
            q[k] = q[k + 1];
        if (++stage == NST) {
            stage = 0;
            parity ^= 1;
        }
    }
}

template <int Q, bool CPML> __global__ void __launch_bounds__(NG_VEL *Cfg<Q>::NTG + 32, 1) kFastVel(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = StageV<Q>;
    constexpr int H = C::H, HX = C::HX;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) Bars bars;
    float *sm = reinterpret_cast<float *>(smraw);
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TX, tz0 = blockIdx.y * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) {
            mbarInit(&bars.full[s], 1);
            mbarInit(&bars.empty[s], NG_VEL * C::WPG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int grp = tid / C::NTG;
    if (grp == 0)
        velConsumer<Q, CPML, 0>(P, sm, &bars, tid, tx0, tz0, yc0, yc1);
    else if (grp == 1)
        velConsumer<Q, CPML, 1>(P, sm, &bars, tid - C::NTG, tx0, tz0, yc0, yc1);
    else if (grp == 2)
        velConsumer<Q, CPML, 2>(P, sm, &bars, tid - 2 * C::NTG, tx0, tz0, yc0, yc1);
    else if (tid == NG_VEL * C::NTG) {
        // ---- producer: one elected thread streams the planes ----
        const int HZP = (P.nzp > 1) ? WS_HALO : 0;
        const int cx = WS_PADX + tx0, cz = HZP + tz0;
        const int nIter = (Q - 1) + (yc1 - yc0);
        int stage = 0;
        uint32_t parity = 1; // first pass over the ring: the stages are free
        for (int it = 0; it < nIter; it++) {
            const int cy = WS_HALO + yc0 - (Q - 1) + it;
            const bool comp = it >= Q - 1;
            if (it >= NST)
                mbarWait(&bars.empty[stage], parity);
            float *st = sm + stage * S::SIZE;
            uint64_t *bar = &bars.full[stage];
            mbarExpectTx(bar, comp ? S::BYTES_FULL : S::BYTES_FEED);
            tmaLoad3D(st + S::FEED, &maps[TM_SXY_P], bar, cx, cz, cy + H - 1);
            tmaLoad3D(st + S::FEED + C::N_P, &maps[TM_SYY_P], bar, cx, cz, cy + H);
            tmaLoad3D(st + S::FEED + 2 * C::N_P, &maps[TM_SYZ_P], bar, cx, cz, cy + H - 1);
            if (comp) {
                tmaLoad3D(st + S::SXX, &maps[TM_SXX_X], bar, cx - HX, cz, cy);
                tmaLoad3D(st + S::SXY, &maps[TM_SXY_X], bar, cx - HX, cz, cy);
                tmaLoad3D(st + S::SXZ, &maps[TM_SXZ_XZ], bar, cx - HX, cz - H, cy);
                tmaLoad3D(st + S::SYZ, &maps[TM_SYZ_Z], bar, cx, cz - H, cy);
                tmaLoad3D(st + S::SZZ, &maps[TM_SZZ_Z], bar, cx, cz - H, cy);
                tmaLoad3D(st + S::OWNV, &maps[TM_VX_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNV + C::N_P, &maps[TM_VY_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNV + 2 * C::N_P, &maps[TM_VZ_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNR, &maps[TM_RIX_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNR + C::N_P, &maps[TM_RIY_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNR + 2 * C::N_P, &maps[TM_RIZ_P], bar, cx, cz, cy);
            }
            if (++stage == NST) {
                stage = 0;
                parity ^= 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stress half-step (ForwardSolver3Delastic.cpp:288-404) incl. free-surface correction
//   group 0: vxx = Dxb vx, vyy = Dyb vy, vzz = Dzb vz -> sxx, syy, szz (+ free surface)
//   group 1: sxy += muxy (Dyf vx + Dxf vy)    group 2: sxz += muxz (Dzf vx + Dxf vz), syz += muyz (Dzf vy + Dyf vz)
// ---------------------------------------------------------------------------------------------------------------------
template <int Q, bool CPML, int G>
__device__ __forceinline__ void strConsumer(const WsParams &P, const float *sm, Bars *bars, int tg, int tx0, int tz0, int yc0, int yc1)
{
    using C = Cfg<Q>;
    using S = StageS<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH, NXZ = C::N_XZ, NP = C::N_P;
    constexpr int FEEDSLOT = (G == 0) ? 1 : (G == 1 ? 0 : 2); // vy | vx | vz
    const int lx = tg % C::LXN, lz = tg / C::LXN;
    const int x0 = tx0 + 4 * lx, z = tz0 + lz;
    const bool active = (x0 < P.nx) && (z < P.nz);
    const int lane = threadIdx.x & 31;

    float c[Q];
    loadInterior<Q>(P, c);
    const long long rowOff = P.base + x0 + (long long)z * P.pitch;

    CpT cpt;
    // group 0 uses the full-grid profiles (vxx, vyy, vzz), the shear groups the half-grid ones (CPML3D.cpp:31-153)
    cpSetup<CPML>(P, cpt, active, x0, z, /*halfX*/ G != 0, /*halfZ*/ G != 0);

    const int oX = (lz + H) * TXH + 4 * lx; // x stencil: own row, column of x0 - HX   (inside an XZ tile)
    const int oZ = lz * TXH + 4 * lx + HX;  // z stencil: row z - H, own column
    const int oP = lz * TX + 4 * lx;

    F4 q[Q];
#pragma unroll
    for (int k = 0; k < Q; k++)
        q[k] = zero4();

    const int nIter = (Q - 1) + (yc1 - yc0);
    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < nIter; it++) {
        const int ly = yc0 - (Q - 1) + it;
        const int gy = P.gy0 + ly;
        const bool comp = it >= Q - 1;
        CpI cpi;
        cpi.ky = -1;
        if (comp) {
            if (G == 0)
                cpLoad<CPML, PSI_VXX, PSI_VYY, PSI_VZZ>(P, cpt, cpi, ly, gy, x0, z, false);
            else if (G == 1)
                cpLoad<CPML, PSI_VYX, PSI_VXY, -1>(P, cpt, cpi, ly, gy, x0, z, true);
            else
                cpLoad<CPML, PSI_VZX, PSI_VZY, PSI_VXZ, PSI_VYZ>(P, cpt, cpi, ly, gy, x0, z, true);
        }
        mbarWait(&bars->full[stage], parity);
        const float *st = sm + stage * S::SIZE;
        const float *tvx = st + S::TV, *tvy = st + S::TV + NXZ, *tvz = st + S::TV + 2 * NXZ;
        q[Q - 1] = ld4(st + S::FEED + FEEDSLOT * NP + oP);
        if (comp) {
            const long long o = rowOff + (long long)ly * P.plane;
            if (G == 0) {
                // normal strain rates; the y derivative of vy is the plain operator even below a free surface (:289)
                F4 vxx = dX<Q, false>(tvx + oX, c);
                F4 vyy = dY<Q>(q, c);
                F4 vzz = dZ<Q, false>(tvz + oZ, TXH, c);
                cpApplyX<CPML>(P, cpt, cpi, vxx, PSI_VXX);
                cpApplyY<CPML>(P, cpt, cpi, vyy, PSI_VYY);
                cpApplyZ<CPML>(P, cpt, cpi, vzz, PSI_VZZ);
                F4 sxx = ld4(st + S::OWNS + 0 * NP + oP), syy = ld4(st + S::OWNS + 1 * NP + oP), szz = ld4(st + S::OWNS + 2 * NP + oP);
                const F4 pi = ld4(st + S::OWNM + 0 * NP + oP), mu = ld4(st + S::OWNM + 1 * NP + oP);
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
                    st4cs(P.fld[F_SXX] + o, sxx);
                    st4cs(P.fld[F_SYY] + o, syy);
                    st4cs(P.fld[F_SZZ] + o, szz);
                }
            } else if (G == 1) {
                F4 u = dY<Q>(q, c);
                cpApplyY<CPML>(P, cpt, cpi, u, PSI_VXY);
                F4 w = dX<Q, true>(tvy + oX, c);
                cpApplyX<CPML>(P, cpt, cpi, w, PSI_VYX);
                F4 s = ld4(st + S::OWNS + 3 * NP + oP);
                const F4 m = ld4(st + S::OWNM + 2 * NP + oP);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float t = A::add(u.v[p], w.v[p]);
                    s.v[p] = A::add(s.v[p], A::mul(t, m.v[p]));
                }
                if (active)
                    st4cs(P.fld[F_SXY] + o, s);
            } else {
                F4 u = dZ<Q, true>(tvx + oZ, TXH, c);
                cpApplyZ<CPML>(P, cpt, cpi, u, PSI_VXZ);
                F4 w = dX<Q, true>(tvz + oX, c);
                cpApplyX<CPML>(P, cpt, cpi, w, PSI_VZX);
                F4 s = ld4(st + S::OWNS + 4 * NP + oP);
                const F4 m = ld4(st + S::OWNM + 3 * NP + oP);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float t = A::add(u.v[p], w.v[p]);
                    s.v[p] = A::add(s.v[p], A::mul(t, m.v[p]));
                }
                if (active)
                    st4cs(P.fld[F_SXZ] + o, s);
                u = dZ<Q, true>(tvy + oZ, TXH, c);
                cpApplyZ2<CPML>(P, cpt, cpi, u, PSI_VYZ);
                w = dY<Q>(q, c);
                cpApplyY<CPML>(P, cpt, cpi, w, PSI_VZY);
                s = ld4(st + S::OWNS + 5 * NP + oP);
                const F4 m2 = ld4(st + S::OWNM + 4 * NP + oP);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float t = A::add(u.v[p], w.v[p]);
                    s.v[p] = A::add(s.v[p], A::mul(t, m2.v[p]));
                }
                if (active)
                    st4cs(P.fld[F_SYZ] + o, s);
            }
        }
        __syncwarp();
        if (lane == 0)
            mbarArrive(&bars->empty[stage]);
#pragma unroll
        for (int k = 0; k < Q - 1; k++)
            q[k] = q[k + 1];
        if (++stage == NST) {
            stage = 0;
            parity ^= 1;
        }
    }
}

template <int Q, bool CPML> __global__ void __launch_bounds__(NG_STR *Cfg<Q>::NTG + 32, 1) kFastStress(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = StageS<Q>;
    constexpr int H = C::H, HX = C::HX;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) Bars bars;
    float *sm = reinterpret_cast<float *>(smraw);
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TX, tz0 = blockIdx.y * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) {
            mbarInit(&bars.full[s], 1);
            mbarInit(&bars.empty[s], NG_STR * C::WPG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int grp = tid / C::NTG;
    if (grp == 0)
        strConsumer<Q, CPML, 0>(P, sm, &bars, tid, tx0, tz0, yc0, yc1);
    else if (grp == 1)
        strConsumer<Q, CPML, 1>(P, sm, &bars, tid - C::NTG, tx0, tz0, yc0, yc1);
    else if (grp == 2)
        strConsumer<Q, CPML, 2>(P, sm, &bars, tid - 2 * C::NTG, tx0, tz0, yc0, yc1);
    else if (tid == NG_STR * C::NTG) {
        const int HZP = (P.nzp > 1) ? WS_HALO : 0;
        const int cx = WS_PADX + tx0, cz = HZP + tz0;
        const int nIter = (Q - 1) + (yc1 - yc0);
        int stage = 0;
        uint32_t parity = 1;
        for (int it = 0; it < nIter; it++) {
            const int cy = WS_HALO + yc0 - (Q - 1) + it;
            const bool comp = it >= Q - 1;
            if (it >= NST)
                mbarWait(&bars.empty[stage], parity);
            float *st = sm + stage * S::SIZE;
            uint64_t *bar = &bars.full[stage];
            mbarExpectTx(bar, comp ? S::BYTES_FULL : S::BYTES_FEED);
            tmaLoad3D(st + S::FEED, &maps[TM_VX_P], bar, cx, cz, cy + H);
            tmaLoad3D(st + S::FEED + C::N_P, &maps[TM_VY_P], bar, cx, cz, cy + H - 1);
            tmaLoad3D(st + S::FEED + 2 * C::N_P, &maps[TM_VZ_P], bar, cx, cz, cy + H);
            if (comp) {
                tmaLoad3D(st + S::TV, &maps[TM_VX_XZ], bar, cx - HX, cz - H, cy);
                tmaLoad3D(st + S::TV + C::N_XZ, &maps[TM_VY_XZ], bar, cx - HX, cz - H, cy);
                tmaLoad3D(st + S::TV + 2 * C::N_XZ, &maps[TM_VZ_XZ], bar, cx - HX, cz - H, cy);
                tmaLoad3D(st + S::OWNS, &maps[TM_SXX_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNS + C::N_P, &maps[TM_SYY_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNS + 2 * C::N_P, &maps[TM_SZZ_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNS + 3 * C::N_P, &maps[TM_SXY_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNS + 4 * C::N_P, &maps[TM_SXZ_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNS + 5 * C::N_P, &maps[TM_SYZ_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNM, &maps[TM_PW_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNM + C::N_P, &maps[TM_MU_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNM + 2 * C::N_P, &maps[TM_MUXY_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNM + 3 * C::N_P, &maps[TM_MUXZ_P], bar, cx, cz, cy);
                tmaLoad3D(st + S::OWNM + 4 * C::N_P, &maps[TM_MUYZ_P], bar, cx, cz, cy);
            }
            if (++stage == NST) {
                stage = 0;
                parity ^= 1;
            }
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
    const int smV = NST * StageV<Q>::SIZE * 4, smS = NST * StageS<Q>::SIZE * 4;
    cudaFuncSetAttribute(kFastVel<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smV);
    cudaFuncSetAttribute(kFastVel<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smV);
    cudaFuncSetAttribute(kFastStress<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS);
    cudaFuncSetAttribute(kFastStress<Q, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS);
    done = true;
}

template <int Q> void launchQ(const WsParams &P, int pass, cudaStream_t st)
{
    setAttrs<Q>();
    const int ny = P.yhi - P.ylo;
    dim3 grid((P.nx + TX - 1) / TX, (P.nz + TZ - 1) / TZ, (ny + P.fastChunk - 1) / P.fastChunk);
    const bool cpml = P.damping == 2;
    if (pass == 0) {
        const size_t sm = (size_t)NST * StageV<Q>::SIZE * 4;
        const int nt = NG_VEL * Cfg<Q>::NTG + 32;
        if (cpml)
            kFastVel<Q, true><<<grid, nt, sm, st>>>(P);
        else
            kFastVel<Q, false><<<grid, nt, sm, st>>>(P);
    } else {
        const size_t sm = (size_t)NST * StageS<Q>::SIZE * 4;
        const int nt = NG_STR * Cfg<Q>::NTG + 32;
        if (cpml)
            kFastStress<Q, true><<<grid, nt, sm, st>>>(P);
        else
            kFastStress<Q, false><<<grid, nt, sm, st>>>(P);
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
    mk(TM_SXY_P, P.fld[F_SXY], TX, TZ);
    mk(TM_SYY_P, P.fld[F_SYY], TX, TZ);
    mk(TM_SYZ_P, P.fld[F_SYZ], TX, TZ);
    mk(TM_VX_P, P.fld[F_VX], TX, TZ);
    mk(TM_VY_P, P.fld[F_VY], TX, TZ);
    mk(TM_VZ_P, P.fld[F_VZ], TX, TZ);
    mk(TM_RIX_P, P.mat[M_RIX], TX, TZ);
    mk(TM_RIY_P, P.mat[M_RIY], TX, TZ);
    mk(TM_RIZ_P, P.mat[M_RIZ], TX, TZ);
    mk(TM_VX_XZ, P.fld[F_VX], TXH, TZH);
    mk(TM_VY_XZ, P.fld[F_VY], TXH, TZH);
    mk(TM_VZ_XZ, P.fld[F_VZ], TXH, TZH);
    mk(TM_SXX_P, P.fld[F_SXX], TX, TZ);
    mk(TM_SZZ_P, P.fld[F_SZZ], TX, TZ);
    mk(TM_SXZ_P, P.fld[F_SXZ], TX, TZ);
    mk(TM_PW_P, P.mat[M_PW], TX, TZ);
    mk(TM_MU_P, P.mat[M_MU], TX, TZ);
    mk(TM_MUXY_P, P.mat[M_MUXY], TX, TZ);
    mk(TM_MUXZ_P, P.mat[M_MUXZ], TX, TZ);
    mk(TM_MUYZ_P, P.mat[M_MUYZ], TX, TZ);
    void *dev = nullptr;
    if (cudaMalloc(&dev, sizeof(CUtensorMap) * TM_COUNT) != cudaSuccess)
        throw std::runtime_error("cudaMalloc for tensor maps failed");
    cudaMemcpy(dev, maps.data(), sizeof(CUtensorMap) * TM_COUNT, cudaMemcpyHostToDevice);
    P.fastMaps = dev;
    // planes per block: enough blocks to fill the 148 SMs several times over, long enough marches to amortise the
    // Q-1 feed-only iterations of the prologue
    const int tiles = ((P.nx + TX - 1) / TX) * ((P.nz + TZ - 1) / TZ);
    int chunks = (6 * 148 + tiles - 1) / tiles;
    if (chunks < 1)
        chunks = 1;
    int chunk = (P.nyl + chunks - 1) / chunks;
    if (chunk < 32)
        chunk = 32;
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

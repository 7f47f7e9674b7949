// ws_kernels_fast.cu — warp-specialised sm_100a kernels for the 3-D elastic hot path (FD3Delastic::run,
// ForwardSolver/ForwardSolver3Delastic.cpp:120-414): one fused kernel per half-step.
//
// Structure (2.5-D blocking, producer / consumer pipeline):
//   * a thread block owns a TX x TZ tile of the x-z plane and marches along y (the slowest axis);
//   * ONE PRODUCER WARP streams, per plane, every operand of that plane into a ring of NST shared-memory stages with TMA
//     (cp.async.bulk.tensor.4d over field arenas, completion on a "full" mbarrier per stage): halo tiles for the x / z
//     stencils, the planes that enter the y-derivative windows, and the own-point operands (updated fields, model
//     parameters).  Consumer threads never wait on DRAM: up to NST-1 planes (~100 KB per SM) are in flight while one
//     is being computed;
//   * CONSUMER GROUPS split the work by OUTPUT COMPONENT (velocity half-step: vx | vy | vz; stress half-step:
//     sxx,syy,szz | sxy | sxz,syz).  Every thread owns 4 consecutive x points (128-bit shared / global accesses) and
//     keeps only the ONE field it differentiates along y in a register queue (Q planes deep);
//   * the march is unrolled Q times with a rotating queue index (no register moves), with the FD weights as
//     constant-bank operands; planes that need per-plane treatment (image-method rows below the free surface, the
//     y-CPML layers, the queue prologue, the remainder) take a generic step with run-time weights;
//   * a stage is handed back to the producer through an "empty" mbarrier (one arrival per consumer warp);
//   * off-grid taps read the zero pads of the HBM layout (StencilMatrix "drop off-grid taps", edge_policy 0);
//   * CPML memory variables live in compact boundary slabs; their loads are issued before the stage wait.
// The arithmetic sequence is the one of the general kernels (ws_kernels_general.cuh) in FMA mode, so both produce
// bit-identical results.
#include "ws_fast_common.cuh"

namespace {

#ifndef WS_NSTV
#define WS_NSTV 4
#endif
#ifndef WS_NSTS
#define WS_NSTS 4
#endif
constexpr int NSTV = WS_NSTV, NSTS = WS_NSTS; // ring depth of the velocity / stress kernel
// tensor-map slots.  Arena order of the wavefields: vx vy vz sxx sxy syy syz szz sxz; of the model parameters:
// rix riy riz pi mu muxy muxz muyz (ws_api.cu) — arrays that one box fetches together are neighbours.
enum {
    TM_V_P = 0,   // {vx,vy,vz} plain            (velocity: own operands)
    TM_R_P,       // {rix,riy,riz} plain
    TM_SX_X,      // {sxx,sxy} with x halo
    TM_SXZ_XZ,    // sxz with x and z halo
    TM_SZ_XZ,     // {syz,szz} with x and z halo
    TM_F1_P,      // one field, plain: planes entering the y windows (4th coordinate selects the field)
    TM_V_XZ,      // {vx,vy,vz} with x and z halo (stress half-step)
    TM_S_P,       // {sxx,sxy,syy,syz,szz,sxz} plain (stress: own operands)
    TM_M_P,       // {pi,mu,muxy,muxz,muyz} plain
    TM_PX,        // CPML memory variables of the x terms: slab rows (2W wide) of the tile's z rows, 3 arrays
    TM_PZ,        // CPML memory variables of the z terms: slab rows of the tile, 3 arrays
    TM_COUNT
};
// positions inside the arenas
enum { AF_VX = 0, AF_VY, AF_VZ, AF_SXX, AF_SXY, AF_SYY, AF_SYZ, AF_SZZ, AF_SXZ, AF_COUNT };
enum { AM_RIX = 0, AM_RIY, AM_RIZ, AM_PW, AM_MU, AM_MUXY, AM_MUXZ, AM_MUYZ, AM_COUNT };
// memory-variable arenas (ws_api.cu): x terms {sxx_x, sxy_x, sxz_x | vxx, vyx, vzx}, z terms {sxz_z, syz_z, szz_z | vzz, vxz, vyz};
// the first three of each belong to the velocity half-step (role order), the last three to the stress half-step
enum { APS_COUNT = 6 };

// ---------------------------------------------------------------------------------------------------------------------
// shared-memory stage layouts (offsets in floats); the order inside a multi-field box is the arena order
// ---------------------------------------------------------------------------------------------------------------------
template <int Q> struct StageV { // velocity half-step
    using C = Cfg<Q>;
    static constexpr int SXX = 0, SXY = SXX + C::N_X;               // TM_SX_X   {sxx,sxy}, x halo
    static constexpr int SYZ = SXY + C::N_X, SZZ = SYZ + C::N_XZ;   // TM_SZ_XZ  {syz,szz}, x and z halo
    static constexpr int SXZ = SZZ + C::N_XZ;                       // TM_SXZ_XZ sxz, x and z halo
    static constexpr int FEED = SXZ + C::N_XZ;                      // 3 plain tiles: Sxy(y+H-1), Syy(y+H), Syz(y+H-1)
    static constexpr int OWNV = FEED + 3 * C::N_P;                  // vx vy vz
    static constexpr int OWNR = OWNV + 3 * C::N_P;                  // rix riy riz
    static constexpr int SIZE = OWNR + 3 * C::N_P;
    static constexpr uint32_t BYTES_FEED = 3u * C::N_P * 4u;
    static constexpr uint32_t BYTES_FULL = (uint32_t)SIZE * 4u;
};
template <int Q> struct StageS { // stress half-step
    using C = Cfg<Q>;
    static constexpr int TV = 0;                   // 3 XZ tiles: vx vy vz
    static constexpr int FEED = TV + 3 * C::N_XZ;  // 3 plain tiles: vx(y+H), vy(y+H-1), vz(y+H)
    static constexpr int OWNS = FEED + 3 * C::N_P; // sxx sxy syy syz szz sxz
    static constexpr int OWNM = OWNS + 6 * C::N_P; // pi mu muxy muxz muyz
    static constexpr int SIZE = OWNM + 5 * C::N_P;
    static constexpr uint32_t BYTES_FEED = 3u * C::N_P * 4u;
    static constexpr uint32_t BYTES_FULL = (uint32_t)SIZE * 4u;
};

// floats of the staged memory variables per stage: 3 x-term slab rows sets (TZ rows of 2W) + 3 z-term tiles
// (PX = row length of the x-term slabs, 2W rounded up to a 16-byte multiple)
__host__ __device__ __forceinline__ int psxFloats(int PX) { return (3 * TZ * PX + 31) / 32 * 32; }
__host__ __device__ __forceinline__ int psiStageFloats(int PX) { return psxFloats(PX) + 3 * TX * TZ; }
// ---------------------------------------------------------------------------------------------------------------------
// velocity half-step (ForwardSolver3Delastic.cpp:181-277)
//   role 0: vx += rix * (Dxf Sxx + Dyb* Sxy + Dzb Sxz)      role 1: vy += riy * (Dxb Sxy + Dyf* Syy + Dzb Syz)
//   role 2: vz += riz * (Dxb Sxz + Dyb* Syz + Dzf Szz)
// The three roles have the same shape (one x, one y, one z derivative), so they share ONE instruction stream whose
// operands are run-time offsets: three role-specific code streams of this size evict each other from the instruction
// cache (measured).  The forward / backward choice of the x operator is folded into a 9-tap weight vector with a
// leading or trailing zero (adding 0*w leaves every partial sum unchanged), of the z operator into the start row, of
// the y operator into the plane the producer feeds.
// ---------------------------------------------------------------------------------------------------------------------
struct VRole {
    int oX, oZ, oF, oV, oR; // shared-memory offsets inside a stage
    int oZ9;                // z stencil from row z - H whatever the operator (order-reducing edge rows)
    int opX;                // x operator of this role (table rows of the edge columns)
    bool xEdge;             // order-reducing edges: some of the thread's 4 columns lie within q/2 of an x face
    const float *wz;        // ... the thread's row lies within q/2 of a z face: its weights (null otherwise)
    float cx[WS_MAXQ + 1];  // x weights over offsets -H..+H
    float *out;             // own row of the output field at plane 0
    float *psx, *psy, *psz; // memory-variable slabs of the x / y / z term
    int opY;                // OP_YF or OP_YB (+ image-method variant) for the generic step
    bool yFwd, halfY;
};

// one plane of one velocity component.  R = position inside a trip; GENERIC = run-time y weights / y-CPML; XZ = this
// warp may sit in an x or z CPML layer.
// POL1 = order-reducing edges (edge_policy 1): own instantiations, so that the code of the default policy is not touched
template <int Q, int R, bool GENERIC, bool XZ, bool POL1>
__device__ __forceinline__ void velPlane(const WsParams &P, const Thr &t, const CpT &cpt, const VRole &ro, F4 (&q)[Cfg<Q>::QL], const float *st, float *gout, float *psx,
                                         float *psz, const YDyn<Q> &yd, float *psy)
{
    using C = Cfg<Q>;
    const bool ycp = GENERIC && yd.ky >= 0 && t.active;
    F4 pz, py;
    if (XZ && cpt.kz >= 0) // staged by the producer together with the operands of this plane
        pz = ld4(st + t.oPZ);
    if (ycp)
        py = ld4(psy);
    F4 u = (POL1 && XZ && ro.xEdge) ? dX9t<Q>(st + ro.oX, P.tab, ro.opX, t.x0, P.nx) : dX9<Q>(st + ro.oX, ro.cx);
    if (XZ && cpt.kxv >= 0)
        cpApplyXv(cpt, st + t.oPX, psx, t.cxTab, P.psiPitchX, u);
    F4 w = dY<Q, R>(q, GENERIC ? yd.w : P.cwy);
    if (ycp)
        cpApply4(psy, py, yd.ya, yd.yb, w);
#pragma unroll
    for (int p = 0; p < 4; p++)
        u.v[p] = A::add(u.v[p], w.v[p]);
    w = (POL1 && XZ && ro.wz) ? dZ9t<Q, C::TXH>(st + ro.oZ9, ro.wz) : dZ<Q, false, C::TXH>(st + ro.oZ, P.cw);
    if (XZ && cpt.kz >= 0)
        cpApply4(psz, pz, cpt.za, cpt.zb, w);
    F4 v = ld4(st + ro.oV);
    const F4 r = ld4(st + ro.oR);
#pragma unroll
    for (int p = 0; p < 4; p++) {
        u.v[p] = A::add(u.v[p], w.v[p]);
        u.v[p] = A::mul(u.v[p], r.v[p]);
        v.v[p] = A::add(v.v[p], u.v[p]);
    }
    if (t.active && P.fastDebug != 2) {
        if (P.fastFlags & 4)
            st4(gout, v);
        else if (P.fastFlags & 8)
            __stcg(reinterpret_cast<float4 *>(gout), make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
        else
            st4cs(gout, v);
    }
}

// UNRV consecutive planes of a trip (template recursion keeps the queue offset R a compile-time constant)
template <int Q, bool XZ, bool POL1, int R> struct VelUnroll {
    static __device__ __forceinline__ void run(const WsParams &P, const Thr &t, const CpT &cpt, const VRole &ro, F4 (&q)[Cfg<Q>::QL], const float *sm, int stage0,
                                               uint32_t parity, float *gout, float *psx, float *psz, long long sxStride, long long szStride)
    {
        using S = StageV<Q>;
        const int stage = stage0 + R;
        mbarWait(t.barFull + 8u * stage, parity);
        const float *st = sm + stage * t.stride;
        q[Q - 1 + R] = ld4(st + ro.oF);
        YDyn<Q> yd;
        yd.ky = -1;
        if (P.fastDebug == 3) { // stores without arithmetic
            if (t.active)
                st4cs(gout, ld4(st + ro.oV));
        } else if (P.fastDebug != 1)
            velPlane<Q, R, false, XZ, POL1>(P, t, cpt, ro, q, st, gout, psx, psz, yd, nullptr);
        consumerRelease(t, stage);
        VelUnroll<Q, XZ, POL1, R + 1>::run(P, t, cpt, ro, q, sm, stage0, parity, gout + P.plane, XZ ? psx + sxStride : psx, XZ ? psz + szStride : psz, sxStride, szStride);
    }
};
template <int Q, bool XZ, bool POL1> struct VelUnroll<Q, XZ, POL1, UNRV> {
    static __device__ __forceinline__ void run(const WsParams &, const Thr &, const CpT &, const VRole &, F4 (&)[Cfg<Q>::QL], const float *, int, uint32_t, float *,
                                               float *, float *, long long, long long)
    {
    }
};

template <int Q, bool CPML, bool EDGE, bool POL1>
__device__ __forceinline__ void velConsumer(const WsParams &P, const float *sm, Thr t, int G, int yc0, int yc1)
{
    using S = StageV<Q>;
    using C = Cfg<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH;
    constexpr bool XZC = CPML && EDGE; // tiles of this launch may touch an x / z CPML layer
    // ---- role set-up (run-time; CPML slots and profiles: CPML3D.cpp:31-153) ----
    VRole ro;
    const int oP = t.lz * TX + 4 * t.lx;
    const int oXrow = t.lz * TXH + 4 * t.lx; // column of x0 - HX in a tile without z halo
    ro.oX = G == 0 ? S::SXX + oXrow : (G == 1 ? S::SXY + oXrow : S::SXZ + H * TXH + oXrow);
    // z stencil: row z - H (backward) or z - H + 1 (forward, role 2), own column of a tile with x and z halo
    ro.oZ = (G == 0 ? S::SXZ : (G == 1 ? S::SYZ : S::SZZ + TXH)) + oXrow + HX;
    ro.oF = S::FEED + G * C::N_P + oP;
    ro.oV = S::OWNV + G * C::N_P + oP;
    ro.oR = S::OWNR + G * C::N_P + oP;
    // order-reducing edges: operators Dxf | Dxb | Dxb and Dzb | Dzb | Dzf of the three roles
    ro.oZ9 = G == 2 ? ro.oZ - TXH : ro.oZ;
    ro.opX = G == 0 ? OP_XF : OP_XB;
    ro.xEdge = POL1 && XZC && t.active && (t.x0 < H || t.x0 + 3 >= P.nx - H);
    ro.wz = (POL1 && XZC && t.active && (t.z < H || t.z >= P.nz - H))
                ? P.tab + ((size_t)(G == 2 ? OP_ZF : OP_ZB) * (2 * H + 1) + wsRowClass(t.z, P.nz, H)) * (Q + 1)
                : nullptr;
#pragma unroll
    for (int j = 0; j <= Q; j++) // forward (role 0): offsets -H+1..H; backward: -H..H-1
        ro.cx[j] = G == 0 ? (j >= 1 ? P.cw[j - 1] : 0.0f) : (j < Q ? P.cw[j] : 0.0f);
    ro.yFwd = G == 1;
    ro.halfY = G == 1;
    ro.opY = P.free_surface == 1 ? (ro.yFwd ? OP_YF_FS : OP_YB_FS) : (ro.yFwd ? OP_YF : OP_YB);
    const int sx = G == 0 ? PSI_SXX_X : (G == 1 ? PSI_SXY_X : PSI_SXZ_X);
    const int sy = G == 0 ? PSI_SXY_Y : (G == 1 ? PSI_SYY_Y : PSI_SYZ_Y);
    const int sz = G == 0 ? PSI_SXZ_Z : (G == 1 ? PSI_SYZ_Z : PSI_SZZ_Z);
    CpT cpt;
    cpSetup<XZC>(P, cpt, t.active, t.x0, t.z, /*halfX*/ G == 0, /*halfZ*/ G == 2);
    // warp-uniform: the unrolled march contains warp-level synchronisation
    const bool warpXZ = XZC && P.fastDebug != 4 && __any_sync(0xffffffffu, cpt.kxv >= 0 || cpt.kz >= 0);
    const int W2 = 2 * P.W, PX = P.psiPitchX;
    // strides of the memory-variable slabs per plane
    const long long sxStride = (long long)P.nz * PX, szStride = (long long)W2 * P.nx;
    ro.out = P.fld[F_VX + G] + P.base + t.x0 + (long long)t.z * P.pitch;
    ro.psx = XZC ? P.psi[sx] + (long long)t.z * PX : nullptr;
    ro.psz = XZC ? P.psi[sz] + (long long)cpt.kz * P.nx + t.x0 : nullptr;
    ro.psy = CPML ? P.psi[sy] + (long long)t.z * P.nx + t.x0 : nullptr;
    // staged x rows: the box holds psiBoxX entries from entry xs of the row (the tile's side of the layer)
    const int xs = (P.psiBoxX == PX || t.x0 - 4 * t.lx < P.W) ? 0 : PX - P.psiBoxX;
    t.oPX = S::SIZE + (G * TZ + t.lz) * P.psiBoxX - xs;
    t.oPZ = S::SIZE + psxFloats(P.psiBoxX) + G * C::N_P + oP;

    F4 q[C::QL];
#pragma unroll
    for (int k = 0; k < C::QL; k++)
        q[k] = zero4();

    const int nIter = (Q - 1) + (yc1 - yc0);
    // generic step: queue in canonical order, shifted by register moves
    auto generic = [&](int it) {
        const int ly = yc0 - (Q - 1) + it, gy = P.gy0 + ly;
        const bool comp = it >= Q - 1;
        const int stage = it % NSTV;
        YDyn<Q> yd;
        yd.ky = -1;
        if (comp) {
            // y weights of this plane: with a free surface every row comes from the image-method operators (their
            // interior rows are scaled (c/DH)*DT, not c*(DT/DH): Derivatives.cpp:407-425 vs FDTD3D.cpp:211-216)
            // (order-reducing edges: the plain operators have their own rows next to both y faces as well)
            const int row = (P.free_surface == 1 && gy < H) ? max(gy, 0) : (POL1 ? wsRowClass(min(max(gy, 0), P.gny - 1), P.gny, H) : H);
            const float *tw = P.tab + ((size_t)ro.opY * (2 * H + 1) + row) * (Q + 1) + (ro.yFwd ? 1 : 0);
#pragma unroll
            for (int j = 0; j < Q; j++)
                yd.w[j] = __ldg(tw + j);
            if (CPML) {
                yd.ky = yCpmlIndex(P, gy);
                if (yd.ky >= 0) {
                    yd.ya = __ldg((ro.halfY ? P.cayh : P.cay) + yd.ky);
                    yd.yb = __ldg((ro.halfY ? P.cbyh : P.cby) + yd.ky);
                }
            }
        }
        mbarWait(t.barFull + 8u * stage, (it / NSTV) & 1);
        const float *st = sm + stage * t.stride;
        q[Q - 1] = ld4(st + ro.oF);
        if (comp)
            velPlane<Q, 0, true, XZC, POL1>(P, t, cpt, ro, q, st, ro.out + (long long)ly * P.plane, XZC ? ro.psx + (long long)ly * sxStride : nullptr,
                                      XZC ? ro.psz + (long long)ly * szStride : nullptr, yd, CPML ? ro.psy + (long long)yd.ky * P.nz * P.nx : nullptr);
        consumerRelease(t, stage);
#pragma unroll
        for (int k = 0; k < Q - 1; k++)
            q[k] = q[k + 1];
    };

    int it = 0;
    while (it < nIter) {
        const int ly0 = yc0 - (Q - 1) + it, gy0 = P.gy0 + ly0;
        // the first Q iterations (queue prologue + first plane) are generic; afterwards aligned trips of UNRV planes
        if (it >= Q && it % UNRV == 0 && it + UNRV <= nIter && !needsGeneric<CPML>(P, gy0, gy0 + UNRV - 1, H, true)) {
            float *gout = ro.out + (long long)ly0 * P.plane;
            const int stage0 = it % NSTV;
            const uint32_t parity = (it / NSTV) & 1;
            if (warpXZ)
                VelUnroll<Q, XZC, POL1, 0>::run(P, t, cpt, ro, q, sm, stage0, parity, gout, ro.psx + (long long)ly0 * sxStride, ro.psz + (long long)ly0 * szStride, sxStride,
                                           szStride);
            else
                VelUnroll<Q, false, POL1, 0>::run(P, t, cpt, ro, q, sm, stage0, parity, gout, nullptr, nullptr, 0, 0);
#pragma unroll
            for (int k = 0; k < Q - 1; k++)
                q[k] = q[k + UNRV];
            it += UNRV;
        } else {
            generic(it);
            it++;
        }
    }
}

template <int Q, bool CPML, bool EDGE, bool POL1 = false> __global__ void __launch_bounds__(NGROUPS *Cfg<Q>::NTG + WS_FAST_PRODUCER_THREADS, 1) kFastVel(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = StageV<Q>;
    constexpr int H = C::H, HX = C::HX;
    constexpr bool XZC = CPML && EDGE;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) Bars bars;
    __shared__ __align__(16) float cxTab[XZC ? 4 * PXMAX : 4];
    float *sm = reinterpret_cast<float *>(smraw);
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);
    if (XZC)
        for (int k = threadIdx.x; k < 4 * P.psiPitchX; k += blockDim.x)
            cxTab[k] = __ldg(P.cxTab + k);

    const int tid = threadIdx.x;
    const int tile = P.fastTiles[P.fastTileBase + blockIdx.x]; // (z tile << 16) | x tile, see wsFastPrepare
    const int tx0 = (tile & 0xffff) * TX, tz0 = (tile >> 16) * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    traceStart(P, 0);
    const uint32_t barFull = smemU32(&bars.full[0]), barEmpty = smemU32(&bars.empty[0]);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTV; s++) {
            mbarInit(barFull + 8u * s, 1);
            mbarInit(barEmpty + 8u * s, NGROUPS * C::WPG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int grp = tid / C::NTG;
    if (grp < NGROUPS) {
        wsConsumerRegs();
        Thr t;
        const int tg = tid - grp * C::NTG;
        t.lx = tg % C::LXN;
        t.lz = tg / C::LXN;
        t.x0 = tx0 + 4 * t.lx;
        t.z = tz0 + t.lz;
        t.lane = tid & 31;
        t.active = (t.x0 < P.nx) && (t.z < P.nz);
        t.barFull = barFull;
        t.barEmpty = barEmpty;
        t.barXFull = t.barXFree = 0;
        t.stride = S::SIZE + (XZC ? psiStageFloats(P.psiBoxX) : 0);
        t.cxTab = cxTab;
        t.oPX = t.oPZ = t.oPZ2 = 0;
        velConsumer<Q, CPML, EDGE, POL1>(P, sm, t, grp, yc0, yc1);
        traceEnd(P, 0, t.lane);
    } else if (wsProducerRegs(), tid == NGROUPS * C::NTG) {
        // ---- producer: one elected thread streams the planes ----
        const int HZP = (P.nzp > 1) ? WS_HALO : 0;
        const int cx = WS_PADX + tx0, cz = HZP + tz0;
        const int nIter = (Q - 1) + (yc1 - yc0);
        const uint32_t smBase = smemU32(sm);
        int stage = 0;
        uint32_t parity = 1; // first pass over the ring: the stages are free
        const bool hints = (P.fastFlags & 1) != 0;
        const uint64_t polOnce = policyEvictFirst(), polKeep = policyEvictLast();
        const int stride = S::SIZE + (XZC ? psiStageFloats(P.psiBoxX) : 0);
        // memory variables of the x / z CPML layers this tile touches (slab rows, ws_api.cu)
        const bool tileX = XZC && P.fastDebug != 4 && (tx0 < P.W || tx0 + TX > P.nx - P.W), tileZ = XZC && P.fastDebug != 4 && (tz0 < P.W || tz0 + TZ > P.nz - P.W);
        const int kz0 = tz0 < P.W ? tz0 : tz0 - (P.nz - 2 * P.W);
        const uint32_t bytesPX = 3u * TZ * (uint32_t)P.psiBoxX * 4u, bytesPZ = 3u * C::N_P * 4u;
        const int xs = (P.psiBoxX == P.psiPitchX || tx0 < P.W) ? 0 : P.psiPitchX - P.psiBoxX;
        const uint32_t bytesFull = S::BYTES_FULL + (tileX ? bytesPX : 0u) + (tileZ ? bytesPZ : 0u);
        for (int it = 0; it < nIter; it++) {
            const int cy = WS_HALO + yc0 - (Q - 1) + it;
            const bool comp = it >= Q - 1;
            if (it >= NSTV)
                mbarWait(barEmpty + 8u * stage, parity);
            const uint32_t st = smBase + (uint32_t)(stage * stride) * 4u;
            const uint32_t bar = barFull + 8u * stage;
            mbarExpectTx(bar, comp ? bytesFull : S::BYTES_FEED);
            if (comp && tileX)
                tmaLoad4D(st + 4u * S::SIZE, &maps[TM_PX], bar, xs, tz0, cy - WS_HALO, 0);
            if (comp && tileZ)
                tmaLoad4D(st + 4u * (S::SIZE + psxFloats(P.psiBoxX)), &maps[TM_PZ], bar, tx0, kz0, cy - WS_HALO, 0);
            if (hints) {
                // Sxy and Syz come back H-1 planes later as stencil tiles; Syy is used by this feed only
                tmaLoad4DHint(st + 4u * S::FEED, &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_SXY, polKeep);
                tmaLoad4DHint(st + 4u * (S::FEED + C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H, AF_SYY, polOnce);
                tmaLoad4DHint(st + 4u * (S::FEED + 2 * C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_SYZ, polKeep);
            } else {
                tmaLoad4D(st + 4u * S::FEED, &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_SXY);
                tmaLoad4D(st + 4u * (S::FEED + C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H, AF_SYY);
                tmaLoad4D(st + 4u * (S::FEED + 2 * C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_SYZ);
            }
            if (comp) {
                tmaLoad4D(st + 4u * S::SXX, &maps[TM_SX_X], bar, cx - HX, cz, cy, AF_SXX);
                tmaLoad4D(st + 4u * S::SYZ, &maps[TM_SZ_XZ], bar, cx - HX, cz - H, cy, AF_SYZ);
                tmaLoad4D(st + 4u * S::SXZ, &maps[TM_SXZ_XZ], bar, cx - HX, cz - H, cy, AF_SXZ);
                if (hints) {
                    tmaLoad4DHint(st + 4u * S::OWNV, &maps[TM_V_P], bar, cx, cz, cy, AF_VX, polOnce);
                    tmaLoad4DHint(st + 4u * S::OWNR, &maps[TM_R_P], bar, cx, cz, cy, AM_RIX, polOnce);
                } else {
                    tmaLoad4D(st + 4u * S::OWNV, &maps[TM_V_P], bar, cx, cz, cy, AF_VX);
                    tmaLoad4D(st + 4u * S::OWNR, &maps[TM_R_P], bar, cx, cz, cy, AM_RIX);
                }
            }
            if (++stage == NSTV) {
                stage = 0;
                parity ^= 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stress half-step (ForwardSolver3Delastic.cpp:288-404) incl. free-surface correction.
// Role r = velocity component: role 0 differentiates vx (Dxb vx = vxx, Dyf vx, Dzf vx), role 1 vy (Dxf vy, Dyb vy = vyy,
// Dzf vy), role 2 vz (Dxf vz, Dyf vz, Dzb vz = vzz) — one x, one y and one z derivative each, the shape of the velocity
// roles, so the three roles again share ONE instruction stream (three role-specific streams do not fit the instruction
// cache: measured) and carry equal work.  Every stress needs derivatives of two or three components, so the roles swap
// them through shared memory with split-phase mbarriers (publish, fetch the own-point operands, then wait):
//   every role publishes its normal strain rate and one shear term and updates its normal stress and one shear stress
//   role 0: sxx, sxy += muxy (Dyf vx + Dxf vy[role 1])    role 1: syy, syz += muyz (Dzf vy + Dyf vz[role 2])
//   role 2: szz, sxz += muxz (Dxf vz + Dzf vx[role 0])
// All y operators of this half-step are the plain ones (ForwardSolver3Delastic.cpp:289), so only the y-CPML layers and
// the free-surface plane need the generic step.
// ---------------------------------------------------------------------------------------------------------------------
struct SRole {
    int oZ9, opX, opY;  // order-reducing edges: z stencil from row z - H, x / y operator of this role
    bool xEdge, yFwd;
    const float *wz;
    int r;
    int oP;             // own point inside a plain tile
    int oX, oZ, oF;     // stage offsets: x stencil row, z stencil start, feed tile
    int oSN, oSS, oMS;  // own normal stress, own shear stress, modulus of the shear stress
    int oED, oES, oIS;  // exchange buffer: published normal strain rate / shear term, fetched shear term
    float cx[WS_MAXQ + 1];
    float *outN, *outS; // output fields
    float *psx, *psy, *psz;
    bool halfY;
};

template <int Q, int R, bool GENERIC, bool XZ, bool POL1>
__device__ __forceinline__ void strPlane(const WsParams &P, const Thr &t, const CpT &cpt, const SRole &ro, F4 (&q)[Cfg<Q>::QL], const float *st, float *xch, int n,
                                         long long o, float *psx, float *psz, const YDyn<Q> &yd, float *psy, int gy)
{
    using S = StageS<Q>;
    using C = Cfg<Q>;
    constexpr int NP = C::N_P;
    const bool ycp = GENERIC && yd.ky >= 0 && t.active;
    F4 pz, py;
    if (XZ && cpt.kz >= 0) // staged by the producer together with the operands of this plane
        pz = ld4(st + t.oPZ);
    if (ycp)
        py = ld4(psy);
    F4 a = (POL1 && XZ && ro.xEdge) ? dX9t<Q>(st + ro.oX, P.tab, ro.opX, t.x0, P.nx) : dX9<Q>(st + ro.oX, ro.cx);
    if (XZ && cpt.kxv >= 0)
        cpApplyXv(cpt, st + t.oPX, psx, t.cxTab, P.psiPitchX, a);
    F4 b = dY<Q, R>(q, (GENERIC && POL1) ? yd.w : P.cw);
    if (ycp)
        cpApply4(psy, py, yd.ya, yd.yb, b);
    F4 c = (POL1 && XZ && ro.wz) ? dZ9t<Q, C::TXH>(st + ro.oZ9, ro.wz) : dZ<Q, false, C::TXH>(st + ro.oZ, P.cw);
    if (XZ && cpt.kz >= 0)
        cpApply4(psz, pz, cpt.za, cpt.zb, c);
    // publish the normal strain rate and the shear term another role needs; keep the shear term of the own shear stress
    F4 dg, ex, kp;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        dg.v[p] = ro.r == 0 ? a.v[p] : (ro.r == 1 ? b.v[p] : c.v[p]);
        ex.v[p] = ro.r == 0 ? c.v[p] : (ro.r == 1 ? a.v[p] : b.v[p]);
        kp.v[p] = ro.r == 0 ? b.v[p] : (ro.r == 1 ? c.v[p] : a.v[p]);
    }
    if (n > 0)
        mbarWait(t.barXFree, (uint32_t)(n - 1) & 1u); // every warp has fetched the terms of the previous plane
    st4(xch + ro.oED, dg);
    st4(xch + ro.oES, ex);
    __syncwarp();
    if (t.lane == 0)
        mbarArrive(t.barXFull);
    F4 sN = ld4(st + ro.oSN), sS = ld4(st + ro.oSS);
    const F4 pi = ld4(st + S::OWNM + ro.oP), mu = ld4(st + S::OWNM + NP + ro.oP), ms = ld4(st + ro.oMS);
    mbarWait(t.barXFull, (uint32_t)n & 1u);
    const F4 d0 = ld4(xch + ro.oP), d1 = ld4(xch + NP + ro.oP), d2 = ld4(xch + 2 * NP + ro.oP), im = ld4(xch + ro.oIS);
    __syncwarp();
    if (t.lane == 0)
        mbarArrive(t.barXFree);
#pragma unroll
    for (int p = 0; p < 4; p++) {
        // ForwardSolver3Delastic.cpp:297-314: S_ii += pi (vxx+vyy+vzz); S_ii -= 2 mu (sum of the other two)
        float u = A::add(d0.v[p], d1.v[p]);
        u = A::add(u, d2.v[p]);
        u = A::mul(u, pi.v[p]);
        sN.v[p] = A::add(sN.v[p], u);
        const float lo = ro.r == 0 ? d1.v[p] : d0.v[p], hi = ro.r == 2 ? d1.v[p] : d2.v[p];
        u = A::mul(A::add(lo, hi), mu.v[p]);
        sN.v[p] = A::msub(2.0f, u, sN.v[p]);
        // :331-382: S_ij += mu_ij (D_j v_i + D_i v_j)
        const float tt = A::add(kp.v[p], im.v[p]);
        sS.v[p] = A::add(sS.v[p], A::mul(tt, ms.v[p]));
    }
    if (GENERIC && P.free_surface == 1 && gy == 0 && t.active) {
        // FreeSurface3Delastic.cpp:15-47, FreeSurface.cpp:13-20
        if (ro.r == 1) {
#pragma unroll
            for (int p = 0; p < 4; p++)
                sN.v[p] = A::mul(sN.v[p], 0.0f);
        } else {
            const F4 sH = ldg4(P.sH + (long long)t.z * P.nx + t.x0), sV = ldg4(P.sV + (long long)t.z * P.nx + t.x0);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float hor = A::add(d0.v[p], d2.v[p]);
                float tt = A::mul(sH.v[p], hor);
                sN.v[p] = A::add(sN.v[p], tt);
                tt = A::mul(sV.v[p], d1.v[p]);
                sN.v[p] = A::sub(sN.v[p], tt);
            }
        }
    }
    if (t.active && P.fastDebug != 2) {
        if (P.fastFlags & 4) {
            st4(ro.outN + o, sN);
            st4(ro.outS + o, sS);
        } else if (P.fastFlags & 8) {
            __stcg(reinterpret_cast<float4 *>(ro.outN + o), make_float4(sN.v[0], sN.v[1], sN.v[2], sN.v[3]));
            __stcg(reinterpret_cast<float4 *>(ro.outS + o), make_float4(sS.v[0], sS.v[1], sS.v[2], sS.v[3]));
        } else {
            st4cs(ro.outN + o, sN);
            st4cs(ro.outS + o, sS);
        }
    }
}

template <int Q, bool XZ, bool POL1, int R> struct StrUnroll {
    static __device__ __forceinline__ void run(const WsParams &P, const Thr &t, const CpT &cpt, const SRole &ro, F4 (&q)[Cfg<Q>::QL], const float *sm, float *xch, int n,
                                               int stage0, uint32_t parity, long long o, float *psx, float *psz, long long sxStride, long long szStride)
    {
        const int stage = stage0 + R;
        mbarWait(t.barFull + 8u * stage, parity);
        const float *st = sm + stage * t.stride;
        q[Q - 1 + R] = ld4(st + ro.oF);
        YDyn<Q> yd;
        yd.ky = -1;
        strPlane<Q, R, false, XZ, POL1>(P, t, cpt, ro, q, st, xch, n, o, psx, psz, yd, nullptr, 1);
        consumerRelease(t, stage);
        StrUnroll<Q, XZ, POL1, R + 1>::run(P, t, cpt, ro, q, sm, xch, n + 1, stage0, parity, o + P.plane, XZ ? psx + sxStride : psx, XZ ? psz + szStride : psz, sxStride,
                                     szStride);
    }
};
template <int Q, bool XZ, bool POL1> struct StrUnroll<Q, XZ, POL1, UNRS> {
    static __device__ __forceinline__ void run(const WsParams &, const Thr &, const CpT &, const SRole &, F4 (&)[Cfg<Q>::QL], const float *, float *, int, int, uint32_t,
                                               long long, float *, float *, long long, long long)
    {
    }
};

template <int Q, bool CPML, bool EDGE, bool POL1>
__device__ __forceinline__ void strConsumer(const WsParams &P, const float *sm, float *xch, Thr t, int G, int yc0, int yc1)
{
    using S = StageS<Q>;
    using C = Cfg<Q>;
    constexpr int H = C::H, HX = C::HX, TXH = C::TXH, NP = C::N_P;
    constexpr bool XZC = CPML && EDGE;
    SRole ro;
    ro.r = G;
    const int oP = t.lz * TX + 4 * t.lx;
    ro.oP = oP;
    const int tv = S::TV + G * C::N_XZ; // own velocity component with x and z halo
    ro.oX = tv + (t.lz + H) * TXH + 4 * t.lx;                    // x stencil: own row, column of x0 - HX
    ro.oZ = tv + (t.lz + (G == 2 ? 0 : 1)) * TXH + 4 * t.lx + HX; // z stencil: row z - H (backward, role 2) or z - H + 1 (forward)
    ro.oF = S::FEED + G * NP + oP;
    // order-reducing edges: operators Dxb | Dxf | Dxf, Dyf | Dyb | Dyf and Dzf | Dzf | Dzb of the three roles
    ro.oZ9 = G == 2 ? ro.oZ : ro.oZ - TXH;
    ro.opX = G == 0 ? OP_XB : OP_XF;
    ro.yFwd = G != 1;
    ro.opY = ro.yFwd ? OP_YF : OP_YB;
    ro.xEdge = POL1 && XZC && t.active && (t.x0 < H || t.x0 + 3 >= P.nx - H);
    ro.wz = (POL1 && XZC && t.active && (t.z < H || t.z >= P.nz - H))
                ? P.tab + ((size_t)(G == 2 ? OP_ZB : OP_ZF) * (2 * H + 1) + wsRowClass(t.z, P.nz, H)) * (Q + 1)
                : nullptr;
#pragma unroll
    for (int j = 0; j <= Q; j++) // offsets -H..H: backward (role 0) -H..H-1, forward -H+1..H
        ro.cx[j] = G == 0 ? (j < Q ? P.cw[j] : 0.0f) : (j >= 1 ? P.cw[j - 1] : 0.0f);
    // own-point tiles: stresses in arena order sxx sxy syy syz szz sxz, moduli in arena order pi mu muxy muxz muyz
    ro.oSN = S::OWNS + (2 * G) * NP + oP;
    ro.oSS = S::OWNS + (2 * G + 1) * NP + oP;
    ro.oMS = S::OWNM + (G == 0 ? 2 : (G == 1 ? 4 : 3)) * NP + oP;
    ro.oED = G * NP + oP;
    ro.oES = (3 + G) * NP + oP;
    ro.oIS = (3 + (G + 1) % 3) * NP + oP;
    ro.outN = P.fld[G == 0 ? F_SXX : (G == 1 ? F_SYY : F_SZZ)];
    ro.outS = P.fld[G == 0 ? F_SXY : (G == 1 ? F_SYZ : F_SXZ)];
    ro.halfY = G != 1;
    // memory-variable slots and profiles (CPML3D.cpp:31-153): normal strain rates full-grid, shear terms half-grid profiles
    const int sx = G == 0 ? PSI_VXX : (G == 1 ? PSI_VYX : PSI_VZX);
    const int sy = G == 0 ? PSI_VXY : (G == 1 ? PSI_VYY : PSI_VZY);
    const int sz = G == 0 ? PSI_VXZ : (G == 1 ? PSI_VYZ : PSI_VZZ);
    CpT cpt;
    cpSetup<XZC>(P, cpt, t.active, t.x0, t.z, /*halfX*/ G != 0, /*halfZ*/ G != 2);
    const bool warpXZ = XZC && P.fastDebug != 4 && __any_sync(0xffffffffu, cpt.kxv >= 0 || cpt.kz >= 0);
    const int W2 = 2 * P.W, PX = P.psiPitchX;
    const long long sxStride = (long long)P.nz * PX, szStride = (long long)W2 * P.nx;
    const long long rowOff = P.base + t.x0 + (long long)t.z * P.pitch;
    ro.psx = XZC ? P.psi[sx] + (long long)t.z * PX : nullptr;
    ro.psz = XZC ? P.psi[sz] + (long long)cpt.kz * P.nx + t.x0 : nullptr;
    ro.psy = CPML ? P.psi[sy] + (long long)t.z * P.nx + t.x0 : nullptr;
    const int xs = (P.psiBoxX == PX || t.x0 - 4 * t.lx < P.W) ? 0 : PX - P.psiBoxX;
    t.oPX = S::SIZE + (G * TZ + t.lz) * P.psiBoxX - xs;
    t.oPZ = S::SIZE + psxFloats(P.psiBoxX) + G * NP + oP;

    F4 q[C::QL];
#pragma unroll
    for (int k = 0; k < C::QL; k++)
        q[k] = zero4();

    const int nIter = (Q - 1) + (yc1 - yc0);
    int n = 0; // planes computed so far (phase of the exchange barriers)
    auto generic = [&](int it) {
        const int ly = yc0 - (Q - 1) + it, gy = P.gy0 + ly;
        const bool comp = it >= Q - 1;
        const int stage = it % NSTS;
        YDyn<Q> yd;
        yd.ky = -1;
        if (POL1) {
            // y weights of this plane: the plain operators (ForwardSolver3Delastic.cpp:289); with order-reducing edges the rows next
            // to the y faces have their own
            const int row = wsRowClass(min(max(gy, 0), P.gny - 1), P.gny, H);
            const float *tw = P.tab + ((size_t)ro.opY * (2 * H + 1) + row) * (Q + 1) + (ro.yFwd ? 1 : 0);
#pragma unroll
            for (int j = 0; j < Q; j++)
                yd.w[j] = __ldg(tw + j);
        }
        if (comp && CPML) {
            yd.ky = yCpmlIndex(P, gy);
            if (yd.ky >= 0) {
                yd.ya = __ldg((ro.halfY ? P.cayh : P.cay) + yd.ky);
                yd.yb = __ldg((ro.halfY ? P.cbyh : P.cby) + yd.ky);
            }
        }
        mbarWait(t.barFull + 8u * stage, (it / NSTS) & 1);
        const float *st = sm + stage * t.stride;
        q[Q - 1] = ld4(st + ro.oF);
        if (comp) {
            strPlane<Q, 0, true, XZC, POL1>(P, t, cpt, ro, q, st, xch, n, rowOff + (long long)ly * P.plane, XZC ? ro.psx + (long long)ly * sxStride : nullptr,
                                      XZC ? ro.psz + (long long)ly * szStride : nullptr, yd, CPML ? ro.psy + (long long)yd.ky * P.nz * P.nx : nullptr, gy);
            n++;
        }
        consumerRelease(t, stage);
#pragma unroll
        for (int k = 0; k < Q - 1; k++)
            q[k] = q[k + 1];
    };

    int it = 0;
    while (it < nIter) {
        const int ly0 = yc0 - (Q - 1) + it, gy0 = P.gy0 + ly0;
        // the first Q iterations (queue prologue + first plane, which may be the free-surface plane) are generic
        if (it >= Q && it % UNRS == 0 && it + UNRS <= nIter && !needsGeneric<CPML>(P, gy0, gy0 + UNRS - 1, H, false)) {
            const long long o = rowOff + (long long)ly0 * P.plane;
            const int stage0 = it % NSTS;
            const uint32_t parity = (it / NSTS) & 1;
            if (warpXZ)
                StrUnroll<Q, XZC, POL1, 0>::run(P, t, cpt, ro, q, sm, xch, n, stage0, parity, o, ro.psx + (long long)ly0 * sxStride, ro.psz + (long long)ly0 * szStride, sxStride,
                                          szStride);
            else
                StrUnroll<Q, false, POL1, 0>::run(P, t, cpt, ro, q, sm, xch, n, stage0, parity, o, nullptr, nullptr, 0, 0);
            n += UNRS;
#pragma unroll
            for (int k = 0; k < Q - 1; k++)
                q[k] = q[k + UNRS];
            it += UNRS;
        } else {
            generic(it);
            it++;
        }
    }
}

template <int Q, bool CPML, bool EDGE, bool POL1 = false> __global__ void __launch_bounds__(NGROUPS *Cfg<Q>::NTG + WS_FAST_PRODUCER_THREADS, 1) kFastStress(const __grid_constant__ WsParams P)
{
    using C = Cfg<Q>;
    using S = StageS<Q>;
    constexpr int H = C::H, HX = C::HX;
    constexpr bool XZC = CPML && EDGE;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ __align__(8) Bars bars;
    __shared__ __align__(16) float cxTab[XZC ? 4 * PXMAX : 4];
    float *sm = reinterpret_cast<float *>(smraw);
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(P.fastMaps);
    if (XZC)
        for (int k = threadIdx.x; k < 4 * P.psiPitchX; k += blockDim.x)
            cxTab[k] = __ldg(P.cxTab + k);

    const int tid = threadIdx.x;
    const int tile = P.fastTiles[P.fastTileBase + blockIdx.x]; // (z tile << 16) | x tile, see wsFastPrepare
    const int tx0 = (tile & 0xffff) * TX, tz0 = (tile >> 16) * TZ;
    const int yc0 = P.ylo + blockIdx.z * P.fastChunk;
    const int yc1 = min(P.yhi, yc0 + P.fastChunk);
    if (yc0 >= yc1)
        return;
    traceStart(P, 1);
    const uint32_t barFull = smemU32(&bars.full[0]), barEmpty = smemU32(&bars.empty[0]);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTS; s++) {
            mbarInit(barFull + 8u * s, 1);
            mbarInit(barEmpty + 8u * s, NGROUPS * C::WPG);
        }
        mbarInit(smemU32(&bars.xfull), NGROUPS * C::WPG);
        mbarInit(smemU32(&bars.xfree), NGROUPS * C::WPG);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int stride = S::SIZE + (XZC ? psiStageFloats(P.psiBoxX) : 0);
    const int grp = tid / C::NTG;
    if (grp < NGROUPS) {
        wsConsumerRegs();
        Thr t;
        const int tg = tid - grp * C::NTG;
        t.lx = tg % C::LXN;
        t.lz = tg / C::LXN;
        t.x0 = tx0 + 4 * t.lx;
        t.z = tz0 + t.lz;
        t.lane = tid & 31;
        t.active = (t.x0 < P.nx) && (t.z < P.nz);
        t.barFull = barFull;
        t.barEmpty = barEmpty;
        t.barXFull = smemU32(&bars.xfull);
        t.barXFree = smemU32(&bars.xfree);
        t.stride = stride;
        t.cxTab = cxTab;
        t.oPX = t.oPZ = t.oPZ2 = 0;
        strConsumer<Q, CPML, EDGE, POL1>(P, sm, sm + NSTS * stride, t, grp, yc0, yc1);
        traceEnd(P, 1, t.lane);
    } else if (wsProducerRegs(), tid == NGROUPS * C::NTG) {
        const int HZP = (P.nzp > 1) ? WS_HALO : 0;
        const int cx = WS_PADX + tx0, cz = HZP + tz0;
        const int nIter = (Q - 1) + (yc1 - yc0);
        const uint32_t smBase = smemU32(sm);
        int stage = 0;
        uint32_t parity = 1;
        const bool hints = (P.fastFlags & 1) != 0;
        const uint64_t polOnce = policyEvictFirst(), polKeep = policyEvictLast();
        const bool tileX = XZC && P.fastDebug != 4 && (tx0 < P.W || tx0 + TX > P.nx - P.W), tileZ = XZC && P.fastDebug != 4 && (tz0 < P.W || tz0 + TZ > P.nz - P.W);
        const int kz0 = tz0 < P.W ? tz0 : tz0 - (P.nz - 2 * P.W);
        const uint32_t bytesPX = 3u * TZ * (uint32_t)P.psiBoxX * 4u, bytesPZ = 3u * C::N_P * 4u;
        const int xs = (P.psiBoxX == P.psiPitchX || tx0 < P.W) ? 0 : P.psiPitchX - P.psiBoxX;
        const uint32_t bytesFull = S::BYTES_FULL + (tileX ? bytesPX : 0u) + (tileZ ? bytesPZ : 0u);
        for (int it = 0; it < nIter; it++) {
            const int cy = WS_HALO + yc0 - (Q - 1) + it;
            const bool comp = it >= Q - 1;
            if (it >= NSTS)
                mbarWait(barEmpty + 8u * stage, parity);
            const uint32_t st = smBase + (uint32_t)(stage * stride) * 4u;
            const uint32_t bar = barFull + 8u * stage;
            mbarExpectTx(bar, comp ? bytesFull : S::BYTES_FEED);
            if (comp && tileX)
                tmaLoad4D(st + 4u * S::SIZE, &maps[TM_PX], bar, xs, tz0, cy - WS_HALO, 3);
            if (comp && tileZ)
                tmaLoad4D(st + 4u * (S::SIZE + psxFloats(P.psiBoxX)), &maps[TM_PZ], bar, tx0, kz0, cy - WS_HALO, 3);
            if (hints) { // the velocities come back H (H-1) planes later as stencil tiles
                tmaLoad4DHint(st + 4u * S::FEED, &maps[TM_F1_P], bar, cx, cz, cy + H, AF_VX, polKeep);
                tmaLoad4DHint(st + 4u * (S::FEED + C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_VY, polKeep);
                tmaLoad4DHint(st + 4u * (S::FEED + 2 * C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H, AF_VZ, polKeep);
            } else {
                tmaLoad4D(st + 4u * S::FEED, &maps[TM_F1_P], bar, cx, cz, cy + H, AF_VX);
                tmaLoad4D(st + 4u * (S::FEED + C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H - 1, AF_VY);
                tmaLoad4D(st + 4u * (S::FEED + 2 * C::N_P), &maps[TM_F1_P], bar, cx, cz, cy + H, AF_VZ);
            }
            if (comp) {
                tmaLoad4D(st + 4u * S::TV, &maps[TM_V_XZ], bar, cx - HX, cz - H, cy, AF_VX);
                if (hints) {
                    tmaLoad4DHint(st + 4u * S::OWNS, &maps[TM_S_P], bar, cx, cz, cy, AF_SXX, polOnce);
                    tmaLoad4DHint(st + 4u * S::OWNM, &maps[TM_M_P], bar, cx, cz, cy, AM_PW, polOnce);
                } else {
                    tmaLoad4D(st + 4u * S::OWNS, &maps[TM_S_P], bar, cx, cz, cy, AF_SXX);
                    tmaLoad4D(st + 4u * S::OWNM, &maps[TM_M_P], bar, cx, cz, cy, AM_PW);
                }
            }
            if (++stage == NSTS) {
                stage = 0;
                parity ^= 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// cpmlEdge: the launch covers tiles in x / z CPML layers (staged memory variables); the stress half-step appends the
// derivative exchange buffer (6 plain tiles)
template <int Q> size_t smemBytes(int pass, bool cpmlEdge, int PX)
{
    const size_t stage = (pass == 0 ? StageV<Q>::SIZE : StageS<Q>::SIZE) + (cpmlEdge ? psiStageFloats(PX) : 0);
    return ((size_t)(pass == 0 ? NSTV : NSTS) * stage + (pass == 1 ? 6 * Cfg<Q>::N_P : 0)) * 4;
}

template <int Q> bool setAttrs()
{
    const void *ks[10] = {reinterpret_cast<const void *>(kFastVel<Q, true, true>),           reinterpret_cast<const void *>(kFastVel<Q, true, false>),
                          reinterpret_cast<const void *>(kFastVel<Q, false, false>),         reinterpret_cast<const void *>(kFastStress<Q, true, true>),
                          reinterpret_cast<const void *>(kFastStress<Q, true, false>),       reinterpret_cast<const void *>(kFastStress<Q, false, false>),
                          reinterpret_cast<const void *>(kFastVel<Q, true, true, true>),     reinterpret_cast<const void *>(kFastVel<Q, true, false, true>),
                          reinterpret_cast<const void *>(kFastStress<Q, true, true, true>), reinterpret_cast<const void *>(kFastStress<Q, true, false, true>)};
    for (const void *k : ks)
        if (wsOptInSmem(k, kMaxDynSmem) != cudaSuccess)
            return false;
    return true;
}

template <int Q> int launchQ(const WsParams &P, int pass, cudaStream_t st)
{
    if (!setAttrs<Q>())
        return 0;
    int launched = 0;
    const int ny = P.yhi - P.ylo;
    const bool cpml = P.damping == 2;
    const int nt = NGROUPS * Cfg<Q>::NTG + WS_FAST_PRODUCER_THREADS;
    // Tiles that touch an x / z CPML layer move up to 1.5x the bytes of an interior tile.  Launched together, the slow
    // tiles desynchronise the marches of neighbouring thread blocks, and halo rows that neighbours would share in L2
    // are fetched from HBM twice (measured: +20 % DRAM reads).  So the layer tiles (longest first) and the interior
    // tiles run as two launches: each is homogeneous enough for its waves of thread blocks to stay in step.
    for (int part = 0; part < 2; part++) {
        WsParams Q2 = P;
        Q2.fastTileBase = part == 0 ? 0 : P.fastNEdge;
        const int n = part == 0 ? P.fastNEdge : P.fastNTiles - P.fastNEdge;
        if (n <= 0)
            continue;
        launched++;
        const int chunk = part == 0 && P.fastNEdge > 0 ? P.fastChunkEdge : P.fastChunk;
        dim3 grid(n, 1, (ny + chunk - 1) / chunk);
        // the second part holds the interior tiles when the layer tiles have a launch of their own (fastNEdge > 0)
        const bool edge = cpml && (part == 0 || P.fastNEdge == 0);
        const size_t sm = smemBytes<Q>(pass, edge, P.psiBoxX);
        Q2.fastChunk = chunk;
        if (P.edge_policy == 1) { // order-reducing edges: CPML variants only (wsFastSupported)
            if (pass == 0) {
                if (edge)
                    kFastVel<Q, true, true, true><<<grid, nt, sm, st>>>(Q2);
                else
                    kFastVel<Q, true, false, true><<<grid, nt, sm, st>>>(Q2);
            } else {
                if (edge)
                    kFastStress<Q, true, true, true><<<grid, nt, sm, st>>>(Q2);
                else
                    kFastStress<Q, true, false, true><<<grid, nt, sm, st>>>(Q2);
            }
        } else if (pass == 0) {
            if (edge)
                kFastVel<Q, true, true><<<grid, nt, sm, st>>>(Q2);
            else if (cpml)
                kFastVel<Q, true, false><<<grid, nt, sm, st>>>(Q2);
            else
                kFastVel<Q, false, false><<<grid, nt, sm, st>>>(Q2);
        } else {
            if (edge)
                kFastStress<Q, true, true><<<grid, nt, sm, st>>>(Q2);
            else if (cpml)
                kFastStress<Q, true, false><<<grid, nt, sm, st>>>(Q2);
            else
                kFastStress<Q, false, false><<<grid, nt, sm, st>>>(Q2);
        }
    }
    return launched;
}

} // namespace

static unsigned long long *g_trace = nullptr;

// pass: 0 = velocity half-step only (the same statements in the elastic and the viscoelastic solver:
// ForwardSolver3Delastic.cpp:181-277 = ForwardSolver3Dviscoelastic.cpp:188-262), -1 = both half-steps (elastic)
bool wsFastSupported(const WsParams &P, bool exact, int pass)
{
    if (exact || P.dim != 3 || !(P.eq == WS_EQ_ELASTIC || (P.eq == WS_EQ_VISCOELASTIC && pass == 0)))
        return false;
    if (P.damping == 1)
        return false;
    // order-reducing edges: the edge columns / rows must lie in the layer tiles (x / z CPML layers at least q/2 wide)
    if (P.edge_policy != 0 && !(P.damping == 2 && P.W >= P.h))
        return false;
    if (P.q != 8 && P.q != 4)
        return false;
    if (P.nx % 4 != 0)
        return false;
    if (!P.fldArena || !P.matArena)
        return false;
    if (P.damping == 2) {
        // the x / z memory variables are staged by TMA: 16-byte slab rows (even W), a tile touches one side of an axis only,
        // and the staged rows must fit next to the operands
        if (!P.psiXArena || !P.psiZArena || !P.cxTab || P.psiPitchX > PXMAX || P.nx - P.W - (P.W + 3) / 4 * 4 < 0)
            return false;
        for (int tz0 = 0; tz0 < P.nz; tz0 += TZ)
            if (tz0 < P.W && tz0 + TZ > P.nz - P.W)
                return false;
        const int PX = P.psiPitchX;
        const size_t need = P.q == 8 ? std::max(smemBytes<8>(0, true, PX), smemBytes<8>(1, true, PX)) : std::max(smemBytes<4>(0, true, PX), smemBytes<4>(1, true, PX));
        if (need > (size_t)kMaxDynSmem)
            return false;
    }
    return true;
}

void *wsFastPrepare(WsParams &P, int nyp)
{
    const int H = P.h, HX = H <= 4 ? 4 : 8;
    const int TXH = TX + 2 * HX, TZH = TZ + 2 * H;
    std::vector<CUtensorMap> maps(TM_COUNT);
    auto mkF = [&](int slot, int bx, int bz, int bf) { maps[slot] = makeMap(P.fldArena, P.pitch, P.nzp, nyp, P.arenaStride, AF_COUNT, bx, bz, bf); };
    auto mkM = [&](int slot, int bx, int bz, int bf) { maps[slot] = makeMap(P.matArena, P.pitch, P.nzp, nyp, P.arenaStride, AM_COUNT, bx, bz, bf); };
    mkF(TM_V_P, TX, TZ, 3);
    mkM(TM_R_P, TX, TZ, 3);
    mkF(TM_SX_X, TXH, TZ, 2);
    mkF(TM_SXZ_XZ, TXH, TZH, 1);
    mkF(TM_SZ_XZ, TXH, TZH, 2);
    mkF(TM_F1_P, TX, TZ, 1);
    mkF(TM_V_XZ, TXH, TZH, 3);
    mkF(TM_S_P, TX, TZ, 6);
    mkM(TM_M_P, TX, TZ, 5);
    if (P.damping == 2) {
        const int W2 = 2 * P.W;
        // a tile stages its own side of the x rows only (entries [0, W4) or [PX - box, PX)) unless one tile spans both layers
        const int W4 = (P.W + 3) / 4 * 4;
        const bool oneSided = TX <= P.nx - P.W; // the first tile ends before the high layer starts, so no tile touches both
        P.psiBoxX = oneSided ? std::max(W4, P.psiPitchX - W4) : P.psiPitchX;
        maps[TM_PX] = makeMapG(P.psiXArena, P.psiPitchX, P.nz, P.nyl, APS_COUNT, P.psiBoxX, TZ, 3);
        maps[TM_PZ] = makeMapG(P.psiZArena, P.nx, W2, P.nyl, APS_COUNT, TX, TZ, 3);
    }
    // bit 0 = L2 eviction-priority hints on the TMA loads, bit 1 = layer tiles in a launch of their own.  With 64-plane chunks
    // the hints no longer pay (36.7 Gpt/s with, 37.1 without at 1024^3).  The separate launch of the layer tiles paid as long as the
    // consumers sat at 128 registers and the layer variant spilled (37.1 against 32.0 Gpt/s); with the producer warpgroup and the
    // register trade (ws_fast_common.cuh) ONE launch per half-step is faster again: 40.0 against 38.8 Gpt/s
    // (profiles/r02_northstar_notes.txt)
    P.fastFlags = getenv("WS_FAST_FLAGS") ? atoi(getenv("WS_FAST_FLAGS")) : 0;
    // tile list: layer tiles first (z layers and corners, then x layers: longest first; neighbours adjacent), then interior
    const int ntx = (P.nx + TX - 1) / TX, ntz = (P.nz + TZ - 1) / TZ;
    std::vector<int> tilesCorner, tilesZ, tilesX, tilesIn;
    const bool split = P.damping == 2 && (P.fastFlags & 2);
    for (int bz = 0; bz < ntz; bz++)
        for (int bx = 0; bx < ntx; bx++) {
            const bool ex = split && (bx * TX < P.W || bx * TX + TX > P.nx - P.W), ez = split && (bz * TZ < P.W || bz * TZ + TZ > P.nz - P.W);
            (ex && ez ? tilesCorner : (ez ? tilesZ : (ex ? tilesX : tilesIn))).push_back((bz << 16) | bx);
        }
    std::stable_sort(tilesX.begin(), tilesX.end(), [](int a, int b) { return (a & 0xffff) < (b & 0xffff); }); // column by column
    std::vector<int> tiles(tilesCorner);
    tiles.insert(tiles.end(), tilesZ.begin(), tilesZ.end());
    tiles.insert(tiles.end(), tilesX.begin(), tilesX.end());
    P.fastNEdge = (int)tiles.size();
    tiles.insert(tiles.end(), tilesIn.begin(), tilesIn.end());
    P.fastNTiles = (int)tiles.size();
    P.fastTileBase = 0;
    void *dev = nullptr;
    if (cudaMalloc(&dev, sizeof(CUtensorMap) * TM_COUNT + sizeof(int) * tiles.size()) != cudaSuccess)
        throw std::runtime_error("cudaMalloc for tensor maps failed");
    cudaMemcpy(dev, maps.data(), sizeof(CUtensorMap) * TM_COUNT, cudaMemcpyHostToDevice);
    cudaMemcpy(static_cast<char *>(dev) + sizeof(CUtensorMap) * TM_COUNT, tiles.data(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice);
    P.fastMaps = dev;
    P.fastTiles = reinterpret_cast<const int *>(static_cast<char *>(dev) + sizeof(CUtensorMap) * TM_COUNT);
    P.fastDebug = getenv("WS_FAST_DEBUG") ? atoi(getenv("WS_FAST_DEBUG")) : 0;
    P.fastTrace = nullptr;
    if (getenv("WS_FAST_TRACE")) {
        cudaMalloc(&g_trace, sizeof(unsigned long long) * 3 * 2 * WS_TRACE_MAX);
        cudaMemset(g_trace, 0, sizeof(unsigned long long) * 3 * 2 * WS_TRACE_MAX);
        P.fastTrace = g_trace;
    }
    P.fastChunk = wsPickChunk(P.fastNTiles - P.fastNEdge, P.nyl, P.q);
    P.fastChunkEdge = P.fastNEdge > 0 ? wsPickChunk(P.fastNEdge, P.nyl, P.q) : P.fastChunk;
    if (getenv("WS_FAST_CHUNK"))
        P.fastChunk = P.fastChunkEdge = atoi(getenv("WS_FAST_CHUNK"));
    return dev;
}

void wsFastRelease(void *maps)
{
    if (g_trace && getenv("WS_FAST_TRACE")) {
        std::vector<unsigned long long> h((size_t)3 * 2 * WS_TRACE_MAX);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), g_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(getenv("WS_FAST_TRACE"), "wb")) {
            fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
            fclose(f);
        }
        cudaFree(g_trace);
        g_trace = nullptr;
    }
    if (maps)
        cudaFree(maps);
}

int wsLaunchFast(const WsParams &P, int pass, cudaStream_t st)
{
    if (!P.fastMaps || P.yhi <= P.ylo)
        return 0;
    if (P.q == 8)
        return launchQ<8>(P, pass, st);
    if (P.q == 4)
        return launchQ<4>(P, pass, st);
    return 0;
}

// ws_kernels_march.cuh — marching kernels: every equation type with a compile-time FD order, for the configurations
// the warp-specialised TMA kernels (ws_kernels_fast.cu, 3-D elastic only) do not cover.
//
// 2.5-D blocking.  A thread block owns a TX x TZ tile of the x-z plane (a TX-wide strip of the row in 2-D) and marches
// along y, the slowest axis, in chunks of <= 64 planes (thread blocks of one y range march in step and share their halo
// rows in L2: ws_kernels_march.cu); a thread owns FOUR consecutive x points of one row (128-bit shared-memory loads,
// 128-bit global stores, index arithmetic shared by the four points) — ONE point in the 3-D elastic / viscoelastic
// half-steps, whose operands are so many that shared memory would otherwise leave 4-8 warps per SM.
//   * every operand of a plane is staged in shared memory by 16-byte cp.async copies into a ring of NST stages (2: the
//     next plane is in flight while one is computed; deeper rings cost resident thread blocks and lose, measured): the
//     halo tiles of the fields differentiated along x or z, the plane that enters each y window, and the own-point
//     operands (updated fields, model parameters, memory variables);
//   * y derivatives: every field differentiated along y lives in a REGISTER QUEUE of q+1 planes per thread, fed from the
//     staged plane y + q/2: each value is read from memory once per thread block instead of q times;
//   * the statement sequence is the one of the per-point kernels: the point type below only supplies the operands of
//     wsgen::passA / passB (derivatives D<F, OP>(), own-point values, CPML / ABS / free-surface terms) as NL-lane
//     values, every lane sees the scalar operation order, and the weights are applied in the same ascending-column
//     order, so the results are bit-identical to the per-point kernels in FMA mode (checked by the tests);
//   * tiles and planes that lie inside the grid on every axis (no edge rows, no CPML / ABS layer) run an instantiation
//     without any of the boundary code (a thread-block-uniform choice).
// Which fields are staged / queued per (equation, dimension, half-step) is the table `spec` below; it restates which
// operator the reference applies to which field (ForwardSolver/ForwardSolver{2D,3D}*.cpp, ForwardSolverEM/*.cpp run()).
#pragma once
#include "ws_kernels_general.cuh"

namespace wsmarch {

struct Lists {
    int nt; int t[5];   // fields differentiated along x or z: staged with their halo
    int nq; int qf[3];  // fields differentiated along y: register queues, fed from staged plain tiles of plane y + q/2
    int nf; int f[6];   // own-point wavefields the half-step reads (and updates)
    int nm; int m[10];  // own-point model parameters
    int nr; int r[6];   // memory-variable components RC_* the half-step updates, per relaxation mechanism
    int nc; int c[3];   // EM coefficient Cd axes, per relaxation mechanism
};

__host__ __device__ constexpr Lists spec(int EQ, int DIM, int PASS)
{
    const bool d3 = DIM == 3;
    switch (EQ) {
    case WS_EQ_ACOUSTIC: // ForwardSolver3Dacoustic.cpp:131-225, ForwardSolver2Dacoustic.cpp:121-190
        if (PASS == 0)
            return Lists{1, {F_P}, 1, {F_P}, d3 ? 3 : 2, {F_VX, F_VY, F_VZ}, d3 ? 3 : 2, {M_RIX, M_RIY, M_RIZ}, 0, {}, 0, {}};
        return d3 ? Lists{2, {F_VX, F_VZ}, 1, {F_VY}, 1, {F_P}, 1, {M_PW}, 0, {}, 0, {}} : Lists{1, {F_VX}, 1, {F_VY}, 1, {F_P}, 1, {M_PW}, 0, {}, 0, {}};
    case WS_EQ_ELASTIC:
    case WS_EQ_VISCOELASTIC: { // ForwardSolver3Delastic.cpp:181-404, ForwardSolver2Delastic.cpp:163-290, ForwardSolver3Dviscoelastic.cpp:188-416
        const bool v = EQ == WS_EQ_VISCOELASTIC;
        if (PASS == 0)
            return d3 ? Lists{5, {F_SXX, F_SXY, F_SXZ, F_SYZ, F_SZZ}, 3, {F_SXY, F_SYY, F_SYZ}, 3, {F_VX, F_VY, F_VZ}, 3, {M_RIX, M_RIY, M_RIZ}, 0, {}, 0, {}}
                      : Lists{2, {F_SXX, F_SXY}, 2, {F_SXY, F_SYY}, 2, {F_VX, F_VY}, 2, {M_RIX, M_RIY}, 0, {}, 0, {}};
        if (d3)
            // (own-point stresses in the order of the 3-D elastic arena, ws_api.cu: one TMA box fetches them together)
            return Lists{3, {F_VX, F_VY, F_VZ}, 3, {F_VX, F_VY, F_VZ}, 6, {F_SXX, F_SXY, F_SYY, F_SYZ, F_SZZ, F_SXZ},
                         v ? 10 : 5, {M_PW, M_MU, M_MUXY, M_MUXZ, M_MUYZ, M_TAUP, M_TAUS, M_TSXY, M_TSXZ, M_TSYZ},
                         v ? 6 : 0, {RC_XX, RC_YY, RC_ZZ, RC_XY, RC_XZ, RC_YZ}, 0, {}};
        return Lists{2, {F_VX, F_VY}, 2, {F_VX, F_VY}, 3, {F_SXX, F_SYY, F_SXY}, v ? 6 : 3, {M_PW, M_MU, M_MUXY, M_TAUP, M_TAUS, M_TSXY},
                     v ? 3 : 0, {RC_XX, RC_YY, RC_XY}, 0, {}};
    }
    case WS_EQ_SH:
    case WS_EQ_VISCOSH: { // ForwardSolver2Dsh.cpp:140-192, ForwardSolver2Dviscosh.cpp:190-236
        const bool v = EQ == WS_EQ_VISCOSH;
        if (PASS == 0)
            return Lists{1, {F_SXZ}, 1, {F_SYZ}, 1, {F_VZ}, 1, {M_INVRHO}, 0, {}, 0, {}};
        return Lists{1, {F_VZ}, 1, {F_VZ}, 2, {F_SXZ, F_SYZ}, v ? 5 : 2, {M_MUXZ, M_MUYZ, M_TAUS, M_TSXZ, M_TSYZ}, v ? 2 : 0, {RC_XZ, RC_YZ}, 0, {}};
    }
    case WS_EQ_TMEM:
    case WS_EQ_VISCOTMEM: // ForwardSolver2Dtmem.cpp:131-163, ForwardSolver2Dviscotmem.cpp:176-197
        if (PASS == 0)
            return Lists{1, {F_EZ}, 1, {F_EZ}, 2, {F_HX, F_HY}, 2, {M_MIYZ, M_MIXZ}, 0, {}, 0, {}};
        return Lists{1, {F_HY}, 1, {F_HX}, 1, {F_EZ}, 2, {M_CAZ, M_CBZ}, 1, {RC_Z}, 1, {RC_Z}};
    default: // EMEM / VISCOEMEM: ForwardSolver2Demem.cpp:136-169, ForwardSolver3Demem.cpp:154-232, ForwardSolver3Dviscoemem.cpp:161-312
        if (d3) {
            if (PASS == 0)
                return Lists{3, {F_EZ, F_EY, F_EX}, 2, {F_EZ, F_EX}, 3, {F_HX, F_HY, F_HZ}, 3, {M_MIYZ, M_MIXZ, M_MIXY}, 0, {}, 0, {}};
            return Lists{3, {F_HZ, F_HY, F_HX}, 2, {F_HZ, F_HX}, 3, {F_EX, F_EY, F_EZ}, 6, {M_CAX, M_CAY, M_CAZ, M_CBX, M_CBY, M_CBZ}, 3, {RC_X, RC_Y, RC_Z},
                         3, {RC_X, RC_Y, RC_Z}};
        }
        if (PASS == 0)
            return Lists{1, {F_EY}, 1, {F_EX}, 1, {F_HZ}, 1, {M_MIXY}, 0, {}, 0, {}};
        return Lists{1, {F_HZ}, 1, {F_HZ}, 2, {F_EX, F_EY}, 4, {M_CAX, M_CAY, M_CBX, M_CBY}, 2, {RC_X, RC_Y}, 2, {RC_X, RC_Y}};
    }
}
__host__ __device__ constexpr int findIn(const int *a, int n, int v)
{
    for (int k = 0; k < n; k++)
        if (a[k] == v)
            return k;
    return -1;
}
// every field a half-step differentiates (along x / z: `t`, along y: `qf`), each once: the halo tiles of the kernels that
// take the y stencils from shared memory too (2-D tile kernels, ws_kernels_tile2d.cuh)
struct HaloSet {
    int n; int f[8];
};
__host__ __device__ constexpr HaloSet haloSet(const Lists &S)
{
    HaloSet U{0, {}};
    for (int k = 0; k < S.nt; k++)
        U.f[U.n++] = S.t[k];
    for (int k = 0; k < S.nq; k++)
        if (findIn(S.t, S.nt, S.qf[k]) < 0)
            U.f[U.n++] = S.qf[k];
    return U;
}

// NL = consecutive x points per thread: 4 (128-bit accesses, index arithmetic shared by the points) or 1 (more resident
// threads per staged byte: the half-steps whose operands fill the shared memory, e.g. 3-D viscoelastic)
template <int DIM, int Q, int NL> struct Geo {
    static constexpr int H = Q / 2;
    static constexpr int HX = H <= 4 ? 4 : 8; // x halo rounded to whole 16-byte copies
#ifndef WS_MARCH_TX2D
#define WS_MARCH_TX2D 256 /* developer switch: width of the 2-D strips */
#endif
    static constexpr int TX = DIM == 3 ? (NL == 4 ? 64 : 32) : (NL == 4 ? WS_MARCH_TX2D : 128), TZ = DIM == 3 ? 8 : 1;
    static constexpr int HZ = DIM == 3 ? H : 0;
    static constexpr int LDX = TX + 2 * HX, NROW = TZ + 2 * HZ, TILE = LDX * NROW; // staged tile with halo
    static constexpr int TS = TILE;                                               // floats between two halo tiles of a stage
    static constexpr int NP = TX * TZ;                                            // plain tile (own points)
    static constexpr int LXN = TX / NL;                                           // threads per tile row
    static constexpr int NTHR = LXN * TZ;
};
constexpr int MAXPLAIN = 3 + 6 + 10 + 4 * (6 + 3); // plain tiles per stage: feeds, own fields, model parameters, L x (R + Cd)

// floats per stage
template <int EQ, int DIM, int Q, int PASS, int NL> __host__ __device__ constexpr int stageFloats(int L)
{
    constexpr Lists S = spec(EQ, DIM, PASS);
    using G = Geo<DIM, Q, NL>;
    return S.nt * G::TILE + (S.nq + S.nf + S.nm + L * (S.nr + S.nc)) * G::NP;
}

#ifdef WS_EMULATE
inline void cpAsync16(float *dst, const float *src) { std::memcpy(dst, src, 16); }
inline void cpCommit() {}
template <int N> inline void cpWait() {}
#else
__device__ __forceinline__ void cpAsync16(float *dst, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif
template <int N> __device__ __forceinline__ FV<N> ldv(const float *p) { return *reinterpret_cast<const FV<N> *>(p); }
template <int N> __device__ __forceinline__ void stv(float *p, const FV<N> &a) { *reinterpret_cast<FV<N> *>(p) = a; }

// NL consecutive x points (x0 .. x0+NL-1) of row z on the plane being computed.
// INTR = the points are known to be interior points of every axis (no edge rows, no CPML / ABS layer, not on the free
// surface): a thread-block-uniform property of (tile, plane), so the boundary code disappears from that instantiation.
// G = tile geometry; PB = planes per stage (the TMA kernels of ws_kernels_tma.cuh stage groups of PB planes in 2-D: every
// entry of a stage then holds PB tiles, one per plane, and `st` points at the tile of the plane being computed).
// YSM = the y stencils are read from the halo tiles as well (2-D tile kernels: a halo tile holds the rows y - H .. y + H of
// every differentiated field, LDX floats apart; no register queues, no feed tiles).
template <int EQ, int DIM, int Q, int PASS, int NL, bool INTR, class G_ = Geo<DIM, Q, NL>, int PB = 1, bool YSM = false>
struct MPt {
    using A = Ar<false>;
    using V = FV<NL>;
    using G = G_;
    static constexpr int H = Q / 2;
    static constexpr int NQ = spec(EQ, DIM, PASS).nq;
    static constexpr int TS = G::TS, PS = PB * G::NP; // floats between two halo-tile entries / two plain-tile entries
    static constexpr int O_Q = 0, O_F = O_Q + (YSM ? 0 : NQ) * PS, O_M = O_F + spec(EQ, DIM, PASS).nf * PS, O_R = O_M + spec(EQ, DIM, PASS).nm * PS;
    const WsParams &P;
    int x0, z, nAct; // nAct = lanes inside the grid (NL except in a ragged last column or an inactive row)
    int ly, gy;
    long long i;     // padded linear index of lane 0
    int ry, rz;      // derivative row classes of the plane / the row
    bool xEdge;      // some lane is an edge row of the x operators
    int ky, kz;      // CPML slab indices of the plane / the row (-1 outside)
    bool xLayer;     // some lane lies in an x CPML layer
    float *st;       // current stage, halo tiles [nt][PB][TILE] (at the tile of the current plane)
    float *sp;       // current stage, plain tiles [feeds | fields | model | R[l][nr] | Cd[l][nc]][PB][NP] (same)
    int so, op;      // lane 0 inside a halo tile / a plain tile
    int oC;          // offset of the Cd tiles (after the L * nr memory-variable tiles)
    bool stR;        // the memory variables and Cd coefficients are staged (else they are read from global memory)
    V (&q)[NQ][Q + 1];

    __device__ __forceinline__ MPt(const WsParams &P_, int x0_, int z_, int nAct_, int so_, int op_, V (&q_)[NQ][Q + 1])
        : P(P_), x0(x0_), z(z_), nAct(nAct_), st(nullptr), sp(nullptr), so(so_), op(op_), oC(O_R + P_.L * spec(EQ, DIM, PASS).nr * PS), stR(P_.marchStageR != 0), q(q_)
    {
        ly = gy = 0;
        i = 0;
        ry = rz = H;
        ky = kz = -1;
        xEdge = xLayer = false;
        if (!INTR) {
            rz = wsRowClass(z, P.nz, H);
            xEdge = x0 < H || x0 + NL - 1 >= P.nx - H;
            if (P.damping == 2) {
                xLayer = x0 < P.W || x0 + NL - 1 >= P.nx - P.W;
                if (DIM == 3)
                    kz = wsCpmlIndex(z, P.nz, P.W);
            }
        }
    }
    // plane ly, whose lane-0 point has the padded linear index i_
    __device__ __forceinline__ void setPlane(int ly_, long long i_)
    {
        ly = ly_;
        gy = P.gy0 + ly_;
        i = i_;
        if (!INTR) {
            ry = wsRowClass(gy, P.gny, H);
            ky = -1;
            if (P.damping == 2) {
                ky = wsCpmlIndex(gy, P.gny, P.W);
                if (P.free_surface != 0 && gy < P.W)
                    ky = -1; // no CPML in the top layer with a free surface (CPML3D.cpp:320-328)
            }
        }
    }

    template <int F, int OP> __device__ __forceinline__ V D() const
    {
        constexpr int axis = OP < 6 ? (OP >> 1) : 1;
        constexpr int fw = (OP & 1) == 0 ? 1 : 0; // forward operators: taps -H+1..H, backward: -H..H-1
        constexpr Lists S = spec(EQ, DIM, PASS);
        V acc(0.0f);
        if constexpr (axis == 1 && YSM) {
            constexpr HaloSet U = haloSet(S);
            constexpr int ti = findIn(U.f, U.n, F);
            if constexpr (ti >= 0) {
                const float *p = st + ti * TS + so - H * G::LDX;
                if (INTR || ry == H) {
#pragma unroll
                    for (int j = 0; j < Q; j++)
                        acc = A::madd(OP >= 6 ? P.cwy[j] : P.cw[j], ldv<NL>(p + (j + fw) * G::LDX), acc);
                } else {
                    const float *__restrict__ w = P.tab + ((size_t)OP * (2 * H + 1) + ry) * (Q + 1);
#pragma unroll
                    for (int j = 0; j <= Q; j++)
                        acc = A::madd(__ldg(w + j), ldv<NL>(p + j * G::LDX), acc);
                }
            }
        } else if constexpr (axis == 1) {
            constexpr int qi = findIn(S.qf, S.nq, F);
            if constexpr (qi >= 0) {
                if (INTR || ry == H) { // interior row: weights from the constant bank, the zero tap of the table row skipped
#pragma unroll
                    for (int j = 0; j < Q; j++)
                        acc = A::madd(OP >= 6 ? P.cwy[j] : P.cw[j], q[qi][j + fw], acc);
                } else {
                    const float *__restrict__ w = P.tab + ((size_t)OP * (2 * H + 1) + ry) * (Q + 1);
#pragma unroll
                    for (int j = 0; j <= Q; j++)
                        acc = A::madd(__ldg(w + j), q[qi][j], acc);
                }
            }
        } else if constexpr (axis == 2) {
            constexpr int ti = YSM ? findIn(haloSet(S).f, haloSet(S).n, F) : findIn(S.t, S.nt, F);
            if constexpr (ti >= 0) {
                const float *p = st + ti * TS + so - H * G::LDX;
                if (INTR || rz == H) {
#pragma unroll
                    for (int j = 0; j < Q; j++)
                        acc = A::madd(P.cw[j], ldv<NL>(p + (j + fw) * G::LDX), acc);
                } else {
                    const float *__restrict__ w = P.tab + ((size_t)OP * (2 * H + 1) + rz) * (Q + 1);
#pragma unroll
                    for (int j = 0; j <= Q; j++)
                        acc = A::madd(__ldg(w + j), ldv<NL>(p + j * G::LDX), acc);
                }
            }
        } else {
            constexpr int ti = YSM ? findIn(haloSet(S).f, haloSet(S).n, F) : findIn(S.t, S.nt, F);
            if constexpr (ti >= 0) {
                // the row around the points: offsets -H .. H+NL-1 (NL = 4: aligned 16-byte loads from -HX on)
                constexpr int HX = G::HX;
                constexpr int W0 = NL == 4 ? HX : H; // w[W0 + o] = value at offset o from lane 0
                float w[NL == 4 ? 2 * HX + 4 : 2 * H + 1];
                if constexpr (NL == 4) {
                    const float *p = st + ti * TS + so - HX;
#pragma unroll
                    for (int k = 0; k < (2 * HX + 4) / 4; k++) {
                        if (4 * k + 3 >= HX - H && 4 * k <= HX + H + 3) { // vectors that hold a tap of some lane
                            const FV<4> t = ldv<4>(p + 4 * k);
                            w[4 * k] = t.v[0]; w[4 * k + 1] = t.v[1]; w[4 * k + 2] = t.v[2]; w[4 * k + 3] = t.v[3];
                        } else
                            w[4 * k] = w[4 * k + 1] = w[4 * k + 2] = w[4 * k + 3] = 0.0f;
                    }
                } else {
                    const float *p = st + ti * TS + so - H;
#pragma unroll
                    for (int k = 0; k <= 2 * H + NL - 1; k++)
                        w[k] = p[k];
                }
                if (INTR || !xEdge) {
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        float a = 0.0f;
#pragma unroll
                        for (int j = 0; j < Q; j++)
                            a = A::madd(P.cw[j], w[W0 - H + j + fw + l], a);
                        acc.v[l] = a;
                    }
                } else {
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        const float *__restrict__ c = P.tab + ((size_t)OP * (2 * H + 1) + wsRowClass(x0 + l, P.nx, H)) * (Q + 1);
                        float a = 0.0f;
#pragma unroll
                        for (int j = 0; j <= Q; j++)
                            a = A::madd(__ldg(c + j), w[W0 - H + j + l], a);
                        acc.v[l] = a;
                    }
                }
            }
        }
        return acc;
    }

    // CPML.cpp:84-95 applyCPML on the lanes inside a layer: psi = b psi + a d ; d = d + psi (all memory variables are read
    // before the first one is written)
    __device__ __forceinline__ V cpx(V d, int slot, bool half) const
    {
        if (INTR || !xLayer)
            return d;
        const float *__restrict__ ca = half ? P.caxh : P.cax, *__restrict__ cb = half ? P.cbxh : P.cbx;
        float *ps = P.psi[slot] + ((long long)ly * P.nz + z) * P.psiPitchX;
        int k[NL], o[NL];
        float old[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) {
            k[l] = l < nAct ? wsCpmlIndex(x0 + l, P.nx, P.W) : -1;
            o[l] = wsPsiXIndex(x0 + l, P.W, P.psiDX);
            old[l] = k[l] >= 0 ? __ldg(ps + o[l]) : 0.0f;
        }
#pragma unroll
        for (int l = 0; l < NL; l++)
            if (k[l] >= 0) {
                float v = A::mul(old[l], __ldg(cb + k[l]));
                const float t = A::mul(__ldg(ca + k[l]), d.v[l]);
                v = A::add(v, t);
                ps[o[l]] = v;
                d.v[l] = A::add(d.v[l], v);
            }
        return d;
    }
    // a memory variable is read once (before it is written) and by this thread only, so the read may take the
    // read-only path: the loads of all terms of a half-step can then be issued ahead of the stores of the earlier terms
    __device__ __forceinline__ V cpRow(V d, float *ps, float a, float b) const
    {
        if (NL == 4 && nAct == 4 && (P.nx & 3) == 0) {
            const float4 o4 = __ldg(reinterpret_cast<const float4 *>(ps));
            const float old[4] = {o4.x, o4.y, o4.z, o4.w};
            V nw;
#pragma unroll
            for (int l = 0; l < NL; l++) {
                float v = A::mul(old[l], b);
                const float t = A::mul(a, d.v[l]);
                v = A::add(v, t);
                nw.v[l] = v;
                d.v[l] = A::add(d.v[l], v);
            }
            stv<NL>(ps, nw);
            return d;
        }
        float old[NL];
#pragma unroll
        for (int l = 0; l < NL; l++)
            old[l] = l < nAct ? __ldg(ps + l) : 0.0f;
#pragma unroll
        for (int l = 0; l < NL; l++)
            if (l < nAct) {
                float v = A::mul(old[l], b);
                const float t = A::mul(a, d.v[l]);
                v = A::add(v, t);
                ps[l] = v;
                d.v[l] = A::add(d.v[l], v);
            }
        return d;
    }
    __device__ __forceinline__ V cpy(V d, int slot, bool half) const
    {
        if (INTR || ky < 0)
            return d;
        return cpRow(d, P.psi[slot] + ((long long)ky * P.nz + z) * P.nx + x0, __ldg((half ? P.cayh : P.cay) + ky), __ldg((half ? P.cbyh : P.cby) + ky));
    }
    __device__ __forceinline__ V cpz(V d, int slot, bool half) const
    {
        if (INTR || kz < 0)
            return d;
        return cpRow(d, P.psi[slot] + ((long long)ly * (2 * P.W) + kz) * P.nx + x0, __ldg((half ? P.cazh : P.caz) + kz), __ldg((half ? P.cbzh : P.cbz) + kz));
    }
    __device__ __forceinline__ V absFactor() const
    {
        V r(1.0f);
        if (!INTR && P.damping == 1) {
#pragma unroll
            for (int l = 0; l < NL; l++)
                r.v[l] = wsgen::wsAbsFactor(P, x0 + l, gy, z);
        }
        return r;
    }
    // own-point operands: from the stage when the half-step's lists name them, else from global memory
    __device__ __forceinline__ V ldGlobal(const float *g) const
    {
        V r(0.0f);
#pragma unroll
        for (int l = 0; l < NL; l++)
            if (l < nAct)
                r.v[l] = g[l];
        return r;
    }
    __device__ __forceinline__ void stGlobal(float *g, const V &v) const
    {
        if (nAct == NL)
            stv<NL>(g, v);
        else {
#pragma unroll
            for (int l = 0; l < NL; l++)
                if (l < nAct)
                    g[l] = v.v[l];
        }
    }
    template <int F> __device__ __forceinline__ V fld() const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.f, S.nf, F);
        if constexpr (k >= 0)
            return ldv<NL>(sp + O_F + k * PS + op);
        else
            return ldGlobal(P.fld[F] + i);
    }
    template <int F> __device__ __forceinline__ void put(const V &v) const { stGlobal(P.fld[F] + i, v); }
    template <int M> __device__ __forceinline__ V mat() const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.m, S.nm, M);
        if constexpr (k >= 0)
            return ldv<NL>(sp + O_M + k * PS + op);
        else
            return ldGlobal(P.mat[M] + i);
    }
    template <int C> __device__ __forceinline__ V rget(int l) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.r, S.nr, C);
        if (k >= 0 && stR)
            return ldv<NL>(sp + O_R + (l * S.nr + k) * PS + op);
        return ldGlobal(P.fld[F_R0 + 6 * l + C] + i);
    }
    // write-through: later statements of the same half-step read the updated memory variable again
    template <int C> __device__ __forceinline__ void rput(int l, const V &v) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.r, S.nr, C);
        stGlobal(P.fld[F_R0 + 6 * l + C] + i, v);
        if (k >= 0 && stR)
            stv<NL>(sp + O_R + (l * S.nr + k) * PS + op, v);
    }
    template <int AXIS> __device__ __forceinline__ V cd(int l) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.c, S.nc, AXIS);
        if (k >= 0 && stR)
            return ldv<NL>(sp + oC + (l * S.nc + k) * PS + op);
        return ldGlobal(P.mat[M_CD0 + 3 * l + AXIS] + i);
    }
    // free-surface scalings of these columns (only read on the plane y = 0)
    __device__ __forceinline__ V sH() const { return ldGlobal(P.sH + (long long)z * P.nx + x0); }
    __device__ __forceinline__ V sV() const { return ldGlobal(P.sV + (long long)z * P.nx + x0); }
    __device__ __forceinline__ V sRH(int l) const { return ldGlobal(P.sRH[l] + (long long)z * P.nx + x0); }
    __device__ __forceinline__ V sRV(int l) const { return ldGlobal(P.sRV[l] + (long long)z * P.nx + x0); }
};

// NST = depth of the stage ring (planes y .. y+NST-2 are in flight while plane y is computed)
template <int EQ, int DIM, int Q, int PASS, int NST, int NL>
__global__ void __launch_bounds__(Geo<DIM, Q, NL>::NTHR) kMarch(const __grid_constant__ WsParams P)
{
    using G = Geo<DIM, Q, NL>;
    using MPG = MPt<EQ, DIM, Q, PASS, NL, false>;
    using MPI = MPt<EQ, DIM, Q, PASS, NL, true>;
    using V = FV<NL>;
    constexpr Lists S = spec(EQ, DIM, PASS);
    constexpr int H = G::H, NQ = S.nq, NT = S.nt, NP = G::NP;
#ifdef WS_EMULATE
    float *sm = reinterpret_cast<float *>(wsemu::g_smem);
    static const float *srcTab[MAXPLAIN]; // one thread block at a time in the emulation
#else
    extern __shared__ float4 wsMarchSmem[];
    float *sm = reinterpret_cast<float *>(wsMarchSmem);
    __shared__ const float *srcTab[MAXPLAIN];
#endif
    const int tid = threadIdx.x;
    const int lx = tid % G::LXN, lz = tid / G::LXN;
    const int tx0 = blockIdx.x * G::TX, tz0 = blockIdx.y * G::TZ;
    const int yc0 = P.ylo + blockIdx.z * P.marchChunk;
    const int yc1 = min(P.yhi, yc0 + P.marchChunk);
    if (yc0 >= yc1)
        return;
    const int x0 = tx0 + NL * lx, z = tz0 + lz;
    const int nAct = z < P.nz ? max(0, min(NL, P.nx - x0)) : 0;
    const bool active = nAct > 0;
    const int L = P.marchStageR ? P.L : 0; // relaxation mechanisms whose memory variables / Cd coefficients are staged
    const int nPlain = NQ + S.nf + S.nm + L * (S.nr + S.nc);
    const int stride = NT * G::TILE + nPlain * NP;
    // interior tile: every point is at least `lo` points away from the x and z faces (no edge rows, no CPML / ABS layer)
    const int lo = max(H, P.damping != 0 ? P.W : 0);
    const bool tileIn = tx0 >= lo && tx0 + G::TX <= P.nx - lo && (DIM == 2 || (tz0 >= lo && tz0 + G::TZ <= P.nz - lo));

    // plain tiles of a stage, in stage order: source = array origin of this tile (feeds: q/2 planes ahead)
    const long long ownOrg = P.base + tx0 + (long long)tz0 * P.pitch;
    if (tid < nPlain) {
        int a = tid;
        const float *p;
        long long yoff = 0;
        if (a < NQ) {
            p = P.fld[S.qf[a]];
            yoff = H;
        } else if ((a -= NQ) < S.nf) {
            p = P.fld[S.f[a]];
        } else if ((a -= S.nf) < S.nm) {
            p = P.mat[S.m[a]];
        } else if ((a -= S.nm) < L * S.nr) {
            const int nr = S.nr > 0 ? S.nr : 1;
            p = P.fld[F_R0 + 6 * (a / nr) + S.r[a % nr]];
        } else {
            a -= L * S.nr;
            const int nc = S.nc > 0 ? S.nc : 1;
            p = P.mat[M_CD0 + 3 * (a / nc) + S.c[a % nc]];
        }
        srcTab[tid] = p + ownOrg + yoff * P.plane;
    }
    __syncthreads();

    // stage the operands of a plane: halo tiles of the x / z differentiated fields, then the plain tiles (NL = 4: a thread
    // copies exactly the 16 bytes of each plain tile it reads itself).  Chunk offsets are the same for every plane.
    constexpr int TCH = G::TILE / 4, NCT = (TCH + G::NTHR - 1) / G::NTHR; // 16-byte chunks per halo tile / per thread
    int tSrc[NCT], tDst[NCT];
#pragma unroll
    for (int n = 0; n < NCT; n++) {
        const int c = tid + n * G::NTHR;
        const int row = c / (G::LDX / 4), col = c - row * (G::LDX / 4);
        tSrc[n] = c < TCH ? row * P.pitch + 4 * col : -1;
        tDst[n] = row * G::LDX + 4 * col;
    }
    const int op = lz * G::TX + NL * lx, so = (lz + G::HZ) * G::LDX + NL * lx + G::HX;
    constexpr int CH = NP / 4, APR = G::NTHR / CH; // 16-byte chunks per plain tile / plain tiles per round of the block
    const int pc = tid % CH, pa0 = tid / CH;
    const int pSrc = (pc / (G::TX / 4)) * P.pitch + 4 * (pc % (G::TX / 4));
    const long long tileOrg = P.base + (tx0 - G::HX) + (long long)(tz0 - G::HZ) * P.pitch;
    auto stage = [&](long long yo, int slot) {
        float *dst = sm + slot * stride;
#pragma unroll
        for (int k = 0; k < NT; k++) {
            const float *src = P.fld[S.t[k]] + tileOrg + yo;
#pragma unroll
            for (int n = 0; n < NCT; n++)
                if (tSrc[n] >= 0)
                    cpAsync16(dst + k * G::TILE + tDst[n], src + tSrc[n]);
        }
        float *dp = dst + NT * G::TILE + 4 * pc;
        for (int a = pa0; a < nPlain; a += APR)
            cpAsync16(dp + a * NP, srcTab[a] + yo + pSrc);
        cpCommit();
    };

    // register queues: q[f][j] = plane ly - H + j of queued field f while plane ly is computed
    V q[NQ][Q + 1];
    const long long own0 = P.base + x0 + (long long)z * P.pitch;
#pragma unroll
    for (int f = 0; f < NQ; f++) {
        const float *src = P.fld[S.qf[f]] + own0;
        q[f][0] = V(0.0f);
#pragma unroll
        for (int j = 1; j <= Q; j++)
            q[f][j] = active ? ldv<NL>(src + (long long)(yc0 - 1 - H + j) * P.plane) : V(0.0f);
    }

    MPG tg(P, x0, z, nAct, so, op, q);
    MPI ti(P, x0, z, nAct, so, op, q);
    long long yo = (long long)yc0 * P.plane; // plane offset of the next plane to stage
#pragma unroll
    for (int s = 0; s < NST - 1; s++) {
        if (yc0 + s < yc1)
            stage(yo, s);
        else
            cpCommit(); // empty group: the group count stays uniform
        yo += P.plane;
    }
    int slot = 0;
    long long iCur = own0 + (long long)yc0 * P.plane;
    for (int ly = yc0; ly < yc1; ly++) {
        cpWait<NST - 2>(); // this thread's copies of plane ly have landed
        __syncthreads();   // ... and everybody else's; everybody has finished plane ly - 1
        {
            const int refill = slot == 0 ? NST - 1 : slot - 1; // the stage of plane ly - 1
            if (ly + NST - 1 < yc1)
                stage(yo, refill);
            else
                cpCommit();
            yo += P.plane;
        }
        float *st = sm + slot * stride;
#pragma unroll
        for (int f = 0; f < NQ; f++) {
#pragma unroll
            for (int j = 0; j < Q; j++)
                q[f][j] = q[f][j + 1];
            q[f][Q] = ldv<NL>(st + NT * G::TILE + f * NP + op);
        }
        const int gy = P.gy0 + ly;
        if (P.marchDebug == 1) { // developer switch: staging only (memory-side ceiling of the skeleton)
        } else if (tileIn && gy >= lo && gy < P.gny - lo) { // uniform over the thread block
            ti.setPlane(ly, iCur);
            ti.st = st;
            ti.sp = st + NT * G::TILE;
            if (PASS == 0)
                wsgen::passA<EQ, DIM, false>(P, ti);
            else
                wsgen::passB<EQ, DIM, false>(P, ti);
        } else if (active) {
            tg.setPlane(ly, iCur);
            tg.st = st;
            tg.sp = st + NT * G::TILE;
            if (PASS == 0)
                wsgen::passA<EQ, DIM, false>(P, tg);
            else
                wsgen::passB<EQ, DIM, false>(P, tg);
        }
        iCur += P.plane;
        slot = slot + 1 == NST ? 0 : slot + 1;
    }
    cpWait<0>();
}

} // namespace wsmarch

// ws_kernels_march.cuh — marching kernels: every equation type with a compile-time FD order, for the configurations
// the warp-specialised TMA kernels (ws_kernels_fast.cu, 3-D elastic only) do not cover.
//
// 2.5-D blocking.  A thread block owns a TX x TZ tile of the x-z plane (a TX-wide strip of the row in 2-D) and marches
// along y, the slowest axis; a thread owns one grid column (x, z).
//   * y derivatives: every field that is differentiated along y lives in a REGISTER QUEUE of q+1 planes per thread;
//     one global load per field and plane feeds it (prefetched one plane ahead), so each value is read from memory once
//     per thread block instead of q times;
//   * x / z derivatives: the plane of every field differentiated along x or z is staged, with its halo, in shared memory
//     by 16-byte cp.async copies into a double buffer (plane y+1 is in flight while plane y is computed);
//   * everything else (own-point operands, CPML memory variables, ABS factors, free surface) and the statement sequence
//     itself are the ones of the per-point kernels: the point type below only replaces the derivative D<F, OP>() of
//     wsgen::passA / passB, and the weights are applied in the same ascending-column order, so the results are
//     bit-identical to the general kernels in FMA mode (checked by the tests).
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
            return Lists{3, {F_VX, F_VY, F_VZ}, 3, {F_VX, F_VY, F_VZ}, 6, {F_SXX, F_SYY, F_SZZ, F_SXY, F_SXZ, F_SYZ},
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

template <int DIM, int Q> struct Geo {
    static constexpr int H = Q / 2;
    static constexpr int HX = H <= 4 ? 4 : 8; // x halo rounded to whole 16-byte copies
    static constexpr int TX = DIM == 3 ? 32 : 128, TZ = DIM == 3 ? 8 : 1;
    static constexpr int HZ = DIM == 3 ? H : 0;
    static constexpr int LDX = TX + 2 * HX, NROW = TZ + 2 * HZ, TILE = LDX * NROW; // staged tile with halo
    static constexpr int NP = TX * TZ;                                            // plain tile (own points)
    static constexpr int NTHR = TX * TZ;
};
constexpr int MAXPLAIN = 3 + 6 + 10 + 4 * (6 + 3); // plain tiles per stage: feeds, own fields, model parameters, L x (R + Cd)

// floats per stage
template <int EQ, int DIM, int Q, int PASS> __host__ __device__ constexpr int stageFloats(int L)
{
    constexpr Lists S = spec(EQ, DIM, PASS);
    using G = Geo<DIM, Q>;
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

// point of the marching kernels: derivatives from the staged plane (x, z) and the register queues (y), own-point
// operands from the staged plain tiles
template <int EQ, int DIM, int Q, int PASS>
struct MPt : wsgen::PtBase<false> {
    using A = Ar<false>;
    using G = Geo<DIM, Q>;
    static constexpr int NQ = spec(EQ, DIM, PASS).nq;
    static constexpr int O_Q = spec(EQ, DIM, PASS).nt * G::TILE, O_F = O_Q + NQ * G::NP, O_M = O_F + spec(EQ, DIM, PASS).nf * G::NP,
                         O_R = O_M + spec(EQ, DIM, PASS).nm * G::NP;
    float *st;  // current stage: halo tiles [nt][TILE], then plain tiles [feeds | fields | model | R[l][nr] | Cd[l][nc]][NP]
    int so, op; // own point inside a halo tile / a plain tile
    int oC;     // offset of the Cd tiles (after the L * nr memory-variable tiles)
    float (&q)[NQ][Q + 1];

    __device__ __forceinline__ MPt(const WsParams &P_, int x_, int ly_, int z_, int so_, int op_, float (&q_)[NQ][Q + 1])
        : wsgen::PtBase<false>(P_, x_, ly_, z_), st(nullptr), so(so_), op(op_), oC(O_R + P_.L * spec(EQ, DIM, PASS).nr * G::NP), q(q_)
    {
    }

    template <int F, int OP> __device__ __forceinline__ float D() const
    {
        constexpr int axis = OP < 6 ? (OP >> 1) : 1;
        constexpr int fw = (OP & 1) == 0 ? 1 : 0; // forward operators: taps -H+1..H, backward: -H..H-1
        constexpr int H = Q / 2;
        constexpr Lists S = spec(EQ, DIM, PASS);
        float acc = 0.0f;
        if constexpr (axis == 1) {
            constexpr int qi = findIn(S.qf, S.nq, F);
            if constexpr (qi >= 0) {
                if (ry == H) { // interior row: weights from the constant bank, the zero tap of the table row skipped
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
        } else {
            constexpr int ti = findIn(S.t, S.nt, F);
            if constexpr (ti >= 0) {
                constexpr int sd = axis == 0 ? 1 : G::LDX;
                const float *p = st + ti * G::TILE + so - H * sd;
                const int r = axis == 0 ? rx : rz;
                if (r == H) {
#pragma unroll
                    for (int j = 0; j < Q; j++)
                        acc = A::madd(P.cw[j], p[(j + fw) * sd], acc);
                } else {
                    const float *__restrict__ w = P.tab + ((size_t)OP * (2 * H + 1) + r) * (Q + 1);
#pragma unroll
                    for (int j = 0; j <= Q; j++)
                        acc = A::madd(__ldg(w + j), p[j * sd], acc);
                }
            }
        }
        return acc;
    }
    // own-point operands: from the stage when the half-step's lists name them, else from global memory
    template <int F> __device__ __forceinline__ float fld() const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.f, S.nf, F);
        if constexpr (k >= 0)
            return st[O_F + k * G::NP + op];
        else
            return P.fld[F][i];
    }
    template <int F> __device__ __forceinline__ void put(float v) const { P.fld[F][i] = v; }
    template <int M> __device__ __forceinline__ float mat() const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.m, S.nm, M);
        if constexpr (k >= 0)
            return st[O_M + k * G::NP + op];
        else
            return P.mat[M][i];
    }
    template <int C> __device__ __forceinline__ float rget(int l) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.r, S.nr, C);
        if constexpr (k >= 0)
            return st[O_R + (l * S.nr + k) * G::NP + op];
        else
            return P.fld[F_R0 + 6 * l + C][i];
    }
    // write-through: later statements of the same half-step read the updated memory variable again
    template <int C> __device__ __forceinline__ void rput(int l, float v) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.r, S.nr, C);
        P.fld[F_R0 + 6 * l + C][i] = v;
        if constexpr (k >= 0)
            st[O_R + (l * S.nr + k) * G::NP + op] = v;
    }
    template <int AXIS> __device__ __forceinline__ float cd(int l) const
    {
        constexpr Lists S = spec(EQ, DIM, PASS);
        constexpr int k = findIn(S.c, S.nc, AXIS);
        if constexpr (k >= 0)
            return st[oC + (l * S.nc + k) * G::NP + op];
        else
            return P.mat[M_CD0 + 3 * l + AXIS][i];
    }
};

// NST = depth of the stage ring (planes y .. y+NST-2 are in flight while plane y is computed)
template <int EQ, int DIM, int Q, int PASS, int NST>
__global__ void __launch_bounds__(Geo<DIM, Q>::NTHR) kMarch(const __grid_constant__ WsParams P)
{
    using G = Geo<DIM, Q>;
    using MP = MPt<EQ, DIM, Q, PASS>;
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
    const int lx = tid % G::TX, lz = tid / G::TX;
    const int tx0 = blockIdx.x * G::TX, tz0 = blockIdx.y * G::TZ;
    const int yc0 = P.ylo + blockIdx.z * P.marchChunk;
    const int yc1 = min(P.yhi, yc0 + P.marchChunk);
    if (yc0 >= yc1)
        return;
    const int x = tx0 + lx, z = tz0 + lz;
    const bool active = x < P.nx && z < P.nz;
    const int L = P.L;
    const int nPlain = NQ + S.nf + S.nm + L * (S.nr + S.nc);
    const int stride = NT * G::TILE + nPlain * NP;

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

    // stage the operands of plane ly: halo tiles of the x / z differentiated fields, then the plain tiles
    const long long tileOrg = P.base + (tx0 - G::HX) + (long long)(tz0 - G::HZ) * P.pitch;
    auto stage = [&](int ly, int slot) {
        float *dst = sm + slot * stride;
        const long long yo = (long long)ly * P.plane;
#pragma unroll
        for (int k = 0; k < NT; k++) {
            const float *src = P.fld[S.t[k]] + tileOrg + yo;
            for (int c = tid; c < G::TILE / 4; c += G::NTHR) {
                const int row = c / (G::LDX / 4), col = c - row * (G::LDX / 4);
                cpAsync16(dst + k * G::TILE + row * G::LDX + 4 * col, src + (long long)row * P.pitch + 4 * col);
            }
        }
        constexpr int CH = NP / 4, CPR = G::TX / 4; // 16-byte chunks per plain tile / per tile row
        float *dp = dst + NT * G::TILE;
        for (int g = tid; g < nPlain * CH; g += G::NTHR) {
            const int a = g / CH, c = g - a * CH;
            const int row = c / CPR, col = c - row * CPR;
            cpAsync16(dp + a * NP + row * G::TX + 4 * col, srcTab[a] + yo + (long long)row * P.pitch + 4 * col);
        }
        cpCommit();
    };

    // register queues: q[f][j] = plane ly - H + j of queued field f while plane ly is computed
    float q[NQ][Q + 1];
    const long long own0 = P.base + x + (long long)z * P.pitch;
#pragma unroll
    for (int f = 0; f < NQ; f++) {
        const float *src = P.fld[S.qf[f]] + own0;
        q[f][0] = 0.0f;
#pragma unroll
        for (int j = 1; j <= Q; j++)
            q[f][j] = active ? __ldg(src + (long long)(yc0 - 1 - H + j) * P.plane) : 0.0f;
    }

    const int op = lz * G::TX + lx;
    MP t(P, x, yc0, z, (lz + G::HZ) * G::LDX + lx + G::HX, op, q);
#pragma unroll
    for (int s = 0; s < NST - 1; s++) {
        if (yc0 + s < yc1)
            stage(yc0 + s, s);
        else
            cpCommit(); // empty group: the group count stays uniform
    }
    int slot = 0;
    for (int ly = yc0; ly < yc1; ly++) {
        cpWait<NST - 2>(); // this thread's copies of plane ly have landed
        __syncthreads();   // ... and everybody else's; everybody has finished plane ly - 1
        {
            const int refill = slot == 0 ? NST - 1 : slot - 1; // the stage of plane ly - 1
            if (ly + NST - 1 < yc1)
                stage(ly + NST - 1, refill);
            else
                cpCommit();
        }
        float *st = sm + slot * stride;
#pragma unroll
        for (int f = 0; f < NQ; f++) {
#pragma unroll
            for (int j = 0; j < Q; j++)
                q[f][j] = q[f][j + 1];
            q[f][Q] = st[MP::O_Q + f * NP + op];
        }
        if (active) {
            t.setY(ly);
            t.st = st;
            if (PASS == 0)
                wsgen::passA<EQ, DIM, false>(P, t);
            else
                wsgen::passB<EQ, DIM, false>(P, t);
        }
        slot = slot + 1 == NST ? 0 : slot + 1;
    }
    cpWait<0>();
}

} // namespace wsmarch

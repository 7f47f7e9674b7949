// ws_kernels_sparse.cuh — operator-given mode: the grid is irregular (variable grid spacing, variable FD order: Coordinates.cpp
// :115-247, Derivatives.cpp:129-1656), so the derivative operators are not stencils but rows assembled point by point.  The
// host layer assembles them by the reference's rules and hands them over in ELL form (ws_set_operator); the kernels below run
// the reference's statement sequence with a gather per operator row, fused per half-step like the regular-grid kernels
// (acoustic solvers: ForwardSolver2Dacoustic.cpp:121-190, ForwardSolver3Dacoustic.cpp:131-229).  Row sums run in ascending
// column order with an accumulator that starts at 0, like the CSR SpMV they replace.
#pragma once
#include "ws_common.cuh"

namespace wssparse {

enum { SP_XF = 0, SP_XB, SP_YF, SP_YB, SP_ZF, SP_ZB, SP_NOPS };
enum { SPSI_P_X = 0, SPSI_P_Y, SPSI_P_Z, SPSI_VXX, SPSI_VYY, SPSI_VZZ, SPSI_COUNT };

struct Params {
    long long n;
    int dim;
    // ELL operators, column-major: entry k of row i at [k * n + i]; col < 0 = no entry
    const int *col[SP_NOPS];
    const float *val[SP_NOPS];
    int taps[SP_NOPS];
    // CPML: point -> entry of the axis profile (-1 outside the layer), coefficients per entry (full / half grid), memory variables
    const int *cpK[3];
    const float *ca[3], *cb[3], *cah[3], *cbh[3];
    float *psi[SPSI_COUNT];
    float *vx, *vy, *vz, *p;
    const float *rix, *riy, *riz, *pw;
};

template <bool EXACT> __device__ __forceinline__ float row(const Params &Q, int op, long long i, const float *__restrict__ x)
{
    using A = Ar<EXACT>;
    const int *__restrict__ c = Q.col[op] + i;
    const float *__restrict__ v = Q.val[op] + i;
    float acc = 0.0f;
    for (int k = 0; k < Q.taps[op]; k++) {
        const int j = __ldg(c + (long long)k * Q.n);
        if (j >= 0)
            acc = A::madd(__ldg(v + (long long)k * Q.n), x[j], acc);
    }
    return acc;
}

// CPML.cpp:84-95 applyCPML: temp = a; Psi *= b; temp *= Vec; Psi += temp; Vec += Psi
template <bool EXACT> __device__ __forceinline__ float cpml(const Params &Q, int axis, int slot, bool half, long long i, float u)
{
    using A = Ar<EXACT>;
    if (!Q.cpK[axis])
        return u;
    const int k = __ldg(Q.cpK[axis] + i);
    if (k < 0)
        return u;
    float ps = A::mul(Q.psi[slot][k], __ldg((half ? Q.cbh[axis] : Q.cb[axis]) + k));
    const float t = A::mul(__ldg((half ? Q.cah[axis] : Q.ca[axis]) + k), u);
    ps = A::add(ps, t);
    Q.psi[slot][k] = ps;
    return A::add(u, ps);
}

// particle velocities: v_a += rho_a^-1 (.) P_a(D_af p)
template <bool EXACT> __global__ void __launch_bounds__(256) kVelAcoustic(const Params Q)
{
    using A = Ar<EXACT>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.n)
        return;
    float u = row<EXACT>(Q, SP_XF, i, Q.p);
    u = cpml<EXACT>(Q, 0, SPSI_P_X, true, i, u);
    u = A::mul(u, Q.rix[i]);
    Q.vx[i] = A::add(Q.vx[i], u);
    u = row<EXACT>(Q, SP_YF, i, Q.p); // the image-method operator when FreeSurface = 1 (the caller passes it as "Dyf")
    u = cpml<EXACT>(Q, 1, SPSI_P_Y, true, i, u);
    u = A::mul(u, Q.riy[i]);
    Q.vy[i] = A::add(Q.vy[i], u);
    if (Q.dim == 3) {
        u = row<EXACT>(Q, SP_ZF, i, Q.p);
        u = cpml<EXACT>(Q, 2, SPSI_P_Z, true, i, u);
        u = A::mul(u, Q.riz[i]);
        Q.vz[i] = A::add(Q.vz[i], u);
    }
}

// pressure: p += M (.) (P_x(D_xb vx) + P_y(D_yb vy) + P_z(D_zb vz))
template <bool EXACT> __global__ void __launch_bounds__(256) kPressure(const Params Q)
{
    using A = Ar<EXACT>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q.n)
        return;
    float u = row<EXACT>(Q, SP_XB, i, Q.vx);
    u = cpml<EXACT>(Q, 0, SPSI_VXX, false, i, u);
    float w = row<EXACT>(Q, SP_YB, i, Q.vy);
    w = cpml<EXACT>(Q, 1, SPSI_VYY, false, i, w);
    u = A::add(u, w);
    if (Q.dim == 3) {
        w = row<EXACT>(Q, SP_ZB, i, Q.vz);
        w = cpml<EXACT>(Q, 2, SPSI_VZZ, false, i, w);
        u = A::add(u, w);
    }
    u = A::mul(u, Q.pw[i]);
    Q.p[i] = A::add(Q.p[i], u);
}

// interpolation on the interface planes (Derivatives.cpp:1252-1566): out[rows[r]] = sum_k vals[r][k] * in[cols[r][k]] for the
// listed rows (every other row of the matrix is the identity).  Two phases, because rows read points other rows overwrite.
template <bool EXACT> __global__ void __launch_bounds__(256) kInterpGather(long long nrows, int taps, const int *__restrict__ cols, const float *__restrict__ vals, const float *__restrict__ in,
                                                                            float *__restrict__ tmp)
{
    using A = Ar<EXACT>;
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows)
        return;
    float acc = 0.0f;
    for (int k = 0; k < taps; k++) {
        const int j = cols[(long long)k * nrows + r];
        if (j >= 0)
            acc = A::madd(vals[(long long)k * nrows + r], in[j], acc);
    }
    tmp[r] = acc;
}
__global__ void __launch_bounds__(256) kInterpScatter(long long nrows, const int *__restrict__ rows, const float *__restrict__ tmp, float *__restrict__ out)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows)
        out[rows[r]] = tmp[r];
}
// FreeSurface.cpp:13-20 setSurfaceZero
__global__ void __launch_bounds__(256) kSurfaceZero(long long n, const int *__restrict__ idx, float *__restrict__ p)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n)
        p[idx[r]] = __fmul_rn(p[idx[r]], 0.0f);
}

// ABS2D.cpp / ABS3D.cpp apply: p, vx, vy (, vz) *= damping on the points of the frame (every other entry of the sparse vector is 1)
template <bool EXACT> __global__ void __launch_bounds__(256) kAbsDamp(long long n, const int *__restrict__ idx, const float *__restrict__ damp, float *__restrict__ p, float *__restrict__ vx,
                                                                       float *__restrict__ vy, float *__restrict__ vz)
{
    using A = Ar<EXACT>;
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
        return;
    const int i = idx[r];
    const float d = damp[r];
    p[i] = A::mul(p[i], d);
    vx[i] = A::mul(vx[i], d);
    vy[i] = A::mul(vy[i], d);
    if (vz)
        vz[i] = A::mul(vz[i], d);
}

} // namespace wssparse

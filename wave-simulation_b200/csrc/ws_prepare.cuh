// ws_prepare.cuh — Modelparameter::prepareForModelling and ForwardSolver*::prepareForModelling as device kernels.
//
// The reference builds averaging matrices (Modelparameter.cpp:336-624) and applies them as SpMVs; here each derived
// vector is one matrix-free pass with the same summation order (ascending column index of the CSR row) and the same
// clamps (ModelparameterSeismic.cpp:421-432, Common.hpp:68-141).  No FMA contraction: these run once per model.
#pragma once
#include "ws_common.cuh"

namespace wsprep {

struct Geo {
    int nx, nyl, nz, gny, gy0, pitch, nzp;
    long long plane, base;
    int ylo, yhi; // local plane range to process (may extend into the halo: -WS_HALO .. nyl+WS_HALO)
    __device__ __forceinline__ long long idx(int x, int ly, int z) const { return base + x + (long long)z * pitch + (long long)ly * plane; }
};

#define WSPREP_POINT                                                                                                   \
    const int x = blockIdx.x * blockDim.x + threadIdx.x;                                                               \
    const int z = WS_POINT_Z(g.nz);                                                                                    \
    const int ly = g.ylo + WS_POINT_PLANE(g.nz);                                                                       \
    if (x >= g.nx || z >= g.nz || ly >= g.yhi)                                                                         \
        return;                                                                                                        \
    const int gy = g.gy0 + ly;                                                                                         \
    if (gy < 0 || gy >= g.gny)                                                                                         \
        return;                                                                                                        \
    const long long i = g.idx(x, ly, z);

// ModelparameterSeismic.cpp:131-136 (+ Viscoelastic.cpp:650-713 when scale != 0: modulus / (1 + sum * tau))
__global__ void kModulus(Geo g, const float *__restrict__ v, const float *__restrict__ rho, const float *__restrict__ tau, float sum, float *__restrict__ out)
{
    WSPREP_POINT
    float m = __fmul_rn(__fmul_rn(rho[i], v[i]), v[i]);
    if (tau) {
        const float t = __fadd_rn(1.0f, __fmul_rn(sum, tau[i]));
        m = __fdiv_rn(m, t);
    }
    out[i] = m;
}

// Common.hpp:68-120 searchAndReplace(vec, thr, val, 1 := "<")
__global__ void kClampLess(Geo g, float *__restrict__ a, float thr, float val)
{
    WSPREP_POINT
    if (a[i] < thr)
        a[i] = val;
}

__device__ __forceinline__ float wsInvalidToZero(float v) { return (isnan(v) || isinf(v)) ? 0.0f : v; }

// out = 1 / a  (ModelparameterSeismic.cpp:164-172 inverseDensity)
__global__ void kInverse(Geo g, const float *__restrict__ a, float *__restrict__ out)
{
    WSPREP_POINT
    out[i] = __fdiv_rn(1.0f, a[i]);
}

// 2-point average along `axis` (Modelparameter.cpp:336-447): 0.5 a[i] + 0.5 a[i+1], weight 1.0 at the far edge.
// mode 0: plain (calcAveragedParameter), 1: inverse + NaN/Inf -> 0 (calcInverseAveragedParameter, :633-639),
// mode 2: harmonic with clamps (calcAveragedSWaveModulus; input already clamped to >= 1)
__device__ __forceinline__ float wsFinish(float s, int mode)
{
    if (mode == 0)
        return s;
    float r = __fdiv_rn(1.0f, s);
    if (mode == 1)
        return wsInvalidToZero(r);
    return r < 4.0f ? 0.0f : r;
}
__device__ __forceinline__ float wsLoad(const float *__restrict__ a, long long i, int mode) { return mode == 2 ? __fdiv_rn(1.0f, a[i]) : a[i]; }

__global__ void kAvg2(Geo g, const float *__restrict__ a, float *__restrict__ out, int axis, int mode)
{
    WSPREP_POINT
    const int c = axis == 0 ? x : (axis == 1 ? gy : z);
    const int n = axis == 0 ? g.nx : (axis == 1 ? g.gny : g.nz);
    const long long st = axis == 0 ? 1 : (axis == 1 ? g.plane : (long long)g.pitch);
    float s = 0.0f;
    if (c + 1 < n) {
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i, mode)));
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i + st, mode)));
    } else
        s = __fadd_rn(s, __fmul_rn(1.0f, wsLoad(a, i, mode)));
    out[i] = wsFinish(s, mode);
}

// 4-point average in the plane of axes (A,B), A the faster axis (Modelparameter.cpp:449-624): full rows 1/4 each, faces
// 1/2 + 1/2, edges 1.0; terms accumulated in ascending column order.
__global__ void kAvg4(Geo g, const float *__restrict__ a, float *__restrict__ out, int axA, int axB, int mode)
{
    WSPREP_POINT
    const int cA = axA == 0 ? x : (axA == 1 ? gy : z), cB = axB == 0 ? x : (axB == 1 ? gy : z);
    const int nA = axA == 0 ? g.nx : (axA == 1 ? g.gny : g.nz), nB = axB == 0 ? g.nx : (axB == 1 ? g.gny : g.nz);
    const long long sA = axA == 0 ? 1 : (axA == 1 ? g.plane : (long long)g.pitch);
    const long long sB = axB == 0 ? 1 : (axB == 1 ? g.plane : (long long)g.pitch);
    const bool inA = cA + 1 < nA, inB = cB + 1 < nB;
    float s = 0.0f;
    if (inA && inB) {
        s = __fadd_rn(s, __fmul_rn(0.25f, wsLoad(a, i, mode)));
        s = __fadd_rn(s, __fmul_rn(0.25f, wsLoad(a, i + sA, mode)));
        s = __fadd_rn(s, __fmul_rn(0.25f, wsLoad(a, i + sB, mode)));
        s = __fadd_rn(s, __fmul_rn(0.25f, wsLoad(a, i + sA + sB, mode)));
    } else if (inA && !inB) {
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i, mode)));
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i + sA, mode)));
    } else if (!inA && inB) {
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i, mode)));
        s = __fadd_rn(s, __fmul_rn(0.5f, wsLoad(a, i + sB, mode)));
    } else
        s = __fadd_rn(s, __fmul_rn(1.0f, wsLoad(a, i, mode)));
    out[i] = wsFinish(s, mode);
}

// FreeSurfaceElastic.cpp:11-47 on the y = 0 plane (dense nz*nx output)
__global__ void kFreeSurfaceElastic(Geo g, const float *__restrict__ pi, const float *__restrict__ mu, float *__restrict__ sH, float *__restrict__ sV, int *__restrict__ bad)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nx || z >= g.nz)
        return;
    const long long i = g.idx(x, 0, z);
    if (!(mu[i] > 0.0f))
        atomicExch(bad, 1);
    float t = __fsub_rn(pi[i], __fmul_rn(2.0f, mu[i]));
    const int k = z * g.nx + x;
    sV[k] = __fmul_rn(1.0f, t);
    t = __fmul_rn(t, t);
    t = __fdiv_rn(t, pi[i]);
    t = __fmul_rn(t, -1.0f);
    sH[k] = __fmul_rn(1.0f, t);
}

// FreeSurfaceViscoelastic.cpp:12-96
struct ViscoFS {
    int L;
    float fL, relaxTime[4], viscoCoeff2[4];
    float *sRH[4], *sRV[4];
};
__global__ void kFreeSurfaceVisco(Geo g, const float *__restrict__ pi, const float *__restrict__ mu, const float *__restrict__ tauP, const float *__restrict__ tauS, float *__restrict__ sH, float *__restrict__ sV, ViscoFS v, int *__restrict__ bad)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.nx || z >= g.nz)
        return;
    const long long i = g.idx(x, 0, z);
    const int k = z * g.nx + x;
    if (!(mu[i] > 0.0f))
        atomicExch(bad, 1);
    const float optp = __fadd_rn(1.0f, __fmul_rn(v.fL, tauP[i]));
    const float opts = __fadd_rn(1.0f, __fmul_rn(v.fL, tauS[i]));
    float temp = __fmul_rn(__fmul_rn(-2.0f, mu[i]), opts);
    float temp2 = __fmul_rn(pi[i], optp);
    temp = __fadd_rn(temp, temp2);
    temp2 = __fdiv_rn(1.0f, temp2);
    sV[k] = __fmul_rn(1.0f, temp);
    float h = __fmul_rn(-1.0f, 1.0f);
    h = __fmul_rn(h, temp);
    h = __fmul_rn(h, temp);
    h = __fmul_rn(h, temp2);
    sH[k] = h;
    for (int l = 0; l < v.L; l++) {
        float t = __fmul_rn(2.0f, mu[i]);
        float t2 = pi[i];
        float t3 = __fmul_rn(t, tauS[i]);
        t3 = __fsub_rn(t3, __fmul_rn(t2, tauP[i]));
        float rv = __fmul_rn(1.0f, t3);
        rv = __fmul_rn(rv, v.viscoCoeff2[l]);
        rv = __fdiv_rn(rv, v.relaxTime[l]);
        v.sRV[l][k] = rv;
        t = __fmul_rn(t, opts);
        t2 = __fmul_rn(t2, optp);
        t = __fdiv_rn(t, t2);
        t = __fsub_rn(t, 1.0f);
        float rh = __fmul_rn(1.0f, t3);
        rh = __fmul_rn(rh, t);
        rh = __fmul_rn(rh, v.viscoCoeff2[l]);
        rh = __fdiv_rn(rh, v.relaxTime[l]);
        v.sRH[l][k] = rh;
    }
}

// ---- EM coefficient builders (ForwardSolverEM.cpp:14-154) --------------------------------------------------------------
__device__ __forceinline__ float wsCinv(float eps, float sig, float DT)
{
    float v = __fdiv_rn(0.5f, eps);
    v = __fmul_rn(v, sig);
    v = __fmul_rn(v, DT);
    v = __fadd_rn(v, 1.0f);
    return __fdiv_rn(1.0f, v);
}
struct EmCoef {
    int L;
    float DT, eps0;
    float sumInvRelax;     // mean of 1/relaxationTime
    float cdScalar[4];     // 1/(1+0.5 DT/tau_l) / (L tau_l^2)
    float *cd[4];
};
// eps, sig (and tauEps, tauSig when visco != 0) are the (averaged) static parameters of one E component
__global__ void kEmCoefficients(Geo g, const float *__restrict__ eps, const float *__restrict__ sig, const float *__restrict__ tauEps, const float *__restrict__ tauSig, int visco, EmCoef c, float *__restrict__ Ca, float *__restrict__ Cb)
{
    WSPREP_POINT
    float e = eps[i], s = sig[i];
    if (visco) {
        // getDielectricPermittivityEffectiveOptical :140-154, getElectricConductivityEffectiveOptical :122-135
        float eo = __fsub_rn(1.0f, tauEps[i]);
        eo = __fmul_rn(eo, eps[i]);
        eo = __fadd_rn(eo, __fmul_rn(sig[i], tauSig[i]));
        if (eo < c.eps0)
            eo = c.eps0;
        float so = __fmul_rn(tauEps[i], c.sumInvRelax);
        so = __fmul_rn(so, eps[i]);
        so = __fadd_rn(so, sig[i]);
        for (int l = 0; l < c.L; l++) { // getAveragedCd :95-117
            float v = __fmul_rn(tauEps[i], c.cdScalar[l]);
            v = __fmul_rn(v, eps[i]);
            c.cd[l][i] = __fmul_rn(v, -c.DT);
        }
        e = eo;
        s = so;
    }
    const float cinv = wsCinv(e, s, c.DT);
    float v = __fdiv_rn(0.5f, e);
    v = __fmul_rn(v, s);
    v = __fmul_rn(v, c.DT);
    v = __fsub_rn(1.0f, v);
    Ca[i] = __fmul_rn(v, cinv);
    Cb[i] = __fmul_rn(__fdiv_rn(1.0f, e), cinv);
}

// pack / unpack between the dense reference layout and the padded HBM layout
__global__ void kPack(Geo g, const float *__restrict__ dense, float *__restrict__ padded, int denseY0 /* local y of dense plane 0 */)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = WS_POINT_Z(g.nz);
    const int ly = g.ylo + WS_POINT_PLANE(g.nz);
    if (x >= g.nx || z >= g.nz || ly >= g.yhi)
        return;
    padded[g.idx(x, ly, z)] = dense[((long long)(ly - denseY0) * g.nz + z) * g.nx + x];
}
__global__ void kUnpack(Geo g, const float *__restrict__ padded, float *__restrict__ dense)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = WS_POINT_Z(g.nz);
    const int ly = g.ylo + WS_POINT_PLANE(g.nz);
    if (x >= g.nx || z >= g.nz || ly >= g.yhi)
        return;
    dense[((long long)(ly - g.ylo) * g.nz + z) * g.nx + x] = padded[g.idx(x, ly, z)];
}

// isFinite over the interior of one field (Wavefields::isFinite, Simulation.cpp:519)
__global__ void kIsFinite(Geo g, const float *__restrict__ a, int *__restrict__ bad)
{
    WSPREP_POINT
    const float v = a[i];
    if (isnan(v) || isinf(v))
        atomicExch(bad, 1);
}

} // namespace wsprep

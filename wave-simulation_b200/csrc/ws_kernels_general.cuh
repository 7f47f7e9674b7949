// ws_kernels_general.cuh — general matrix-free kernels: every equation type, every FD order (2..12), both edge
// policies, image-method free surface, CPML and ABS, one thread per grid point.  These kernels are the correctness
// workhorse (bit-exact against the CPU oracle in EXACT mode) and the fallback for configurations the tiled fast
// kernels (ws_kernels_fast.cu) do not cover.
//
// Statement order follows the reference run() functions line by line; citations are relative to src/.
#pragma once
#include "ws_common.cuh"

namespace wsgen {

template <int N> struct IC { // compile-time integer tag (array slots as arguments of generic lambdas)
    static constexpr int value = N;
};

// ABS3D.cpp:182-213 / ABS2D.cpp:141-172: damping factor of a point (1 outside the frame)
__device__ __forceinline__ float wsAbsFactor(const WsParams &P, int x, int gy, int z)
{
    if (P.damping != 1)
        return 1.0f;
    const int W = P.W;
    const int dx = min(x, P.nx - 1 - x), dy = min(gy, P.gny - 1 - gy);
    int m;
    if (P.dim == 3) {
        const int dz = min(z, P.nz - 1 - z);
        if (P.free_surface == 0) {
            m = min(min(dx, dy), dz);
        } else if (gy < W) {
            m = (dz < W || dx < W) ? min(dx, dz) : W;
        } else
            m = min(min(dx, dy), dz);
    } else {
        if (P.free_surface == 0)
            m = min(dx, dy);
        else if (gy < W)
            m = dx;
        else
            m = min(dx, dy);
    }
    return m < W ? __ldg(P.absCoeff + m) : 1.0f;
}

// Per-point bookkeeping shared by the per-point kernels below and the marching kernels (ws_kernels_march.cuh): indices,
// row classes of the derivative tables, CPML slab positions, ABS factor.  The derivative itself (D<F, OP>) is supplied
// by the derived point type, so both kernel families run the SAME statement sequence (passA / passB).
template <bool EXACT>
struct PtBase {
    using A = Ar<EXACT>;
    const WsParams &P;
    int x, ly, z, gy;
    long long i;    // padded linear index
    int rx, ry, rz; // derivative row classes
    int kx, ky, kz; // CPML slab indices (-1 outside)
    long long px, py, pz; // psi offsets

    __device__ __forceinline__ PtBase(const WsParams &P_, int x_, int ly_, int z_) : P(P_), x(x_), z(z_)
    {
        rx = wsRowClass(x, P.nx, P.h);
        rz = wsRowClass(z, P.nz, P.h);
        kx = kz = -1;
        if (P.damping == 2) {
            kx = wsCpmlIndex(x, P.nx, P.W);
            if (P.dim == 3)
                kz = wsCpmlIndex(z, P.nz, P.W);
        }
        setY(ly_);
    }
    // everything that depends on the plane (the marching kernels advance a point along y)
    __device__ __forceinline__ void setY(int ly_)
    {
        ly = ly_;
        gy = P.gy0 + ly;
        i = P.base + x + (long long)z * P.pitch + (long long)ly * P.plane;
        ry = wsRowClass(gy, P.gny, P.h);
        ky = -1;
        px = py = pz = 0;
        if (P.damping == 2) {
            const int W = P.W;
            ky = wsCpmlIndex(gy, P.gny, W);
            if (P.free_surface != 0 && gy < W)
                ky = -1; // no CPML in the top layer with a free surface (CPML3D.cpp:320-328)
            px = ((long long)ly * P.nz + z) * P.psiPitchX + wsPsiXIndex(x, W, P.psiDX);
            py = ((long long)ky * P.nz + z) * P.nx + x;
            pz = ((long long)ly * (2 * W) + kz) * P.nx + x;
        }
    }

    // CPML.cpp:84-95 applyCPML
    __device__ __forceinline__ float cp(float d, int slot, int k, long long off, const float *__restrict__ ca, const float *__restrict__ cb) const
    {
        if (k < 0)
            return d;
        float *ps = P.psi[slot] + off;
        float v = A::mul(*ps, __ldg(cb + k));
        const float t = A::mul(__ldg(ca + k), d);
        v = A::add(v, t);
        *ps = v;
        return A::add(d, v);
    }
    __device__ __forceinline__ float cpx(float d, int slot, bool half) const { return cp(d, slot, kx, px, half ? P.caxh : P.cax, half ? P.cbxh : P.cbx); }
    __device__ __forceinline__ float cpy(float d, int slot, bool half) const { return cp(d, slot, ky, py, half ? P.cayh : P.cay, half ? P.cbyh : P.cby); }
    __device__ __forceinline__ float cpz(float d, int slot, bool half) const { return cp(d, slot, kz, pz, half ? P.cazh : P.caz, half ? P.cbzh : P.cbz); }

    __device__ __forceinline__ float absFactor() const { return wsAbsFactor(P, x, gy, z); }
    __device__ __forceinline__ int surfaceIndex() const { return z * P.nx + x; }
};

// one thread per grid point, every tap read from global memory
template <bool EXACT>
struct Pt : PtBase<EXACT> {
    using A = Ar<EXACT>;
    using V = float;
    using PtBase<EXACT>::P;
    __device__ __forceinline__ Pt(const WsParams &P_, int x_, int ly_, int z_) : PtBase<EXACT>(P_, x_, ly_, z_) {}
    // row of matrix `op` applied to field f: ascending-column accumulation like a CSR SpMV
    __device__ __forceinline__ float D(const float *__restrict__ f, int op) const
    {
        const int axis = (op < 6) ? (op >> 1) : 1;
        const int r = axis == 0 ? this->rx : (axis == 1 ? this->ry : this->rz);
        const long long s = axis == 0 ? 1 : (axis == 1 ? P.plane : (long long)P.pitch);
        const float *__restrict__ w = P.tab + ((size_t)op * (2 * P.h + 1) + r) * (P.q + 1);
        const float *__restrict__ p = f + this->i - (long long)P.h * s;
        float acc = 0.0f;
        for (int j = 0; j <= P.q; j++) {
            acc = A::madd(__ldg(w + j), p[0], acc);
            p += s;
        }
        return acc;
    }
    template <int F, int OP> __device__ __forceinline__ float D() const { return D(P.fld[F], OP); }
    // own-point operands: wavefield F, model parameter M, memory variable C of relaxation mechanism l, EM coefficient Cd
    template <int F> __device__ __forceinline__ float fld() const { return P.fld[F][this->i]; }
    template <int F> __device__ __forceinline__ void put(float v) const { P.fld[F][this->i] = v; }
    template <int M> __device__ __forceinline__ float mat() const { return P.mat[M][this->i]; }
    template <int C> __device__ __forceinline__ float rget(int l) const { return P.fld[F_R0 + 6 * l + C][this->i]; }
    template <int C> __device__ __forceinline__ void rput(int l, float v) const { P.fld[F_R0 + 6 * l + C][this->i] = v; }
    template <int AXIS> __device__ __forceinline__ float cd(int l) const { return P.mat[M_CD0 + 3 * l + AXIS][this->i]; }
    // free-surface scalings of this column (FreeSurfaceElastic.cpp:35-46, FreeSurfaceViscoelastic.cpp:38-95)
    __device__ __forceinline__ float sH() const { return P.sH[this->surfaceIndex()]; }
    __device__ __forceinline__ float sV() const { return P.sV[this->surfaceIndex()]; }
    __device__ __forceinline__ float sRH(int l) const { return P.sRH[l][this->surfaceIndex()]; }
    __device__ __forceinline__ float sRV(int l) const { return P.sRV[l][this->surfaceIndex()]; }
};

// --------------------------------------------------------------------------------------------------------------------
// first half-step: particle velocities / magnetic field
// --------------------------------------------------------------------------------------------------------------------
template <int EQ, int DIM, bool EXACT, typename PT>
__device__ __forceinline__ void passA(const WsParams &P, const PT &t)
{
    using A = Ar<EXACT>;
    using V = typename PT::V; // float, or 4 consecutive x points (marching kernels)
    const bool fs = P.free_surface == 1;
    if (EQ == WS_EQ_ACOUSTIC) {
        // ForwardSolver3Dacoustic.cpp:131-187, ForwardSolver2Dacoustic.cpp:121-160
        V u = t.template D<F_P, OP_XF>();
        u = t.cpx(u, PSI_P_X, true);
        u = A::mul(u, t.template mat<M_RIX>());
        t.template put<F_VX>(A::add(t.template fld<F_VX>(), u));
        u = (fs ? t.template D<F_P, OP_YF_FS>() : t.template D<F_P, OP_YF>());
        u = t.cpy(u, PSI_P_Y, true);
        u = A::mul(u, t.template mat<M_RIY>());
        t.template put<F_VY>(A::add(t.template fld<F_VY>(), u));
        if (DIM == 3) {
            u = t.template D<F_P, OP_ZF>();
            u = t.cpz(u, PSI_P_Z, true);
            u = A::mul(u, t.template mat<M_RIZ>());
            t.template put<F_VZ>(A::add(t.template fld<F_VZ>(), u));
        }
    } else if (EQ == WS_EQ_ELASTIC || EQ == WS_EQ_VISCOELASTIC) {
        // ForwardSolver3Delastic.cpp:181-277, ForwardSolver2Delastic.cpp:163-208, ForwardSolver3Dviscoelastic.cpp:188-262
        V u = t.template D<F_SXX, OP_XF>();
        u = t.cpx(u, PSI_SXX_X, true);
        V w = (fs ? t.template D<F_SXY, OP_YB_FS>() : t.template D<F_SXY, OP_YB>());
        w = t.cpy(w, PSI_SXY_Y, false);
        u = A::add(u, w);
        if (DIM == 3) {
            w = t.template D<F_SXZ, OP_ZB>();
            w = t.cpz(w, PSI_SXZ_Z, false);
            u = A::add(u, w);
        }
        u = A::mul(u, t.template mat<M_RIX>());
        t.template put<F_VX>(A::add(t.template fld<F_VX>(), u));

        u = t.template D<F_SXY, OP_XB>();
        u = t.cpx(u, PSI_SXY_X, false);
        w = (fs ? t.template D<F_SYY, OP_YF_FS>() : t.template D<F_SYY, OP_YF>());
        w = t.cpy(w, PSI_SYY_Y, true);
        u = A::add(u, w);
        if (DIM == 3) {
            w = t.template D<F_SYZ, OP_ZB>();
            w = t.cpz(w, PSI_SYZ_Z, false);
            u = A::add(u, w);
        }
        u = A::mul(u, t.template mat<M_RIY>());
        t.template put<F_VY>(A::add(t.template fld<F_VY>(), u));

        if (DIM == 3) {
            u = t.template D<F_SXZ, OP_XB>();
            u = t.cpx(u, PSI_SXZ_X, false);
            w = (fs ? t.template D<F_SYZ, OP_YB_FS>() : t.template D<F_SYZ, OP_YB>());
            w = t.cpy(w, PSI_SYZ_Y, false);
            u = A::add(u, w);
            w = t.template D<F_SZZ, OP_ZF>();
            w = t.cpz(w, PSI_SZZ_Z, true);
            u = A::add(u, w);
            u = A::mul(u, t.template mat<M_RIZ>());
            t.template put<F_VZ>(A::add(t.template fld<F_VZ>(), u));
        }
    } else if (EQ == WS_EQ_SH || EQ == WS_EQ_VISCOSH) {
        // ForwardSolver2Dsh.cpp:140-158
        V u = t.template D<F_SXZ, OP_XB>();
        V w = (fs ? t.template D<F_SYZ, OP_YB_FS>() : t.template D<F_SYZ, OP_YB>());
        u = t.cpx(u, PSI_SXZ_X, false);
        w = t.cpy(w, PSI_SYZ_Y, false);
        u = A::add(u, w);
        u = A::mul(u, t.template mat<M_INVRHO>());
        t.template put<F_VZ>(A::add(t.template fld<F_VZ>(), u));
    } else if (EQ == WS_EQ_TMEM || EQ == WS_EQ_VISCOTMEM) {
        // ForwardSolver2Dtmem.cpp:131-146
        V u = t.template D<F_EZ, OP_YF>();
        u = t.cpy(u, PSI_EZY, true);
        u = A::mul(u, t.template mat<M_MIYZ>());
        t.template put<F_HX>(A::sub(t.template fld<F_HX>(), u));
        V w = t.template D<F_EZ, OP_XF>();
        w = t.cpx(w, PSI_EZX, true);
        u = A::mul(-1.0f, w);
        u = A::mul(u, t.template mat<M_MIXZ>());
        t.template put<F_HY>(A::sub(t.template fld<F_HY>(), u));
    } else if (EQ == WS_EQ_EMEM || EQ == WS_EQ_VISCOEMEM) {
        if (DIM == 3) {
            // ForwardSolver3Demem.cpp:154-187
            V u = t.template D<F_EZ, OP_YF>();
            V w = t.template D<F_EY, OP_ZF>();
            u = t.cpy(u, PSI_EZY, true);
            w = t.cpz(w, PSI_EYZ, true);
            u = A::sub(u, w);
            u = A::mul(u, t.template mat<M_MIYZ>());
            t.template put<F_HX>(A::sub(t.template fld<F_HX>(), u));
            u = t.template D<F_EX, OP_ZF>();
            w = t.template D<F_EZ, OP_XF>();
            u = t.cpz(u, PSI_EXZ, true);
            w = t.cpx(w, PSI_EZX, true);
            u = A::sub(u, w);
            u = A::mul(u, t.template mat<M_MIXZ>());
            t.template put<F_HY>(A::sub(t.template fld<F_HY>(), u));
        }
        // ForwardSolver2Demem.cpp:136-146
        V u = t.template D<F_EY, OP_XF>();
        V w = t.template D<F_EX, OP_YF>();
        u = t.cpx(u, PSI_EYX, true);
        w = t.cpy(w, PSI_EXY, true);
        u = A::sub(u, w);
        u = A::mul(u, t.template mat<M_MIXY>());
        t.template put<F_HZ>(A::sub(t.template fld<F_HZ>(), u));
    }
}

// viscoelastic helpers (ForwardSolver3Dviscoelastic.cpp:284-416) -------------------------------------------------------
template <bool EXACT, int RC, typename PT>
__device__ __forceinline__ typename PT::V viscoShear(const WsParams &P, const PT &t, typename PT::V S, typename PT::V u, typename PT::V muAvg, typename PT::V tauAvg,
                                                     typename PT::V onePlusLtauS)
{
    using A = Ar<EXACT>;
    using V = typename PT::V;
    u = A::mul(u, muAvg);
#pragma unroll 1
    for (int l = 0; l < P.L; l++) {
        V R = t.template rget<RC>(l);
        S = A::madd(P.DThalf, R, S);
        R = A::mul(R, P.viscoCoeff1[l]);
        V u2 = A::mul(P.invRelaxTime[l], u);
        u2 = A::mul(u2, tauAvg);
        R = A::sub(R, u2);
        R = A::mul(R, P.viscoCoeff2[l]);
        S = A::madd(P.DThalf, R, S);
        t.template rput<RC>(l, R);
    }
    u = A::mul(u, onePlusLtauS);
    return A::add(S, u);
}

// --------------------------------------------------------------------------------------------------------------------
// second half-step: stresses / pressure / electric field, free surface, ABS on the fields written here
// --------------------------------------------------------------------------------------------------------------------
template <int EQ, int DIM, bool EXACT, typename PT>
__device__ __forceinline__ void passB(const WsParams &P, const PT &t)
{
    using A = Ar<EXACT>;
    using V = typename PT::V;
    const bool fs = P.free_surface == 1;
    const V damp = t.absFactor();
    const bool surf = fs && t.gy == 0;
    if (EQ == WS_EQ_ACOUSTIC) {
        // ForwardSolver3Dacoustic.cpp:192-225
        V u = t.template D<F_VX, OP_XB>();
        u = t.cpx(u, PSI_VXX, false);
        V w = t.template D<F_VY, OP_YB>();
        w = t.cpy(w, PSI_VYY, false);
        u = A::add(u, w);
        if (DIM == 3) {
            w = t.template D<F_VZ, OP_ZB>();
            w = t.cpz(w, PSI_VZZ, false);
            u = A::add(u, w);
        }
        u = A::mul(u, t.template mat<M_PW>());
        V p = A::add(t.template fld<F_P>(), u);
        p = A::mul(p, damp);
        if (surf)
            p = A::mul(p, 0.0f);
        t.template put<F_P>(p);
    } else if (EQ == WS_EQ_ELASTIC || EQ == WS_EQ_VISCOELASTIC) {
        V vxx = t.template D<F_VX, OP_XB>();
        V vyy = t.template D<F_VY, OP_YB>(); // plain Dyb even with a free surface (ForwardSolver3Delastic.cpp:289)
        V vzz = 0.0f;
        if (DIM == 3)
            vzz = t.template D<F_VZ, OP_ZB>();
        vxx = t.cpx(vxx, PSI_VXX, false);
        vyy = t.cpy(vyy, PSI_VYY, false);
        if (DIM == 3)
            vzz = t.cpz(vzz, PSI_VZZ, false);
        V sxx = t.template fld<F_SXX>(), syy = t.template fld<F_SYY>(), szz = 0.0f;
        if (DIM == 3)
            szz = t.template fld<F_SZZ>();
        const V pi = t.template mat<M_PW>(), mu = t.template mat<M_MU>();
        V optp = 0.f, opts = 0.f, tauP = 0.f, tauS = 0.f;
        if (EQ == WS_EQ_ELASTIC) {
            // ForwardSolver3Delastic.cpp:297-314, ForwardSolver2Delastic.cpp:224-241
            V u = A::add(vxx, vyy);
            if (DIM == 3)
                u = A::add(u, vzz);
            u = A::mul(u, pi);
            sxx = A::add(sxx, u);
            syy = A::add(syy, u);
            if (DIM == 3)
                szz = A::add(szz, u);
            if (DIM == 3) {
                u = A::mul(A::add(vyy, vzz), mu);
                sxx = A::msub(2.0f, u, sxx);
                u = A::mul(A::add(vxx, vzz), mu);
                syy = A::msub(2.0f, u, syy);
                u = A::mul(A::add(vxx, vyy), mu);
                szz = A::msub(2.0f, u, szz);
            } else {
                u = A::mul(vyy, mu);
                sxx = A::msub(2.0f, u, sxx);
                u = A::mul(vxx, mu);
                syy = A::msub(2.0f, u, syy);
            }
        } else {
            // ForwardSolver3Dviscoelastic.cpp:279-352, ForwardSolver2Dviscoelastic.cpp:220-262
            tauP = t.template mat<M_TAUP>();
            tauS = t.template mat<M_TAUS>();
            optp = A::add(1.0f, A::mul(P.fL, tauP)); // onePlusLtauP = 1 + L*tauP (:109-112)
            opts = A::add(1.0f, A::mul(P.fL, tauS));
            V u = A::add(vxx, vyy);
            if (DIM == 3)
                u = A::add(u, vzz);
            u = A::mul(u, pi);
        #pragma unroll 1
    for (int l = 0; l < P.L; l++) {
                V u2 = A::mul(P.invRelaxTime[l], u);
                u2 = A::mul(u2, tauP);
                V r = t.template rget<RC_XX>(l);
                sxx = A::madd(P.DThalf, r, sxx);
                t.template rput<RC_XX>(l, A::sub(A::mul(r, P.viscoCoeff1[l]), u2));
                r = t.template rget<RC_YY>(l);
                syy = A::madd(P.DThalf, r, syy);
                t.template rput<RC_YY>(l, A::sub(A::mul(r, P.viscoCoeff1[l]), u2));
                if (DIM == 3) {
                    r = t.template rget<RC_ZZ>(l);
                    szz = A::madd(P.DThalf, r, szz);
                    t.template rput<RC_ZZ>(l, A::sub(A::mul(r, P.viscoCoeff1[l]), u2));
                }
            }
            u = A::mul(u, optp);
            sxx = A::add(sxx, u);
            syy = A::add(syy, u);
            if (DIM == 3)
                szz = A::add(szz, u);
            auto normalPart = [&](V S, auto rc, V e) {
                constexpr int RC = decltype(rc)::value;
                V uu = A::mul(e, mu);
                uu = A::mul(uu, 2.0f);
            #pragma unroll 1
    for (int l = 0; l < P.L; l++) {
                    V u2 = A::mul(P.invRelaxTime[l], uu);
                    u2 = A::mul(u2, tauS);
                    V R = A::add(t.template rget<RC>(l), u2);
                    R = A::mul(R, P.viscoCoeff2[l]);
                    S = A::madd(P.DThalf, R, S);
                    t.template rput<RC>(l, R);
                }
                uu = A::mul(uu, opts);
                return A::sub(S, uu);
            };
            if (DIM == 3) {
                sxx = normalPart(sxx, IC<RC_XX>{}, A::add(vyy, vzz));
                syy = normalPart(syy, IC<RC_YY>{}, A::add(vxx, vzz));
                szz = normalPart(szz, IC<RC_ZZ>{}, A::add(vxx, vyy));
            } else {
                sxx = normalPart(sxx, IC<RC_XX>{}, vyy);
                syy = normalPart(syy, IC<RC_YY>{}, vxx);
            }
        }
        // shear stresses: ForwardSolver3Delastic.cpp:331-382, ForwardSolver3Dviscoelastic.cpp:355-416
        {
            V u = t.template D<F_VX, OP_YF>();
            u = t.cpy(u, PSI_VXY, true);
            V w = t.template D<F_VY, OP_XF>();
            w = t.cpx(w, PSI_VYX, true);
            u = A::add(u, w);
            V s = t.template fld<F_SXY>();
            if (EQ == WS_EQ_ELASTIC)
                s = A::add(s, A::mul(u, t.template mat<M_MUXY>()));
            else
                s = viscoShear<EXACT, RC_XY>(P, t, s, u, t.template mat<M_MUXY>(), t.template mat<M_TSXY>(), opts);
            t.template put<F_SXY>(A::mul(s, damp));
        }
        if (DIM == 3) {
            V u = t.template D<F_VX, OP_ZF>();
            u = t.cpz(u, PSI_VXZ, true);
            V w = t.template D<F_VZ, OP_XF>();
            w = t.cpx(w, PSI_VZX, true);
            u = A::add(u, w);
            V s = t.template fld<F_SXZ>();
            if (EQ == WS_EQ_ELASTIC)
                s = A::add(s, A::mul(u, t.template mat<M_MUXZ>()));
            else
                s = viscoShear<EXACT, RC_XZ>(P, t, s, u, t.template mat<M_MUXZ>(), t.template mat<M_TSXZ>(), opts);
            t.template put<F_SXZ>(A::mul(s, damp));

            u = t.template D<F_VY, OP_ZF>();
            u = t.cpz(u, PSI_VYZ, true);
            w = t.template D<F_VZ, OP_YF>();
            w = t.cpy(w, PSI_VZY, true);
            u = A::add(u, w);
            s = t.template fld<F_SYZ>();
            if (EQ == WS_EQ_ELASTIC)
                s = A::add(s, A::mul(u, t.template mat<M_MUYZ>()));
            else
                s = viscoShear<EXACT, RC_YZ>(P, t, s, u, t.template mat<M_MUYZ>(), t.template mat<M_TSYZ>(), opts);
            t.template put<F_SYZ>(A::mul(s, damp));
        }
        if (surf) {
            const V hor = DIM == 3 ? A::add(vxx, vzz) : vxx;
            if (EQ == WS_EQ_ELASTIC) {
                // FreeSurface3Delastic.cpp:15-47, FreeSurface2Delastic.cpp:14-46, FreeSurface.cpp:13-20
                V tmp = A::mul(t.sH(), hor);
                sxx = A::add(sxx, tmp);
                if (DIM == 3)
                    szz = A::add(szz, tmp);
                tmp = A::mul(t.sV(), vyy);
                sxx = A::sub(sxx, tmp);
                if (DIM == 3)
                    szz = A::sub(szz, tmp);
                syy = A::mul(syy, 0.0f);
            } else {
                // FreeSurface3Dviscoelastic.cpp:17-75, FreeSurface2Dviscoelastic.cpp:15-63
            #pragma unroll 1
    for (int l = 0; l < P.L; l++) {
                    sxx = A::msub(P.DThalf, A::mul(1.0f, t.template rget<RC_XX>(l)), sxx);
                    if (DIM == 3)
                        szz = A::msub(P.DThalf, A::mul(1.0f, t.template rget<RC_ZZ>(l)), szz);
                }
                V tmp = A::mul(t.sH(), hor);
                sxx = A::add(sxx, tmp);
                if (DIM == 3)
                    szz = A::add(szz, tmp);
                tmp = A::mul(t.sV(), vyy);
                sxx = A::sub(sxx, tmp);
                if (DIM == 3)
                    szz = A::sub(szz, tmp);
            #pragma unroll 1
    for (int l = 0; l < P.L; l++) {
                    const V th = A::mul(t.sRH(l), hor);
                    const V tv = A::mul(t.sRV(l), vyy);
                    V R = A::sub(A::add(t.template rget<RC_XX>(l), th), tv);
                    t.template rput<RC_XX>(l, R);
                    sxx = A::madd(P.DThalf, A::mul(1.0f, R), sxx);
                    if (DIM == 3) {
                        R = A::sub(A::add(t.template rget<RC_ZZ>(l), th), tv);
                        t.template rput<RC_ZZ>(l, R);
                        szz = A::madd(P.DThalf, A::mul(1.0f, R), szz);
                    }
                    t.template rput<RC_YY>(l, A::mul(t.template rget<RC_YY>(l), 0.0f));
                }
                syy = A::mul(syy, 0.0f);
            }
        }
        t.template put<F_SXX>(A::mul(sxx, damp));
        t.template put<F_SYY>(A::mul(syy, damp));
        if (DIM == 3)
            t.template put<F_SZZ>(A::mul(szz, damp));
    } else if (EQ == WS_EQ_SH || EQ == WS_EQ_VISCOSH) {
        // ForwardSolver2Dsh.cpp:162-192, ForwardSolver2Dviscosh.cpp:190-236
        V opts = 0.f;
        if (EQ == WS_EQ_VISCOSH)
            opts = A::add(1.0f, A::mul(P.fL, t.template mat<M_TAUS>()));
        V u = t.template D<F_VZ, OP_XF>();
        u = t.cpx(u, PSI_VZX, true);
        V s = t.template fld<F_SXZ>();
        if (EQ == WS_EQ_SH)
            s = A::add(s, A::mul(u, t.template mat<M_MUXZ>()));
        else
            s = viscoShear<EXACT, RC_XZ>(P, t, s, u, t.template mat<M_MUXZ>(), t.template mat<M_TSXZ>(), opts);
        t.template put<F_SXZ>(A::mul(s, damp));
        u = t.template D<F_VZ, OP_YF>();
        u = t.cpy(u, PSI_VZY, true);
        s = t.template fld<F_SYZ>();
        if (EQ == WS_EQ_SH)
            s = A::add(s, A::mul(u, t.template mat<M_MUYZ>()));
        else
            s = viscoShear<EXACT, RC_YZ>(P, t, s, u, t.template mat<M_MUYZ>(), t.template mat<M_TSYZ>(), opts);
        t.template put<F_SYZ>(A::mul(s, damp));
    } else {
        // EM: r_l = Cc_l r_l + Cd_l e ;  e = Ca e + Cb (curl - DT sum r_l)
        auto updateE = [&](auto fslot, auto axisTag, V curl) {
            constexpr int FS = decltype(fslot)::value, AXIS = decltype(axisTag)::value;
            V e = t.template fld<FS>();
        #pragma unroll 1
    for (int l = 0; l < P.L; l++) {
                const V a = A::mul(P.Cc[l], t.template rget<AXIS>(l));
                const V b = A::mul(t.template cd<AXIS>(l), e);
                t.template rput<AXIS>(l, A::add(b, a));
            }
        #pragma unroll 1
    for (int l = 0; l < P.L; l++)
                curl = A::msub(P.DT, t.template rget<AXIS>(l), curl);
            curl = A::mul(curl, t.template mat<M_CBX + AXIS>());
            const V ca = A::mul(t.template mat<M_CAX + AXIS>(), e);
            e = A::add(ca, curl);
            t.template put<FS>(A::mul(e, damp));
        };
        if (EQ == WS_EQ_TMEM || EQ == WS_EQ_VISCOTMEM) {
            // ForwardSolver2Dtmem.cpp:148-163, ForwardSolver2Dviscotmem.cpp:176-197
            V u = t.template D<F_HY, OP_XB>();
            V w = t.template D<F_HX, OP_YB>();
            u = t.cpx(u, PSI_HYX, false);
            w = t.cpy(w, PSI_HXY, false);
            u = A::sub(u, w);
            updateE(IC<F_EZ>{}, IC<RC_Z>{}, u);
        } else {
            if (DIM == 3) {
                // ForwardSolver3Demem.cpp:189-232
                V u = t.template D<F_HZ, OP_YB>();
                V w = t.template D<F_HY, OP_ZB>();
                u = t.cpy(u, PSI_HZY, false);
                w = t.cpz(w, PSI_HYZ, false);
                u = A::sub(u, w);
                updateE(IC<F_EX>{}, IC<RC_X>{}, u);
                u = t.template D<F_HX, OP_ZB>();
                w = t.template D<F_HZ, OP_XB>();
                u = t.cpz(u, PSI_HXZ, false);
                w = t.cpx(w, PSI_HZX, false);
                u = A::sub(u, w);
                updateE(IC<F_EY>{}, IC<RC_Y>{}, u);
                u = t.template D<F_HY, OP_XB>();
                w = t.template D<F_HX, OP_YB>();
                u = t.cpx(u, PSI_HYX, true); // half profile: CPMLEM3D.cpp:69
                w = t.cpy(w, PSI_HXY, false);
                u = A::sub(u, w);
                updateE(IC<F_EZ>{}, IC<RC_Z>{}, u);
            } else {
                // ForwardSolver2Demem.cpp:148-169, ForwardSolver2Dviscoemem.cpp:195-222
                V u = t.template D<F_HZ, OP_YB>();
                u = t.cpy(u, PSI_HZY, false);
                updateE(IC<F_EX>{}, IC<RC_X>{}, u);
                V w = t.template D<F_HZ, OP_XB>();
                w = t.cpx(w, PSI_HZX, false);
                u = A::mul(-1.0f, w);
                updateE(IC<F_EY>{}, IC<RC_Y>{}, u);
            }
        }
    }
}

template <int EQ, int DIM, bool EXACT, int PASS>
__global__ void __launch_bounds__(256) kGeneral(const __grid_constant__ WsParams P)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = WS_POINT_Z(P.nz);
    const int ly = P.ylo + WS_POINT_PLANE(P.nz);
    if (x >= P.nx || z >= P.nz || ly >= P.yhi)
        return;
    Pt<EXACT> t(P, x, ly, z);
    if (PASS == 0)
        passA<EQ, DIM, EXACT>(P, t);
    else
        passB<EQ, DIM, EXACT>(P, t);
}

// ABS on the fields of the first half-step (they are read as neighbours by the second half-step, so they are damped
// afterwards, only inside the frame): ABS3D.cpp:15-90 apply(...)
template <bool EXACT>
__global__ void __launch_bounds__(256) kAbsFirstHalf(const __grid_constant__ WsParams P, int f0, int f1, int f2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = WS_POINT_Z(P.nz);
    const int ly = P.ylo + WS_POINT_PLANE(P.nz);
    if (x >= P.nx || z >= P.nz || ly >= P.yhi)
        return;
    Pt<EXACT> t(P, x, ly, z);
    const float d = t.absFactor();
    if (d == 1.0f)
        return;
    const int fs[3] = {f0, f1, f2};
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (fs[k] >= 0)
            P.fld[fs[k]][t.i] = Ar<EXACT>::mul(P.fld[fs[k]][t.i], d);
}

// snapType 3: energy measures of the rotational and the divergent part of the first-half-step fields after Dougherty
// and Stephen (1988), reference statement order.  Seismic (Wavefields3Delastic.cpp:197-245, Wavefields2Delastic.cpp:217-248):
// particle velocities, curl scaled by the S-wave modulus, div by the P-wave modulus.  EM (WavefieldsEM/Wavefields2Dtmem.cpp,
// Wavefields3Demem.cpp, Wavefields3Dviscoemem.cpp getCurl / getDiv and the arguments of their write()): magnetic field,
// curl scaled by the dielectric permittivity, div by the EM velocity 1/sqrt(eps mu) (divCoef = 0) or, in the 3-D visco
// case, by the conductivity (divCoef = 1).  which = 0: curl, 1: div.  `out` is a padded array like the wavefields.
template <int DIM, bool EM>
__global__ void __launch_bounds__(256) kDivCurl(const __grid_constant__ WsParams P, float *out, int which, int divCoef)
{
    using A = Ar<true>;
    constexpr int FX = EM ? F_HX : F_VX, FY = EM ? F_HY : F_VY, FZ = EM ? F_HZ : F_VZ;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = WS_POINT_Z(P.nz);
    const int ly = P.ylo + WS_POINT_PLANE(P.nz);
    if (x >= P.nx || z >= P.nz || ly >= P.yhi)
        return;
    Pt<true> t(P, x, ly, z);
    float cDiv, cCurl;
    if (EM) {
        cCurl = t.template mat<M_EPS>();
        if (divCoef == 1)
            cDiv = t.template mat<M_SIG>();
        else { // ModelparameterEM.cpp:319-328 calcVelocityFromModulus (Inf / NaN -> 0)
            cDiv = A::div(1.0f, sqrtf(A::mul(t.template mat<M_EPS>(), t.template mat<M_MUM>())));
            if (isnan(cDiv) || isinf(cDiv))
                cDiv = 0.0f;
        }
    } else {
        cCurl = t.template mat<M_MU>();
        cDiv = t.template mat<M_PW>();
    }
    float r;
    if (which == 1) {
        r = t.template D<FX, OP_XB>();
        r = A::add(r, t.template D<FY, OP_YB>());
        if (DIM == 3) {
            r = A::add(r, t.template D<FZ, OP_ZB>());
            r = A::mul(r, r);
            r = A::mul(r, cDiv);
            r = sqrtf(r);
        } else
            r = A::mul(r, sqrtf(cDiv));
    } else if (DIM == 3) {
        float u = t.template D<FZ, OP_YF>();
        u = A::sub(u, t.template D<FY, OP_ZF>());
        r = A::mul(u, u);
        u = t.template D<FX, OP_ZF>();
        u = A::sub(u, t.template D<FZ, OP_XF>());
        r = A::add(r, A::mul(u, u));
        u = t.template D<FY, OP_XF>();
        u = A::sub(u, t.template D<FX, OP_YF>());
        r = A::add(r, A::mul(u, u));
        r = A::mul(r, cCurl);
        r = sqrtf(r);
    } else {
        r = t.template D<FX, OP_YF>();
        r = A::sub(r, t.template D<FY, OP_XF>());
        r = A::mul(r, sqrtf(cCurl));
    }
    out[t.i] = r;
}

} // namespace wsgen

// wave_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A literal CPU restatement of the staggered-grid FD time-stepping path of WAVE-Simulation, in the reference's own
// formulation: every spatial derivative is an explicit CSR sparse matrix assembled by the rules of
// `src/ForwardSolver/Derivatives/Derivatives.cpp` (calcDxf/calcDyf/.../calcDybFreeSurface), and every line of the
// reference `run()` functions is one sparse-matrix-vector product or one full-length vector operation, in the same
// order and with the same rounding points (compiled with -ffp-contract=off; fp32 by default, fp64 on request).
// It therefore doubles as the "LAMA-equivalent" CPU timing baseline (same algorithm and same memory traffic as the
// LAMA host back-end; it is a restatement, not the LAMA binary: LAMA/SCAI is not available in this environment).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// The product (wave-simulation_b200/) never links, imports or calls it.
//
// Parity pin: checked against the reference's golden seismograms par/ci/seismogram.*.ref.*.mtx
// (tests/test_oracle_golden.py; fixtures copied to tests/golden/).
//
// All citations `File.cpp:N` are relative to /root/reference/src/.
#include "../include/wavesim.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Idx = int32_t;
constexpr int WS_MAXQ_ORACLE = 12; // Derivatives.cpp:2001-2042: orders 2..12
using std::vector;

thread_local std::string g_err;

#define ORACLE_REQUIRE(cond, msg)                                                                                      \
    do {                                                                                                               \
        if (!(cond))                                                                                                   \
            throw std::runtime_error(std::string(msg));                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// FD coefficients: Taylor coefficients of the staggered first derivative, orders 2..12
// (Derivatives.cpp:2001-2042 setFDCoef; the reference stores them as ValueType).
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
vector<T> fdCoef(int q)
{
    static const double c2[] = {-1.0, 1.0};
    static const double c4[] = {1.0 / 24.0, -9.0 / 8.0, 9.0 / 8.0, -1.0 / 24.0};
    static const double c6[] = {-3.0 / 640.0, 25.0 / 384.0, -75.0 / 64.0, 75.0 / 64.0, -25.0 / 384.0, 3.0 / 640.0};
    static const double c8[] = {5.0 / 7168.0, -49.0 / 5120.0, 245.0 / 3072.0, -1225.0 / 1024.0,
                                1225.0 / 1024.0, -245.0 / 3072.0, 49.0 / 5120.0, -5.0 / 7168.0};
    static const double c10[] = {-35.0 / 294912.0, 405.0 / 229376.0, -567.0 / 40960.0, 735.0 / 8192.0, -19845.0 / 16384.0,
                                 19845.0 / 16384.0, -735.0 / 8192.0, 567.0 / 40960.0, -405.0 / 229376.0, 35.0 / 294912.0};
    static const double c12[] = {63.0 / 2883584.0, -847.0 / 2359296.0, 5445.0 / 1835008.0, -22869.0 / 1310720.0,
                                 12705.0 / 131072.0, -160083.0 / 131072.0, 160083.0 / 131072.0, -12705.0 / 131072.0,
                                 22869.0 / 1310720.0, -5445.0 / 1835008.0, 847.0 / 2359296.0, -63.0 / 2883584.0};
    const double *p = nullptr;
    switch (q) {
    case 2: p = c2; break;
    case 4: p = c4; break;
    case 6: p = c6; break;
    case 8: p = c8; break;
    case 10: p = c10; break;
    case 12: p = c12; break;
    default: throw std::runtime_error("spatialFDorder = " + std::to_string(q) + " Unsupported spatialFDorder value.");
    }
    vector<T> r(q);
    for (int j = 0; j < q; j++)
        r[j] = (T)p[j];
    return r;
}

// ------------------------------------------------------------------------------------------------------------------
// CSR matrix + SpMV (stand-in for lama::CSRSparseMatrix / StencilMatrix times DenseVector)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Csr {
    Idx n = 0;
    vector<int64_t> ia;
    vector<Idx> ja;
    vector<T> va;
    bool empty() const { return n == 0; }
};

constexpr int MAXROW = 16;

// rowfn(i, cols, vals) -> nnz of row i (unsorted); rows are sorted by column like LAMA's fillFromAssembly
template <typename T, typename F>
Csr<T> assemble(Idx n, F rowfn)
{
    vector<Idx> cols((size_t)n * MAXROW);
    vector<T> vals((size_t)n * MAXROW);
    vector<int> cnt(n);
#pragma omp parallel for schedule(static)
    for (Idx i = 0; i < n; i++) {
        Idx *c = &cols[(size_t)i * MAXROW];
        T *v = &vals[(size_t)i * MAXROW];
        int k = rowfn(i, c, v);
        // insertion sort by column
        for (int a = 1; a < k; a++) {
            Idx cc = c[a];
            T vv = v[a];
            int b = a - 1;
            while (b >= 0 && c[b] > cc) {
                c[b + 1] = c[b];
                v[b + 1] = v[b];
                b--;
            }
            c[b + 1] = cc;
            v[b + 1] = vv;
        }
        cnt[i] = k;
    }
    Csr<T> A;
    A.n = n;
    A.ia.resize((size_t)n + 1);
    A.ia[0] = 0;
    for (Idx i = 0; i < n; i++)
        A.ia[i + 1] = A.ia[i] + cnt[i];
    A.ja.resize(A.ia[n]);
    A.va.resize(A.ia[n]);
#pragma omp parallel for schedule(static)
    for (Idx i = 0; i < n; i++) {
        for (int k = 0; k < cnt[i]; k++) {
            A.ja[A.ia[i] + k] = cols[(size_t)i * MAXROW + k];
            A.va[A.ia[i] + k] = vals[(size_t)i * MAXROW + k];
        }
    }
    return A;
}

template <typename T>
void scaleCsr(Csr<T> &A, T s)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)A.va.size(); k++)
        A.va[k] = A.va[k] * s;
}

// y = A * x  (row sum in ascending column order, accumulator starts at 0)
template <typename T>
void spmv(const Csr<T> &A, const vector<T> &x, vector<T> &y)
{
    y.resize(A.n);
    const int64_t *ia = A.ia.data();
    const Idx *ja = A.ja.data();
    const T *va = A.va.data();
    const T *xp = x.data();
    T *yp = y.data();
#pragma omp parallel for schedule(static)
    for (Idx i = 0; i < A.n; i++) {
        T s = 0;
        for (int64_t k = ia[i]; k < ia[i + 1]; k++)
            s += va[k] * xp[ja[k]];
        yp[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// full-length vector operations (one per reference statement)
// ------------------------------------------------------------------------------------------------------------------
#define VFOR(n) _Pragma("omp parallel for schedule(static)") for (Idx i = 0; i < (Idx)(n); i++)

template <typename T> void vadd(vector<T> &a, const vector<T> &b) { T *p = a.data(); const T *q = b.data(); VFOR(a.size()) p[i] = p[i] + q[i]; }
template <typename T> void vsub(vector<T> &a, const vector<T> &b) { T *p = a.data(); const T *q = b.data(); VFOR(a.size()) p[i] = p[i] - q[i]; }
template <typename T> void vmul(vector<T> &a, const vector<T> &b) { T *p = a.data(); const T *q = b.data(); VFOR(a.size()) p[i] = p[i] * q[i]; }
template <typename T> void vscale(vector<T> &a, T s) { T *p = a.data(); VFOR(a.size()) p[i] = p[i] * s; }
template <typename T> void vset(vector<T> &a, const vector<T> &b) { a = b; }
// a = b + c
template <typename T> void vsum(vector<T> &a, const vector<T> &b, const vector<T> &c) { a.resize(b.size()); T *p = a.data(); const T *q = b.data(); const T *r = c.data(); VFOR(b.size()) p[i] = q[i] + r[i]; }
// a = s * b
template <typename T> void vscaled(vector<T> &a, T s, const vector<T> &b) { a.resize(b.size()); T *p = a.data(); const T *q = b.data(); VFOR(b.size()) p[i] = s * q[i]; }
// a += s * b   (a[i] = a[i] + (s*b[i]), two roundings)
template <typename T> void vaxpy(vector<T> &a, T s, const vector<T> &b) { T *p = a.data(); const T *q = b.data(); VFOR(a.size()) p[i] = p[i] + s * q[i]; }
// a -= s * b
template <typename T> void vaxmy(vector<T> &a, T s, const vector<T> &b) { T *p = a.data(); const T *q = b.data(); VFOR(a.size()) p[i] = p[i] - s * q[i]; }

// Common.hpp:128-141 replaceInvalid, Common.hpp:68-120 searchAndReplace (compareType 1 := <)
template <typename T> void replaceInvalid(vector<T> &a, T v)
{
    T *p = a.data();
    VFOR(a.size()) if (std::isnan(p[i]) || std::isinf(p[i])) p[i] = v;
}
template <typename T> void searchAndReplaceLess(vector<T> &a, T thr, T v)
{
    T *p = a.data();
    VFOR(a.size()) if (p[i] < thr) p[i] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// sparse vector (boundary-only) for CPML
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Profile { // pattern + (a,b) full-grid and half-grid coefficient values on that pattern
    vector<Idx> idx;
    vector<T> a, b, ah, bh;
};

// ------------------------------------------------------------------------------------------------------------------
// Variable grid (Acquisition/Coordinates.cpp:115-247, 383-407, 464-536, 623-694): horizontal layers whose grid spacing is
// dhFactor = 3^n times the finest one; an interface plane is stored with the spacing of its FINER neighbour.  Coordinates
// are always fine-grid coordinates; the model vector is the concatenation of the layers' own regular grids.
// ------------------------------------------------------------------------------------------------------------------
struct VarGrid {
    bool active = false;
    bool variableSpacing = false; // useVariableGrid (some dhFactor > 1); false = variable FD order on a regular grid
    int numLayers = 1;
    Idx NX = 0, NY = 0, NZ = 0, n = 0;
    vector<int> iface;      // interface[0] = -1 ... interface[numLayers] = NY - 1
    vector<int> dhFactor, transition, layerStart, layerEnd, varNX, varNY, varNZ, fdOrder;
    vector<Idx> nPerLayer;

    // Coordinates.cpp:115-247 (init(dhFactors, interfaces)); interfaces = the file column without its leading 0
    void init(Idx nx, Idx ny, Idx nz, const vector<int> &dhf, vector<int> interfaces)
    {
        NX = nx; NY = ny; NZ = nz;
        dhFactor = dhf;
        numLayers = (int)dhFactor.size();
        ORACLE_REQUIRE(numLayers > 0 && (int)interfaces.size() == numLayers - 1, "number of interfaces doesn't match to the number of different grid spacings");
        iface = interfaces;
        iface.push_back((int)NY - 1);
        iface.insert(iface.begin(), -1);
        int dhMax = 0;
        for (int l = 0; l < numLayers; l++) {
            int f = dhFactor[l];
            ORACLE_REQUIRE(f >= 1, "incompatible dhFactor, dhFactor must be 3^n");
            while (f % 3 == 0)
                f /= 3;
            ORACLE_REQUIRE(f == 1, "incompatible dhFactor, dhFactor must be 3^n");
            dhMax = std::max(dhMax, dhFactor[l]);
        }
        variableSpacing = dhMax > 1;
        Idx NXmax = NX, NZmax = NZ;
        if (dhMax != 1) {
            while (NXmax != (NXmax / dhMax) * dhMax + 1 + dhMax / 2)
                NXmax--;
            while (NZ != 1 && NZmax != (NZmax / dhMax) * dhMax + 1 + dhMax / 2)
                NZmax--;
            NX = NXmax;
            NZ = NZmax;
            // the layer thicknesses must be multiples of the layer's spacing (:178-197): interfaces move up until they are
            int layer = 0;
            while (layer < 1) {
                layer++;
                if ((iface[layer] - iface[layer - 1] - 1) % dhFactor[layer - 1] != 0) {
                    iface[layer]--;
                    layer--;
                }
            }
            layer = 1;
            while (layer < numLayers) {
                layer++;
                if ((iface[layer] - iface[layer - 1]) % dhFactor[layer - 1] != 0) {
                    iface[layer]--;
                    layer--;
                }
            }
        }
        NY = iface[numLayers] + 1;
        transition.assign(numLayers, 0);
        layerStart.assign(numLayers, 0);
        layerEnd.assign(numLayers, 0);
        for (int l = 0; l < numLayers - 1; l++) {
            if (dhFactor[l] < dhFactor[l + 1]) {
                transition[l] = 1;
                layerEnd[l] = iface[l + 1];
                layerStart[l + 1] = iface[l + 1] + dhFactor[l + 1];
            } else if (dhFactor[l] > dhFactor[l + 1]) {
                transition[l] = -1;
                layerEnd[l] = iface[l + 1] - dhFactor[l];
                layerStart[l + 1] = iface[l + 1];
            } else {
                transition[l] = 0;
                layerEnd[l] = iface[l + 1] - dhFactor[l];
                layerStart[l + 1] = iface[l + 1];
            }
        }
        transition[numLayers - 1] = 0;
        layerEnd[numLayers - 1] = iface[numLayers];
        varNX.assign(numLayers, 0); varNY.assign(numLayers, 0); varNZ.assign(numLayers, 0);
        nPerLayer.assign(numLayers, 0);
        n = 0;
        for (int l = 0; l < numLayers; l++) {
            varNY[l] = (layerEnd[l] - layerStart[l]) / dhFactor[l] + 1;
            varNX[l] = (int)(NXmax / dhFactor[l]);
            varNZ[l] = (int)(NZmax / dhFactor[l]);
            if (dhFactor[l] > 1) {
                varNX[l]++;
                varNZ[l]++;
            }
            nPerLayer[l] = (Idx)varNX[l] * varNY[l] * varNZ[l];
            n += nPerLayer[l];
        }
        active = true;
    }
    // Coordinates.cpp:383-398: the coarse grids own the interface of a fine -> coarse transition
    int getLayer(Idx y) const
    {
        int layer = 0;
        for (layer = 0; layer < numLayers; layer++) {
            if ((int)y < iface[layer + 1] && (int)y > iface[layer])
                break;
            else if ((int)y == iface[layer + 1]) {
                if (transition.at(layer) > 0)
                    layer += 1;
                break;
            }
        }
        return layer;
    }
    bool onInterface(Idx y) const // :474-483 (interface[numLayers] = NY - 1 is not tested, interface[0] = -1 never matches)
    {
        for (int l = 0; l < numLayers; l++)
            if ((int)y == iface[l])
                return true;
        return false;
    }
    Idx distToInterface(Idx y) const // :490-499
    {
        Idx dist = NY;
        for (int k = 1; k < numLayers; k++)
            dist = std::min<Idx>(dist, std::abs((int)y - iface[k]));
        return dist;
    }
    int getTransition(Idx y) const // :517-530
    {
        ORACLE_REQUIRE(onInterface(y), "Y Coordinate is not located on an variable grid interface");
        int t = 0;
        for (int l = 0; l < numLayers; l++)
            if ((int)y == iface[l + 1])
                t = transition[l];
        return t;
    }
    int factorAt(Idx y) const { return dhFactor[getLayer(y)]; }
    void index2coord(Idx index, Idx &x, Idx &y, Idx &z) const // :623-651
    {
        int layer;
        for (layer = 0; layer < numLayers; layer++) {
            if (index >= nPerLayer[layer])
                index -= nPerLayer[layer];
            else
                break;
        }
        y = index / ((Idx)varNX[layer] * varNZ[layer]);
        index -= y * ((Idx)varNX[layer] * varNZ[layer]);
        z = index / varNX[layer];
        index -= z * varNX[layer];
        x = index;
        x *= dhFactor[layer];
        y *= dhFactor[layer];
        z *= dhFactor[layer];
        y += layerStart[layer];
    }
    Idx coord2index(Idx X, Idx Y, Idx Z) const // :668-694
    {
        ORACLE_REQUIRE(X >= 0 && X < NX && Y >= 0 && Y < NY && Z >= 0 && Z < NZ, "Could not map from coordinate to index!");
        int layer;
        for (layer = 0; layer < numLayers; layer++)
            if (Y <= layerEnd[layer] && Y >= layerStart[layer]) {
                Y -= layerStart[layer];
                break;
            }
        ORACLE_REQUIRE(layer < numLayers, "Could not map from coordinate to index!");
        Idx index = (X / dhFactor[layer]) + (Z / dhFactor[layer]) * varNX[layer] + (Y / dhFactor[layer]) * (Idx)varNX[layer] * varNZ[layer];
        for (int l = 1; l <= layer; l++)
            index += nPerLayer[l - 1];
        return index;
    }
};

template <typename T>
struct Oracle {
    ws_desc d{};
    Idx NX = 0, NY = 0, NZ = 0, N = 0;
    bool prepared = false;
    T DT = 0, DH = 0;
    int L = 0;

    std::map<std::string, vector<T>> mat; // model parameters, raw + derived, by reference getter name
    std::map<std::string, vector<T>> fld; // wavefields by reference component name

    Csr<T> Dxf, Dxb, Dyf, Dyb, Dzf, Dzb, DyfFS, DybFS;
    VarGrid vg;                                   // variable grid (inactive on a regular grid)
    Csr<T> InterFull, InterStagX, InterStagZ;     // interpolation on the interface planes (Derivatives.cpp:1252-1566)

    // boundaries
    vector<T> damping;                 // ABS (dense, default 1.0)
    vector<Idx> surfIdx;               // indices with y == 0
    vector<T> sH, sV;                  // FreeSurfaceElastic: scaleHorizontalUpdate / scaleVerticalUpdate on surfIdx
    vector<vector<T>> sRH, sRV;        // FreeSurfaceViscoelastic relaxation scalings
    Profile<T> px, py, pz;             // CPML coefficient patterns per axis
    std::map<std::string, vector<T>> psi; // CPML memory variables on the axis pattern

    // visco
    vector<T> relaxationTime, inverseRelaxationTime, viscoCoeff1, viscoCoeff2;
    T DThalf = 0;
    vector<T> onePlusLtauP, onePlusLtauS;
    // EM
    vector<T> Cc;

    // acquisition
    vector<Idx> srcType, srcIdx, recType, recIdx;
    vector<T> srcSig; // nsrc x nt
    vector<T> seis;   // nrec x nt

    // temporaries
    vector<T> update, update_temp, update2, vxx, vyy, vzz;

    Idx index(Idx x, Idx y, Idx z) const { return vg.active ? vg.coord2index(x, y, z) : x + z * NX + y * NX * NZ; } // Coordinates.cpp:687
    void coord(Idx i, Idx &x, Idx &y, Idx &z) const
    { // Coordinates.cpp:615-645
        if (vg.active) {
            vg.index2coord(i, x, y, z);
            return;
        }
        y = i / (NX * NZ);
        i -= y * (NX * NZ);
        z = i / NX;
        i -= z * NX;
        x = i;
    }

    vector<T> &M(const std::string &k)
    {
        auto it = mat.find(k);
        ORACLE_REQUIRE(it != mat.end(), "model parameter '" + k + "' is not set");
        return it->second;
    }
    vector<T> &F(const std::string &k)
    {
        auto it = fld.find(k);
        ORACLE_REQUIRE(it != fld.end(), "wavefield '" + k + "' does not exist");
        return it->second;
    }
    bool isEq(int e) const { return d.eq == e; }
    bool seismic() const { return d.eq <= WS_EQ_VISCOSH; }
    bool visco() const { return d.eq == WS_EQ_VISCOELASTIC || d.eq == WS_EQ_VISCOSH || d.eq == WS_EQ_VISCOTMEM || d.eq == WS_EQ_VISCOEMEM; }

    // ------------------------------------------------------------------------------------------------------------
    // Derivative matrices
    // ------------------------------------------------------------------------------------------------------------
    // axis: 0 = x, 1 = y, 2 = z
    Idx axisN(int axis) const { return axis == 0 ? NX : (axis == 1 ? NY : NZ); }
    Idx axisStride(int axis) const { return axis == 0 ? 1 : (axis == 1 ? NX * NZ : NX); }
    Idx axisCoord(Idx i, int axis) const
    {
        Idx x, y, z;
        coord(i, x, y, z);
        return axis == 0 ? x : (axis == 1 ? y : z);
    }

    // Sparse assembly with order reduction at the domain edges:
    // forward: Derivatives.cpp:129-186 (x), :210-280 (y), :304-359 (z); backward: :706-763 (x), :771-843 (y), calcDzb.
    // Values: coefficient / DH, later scaled by DT (FDTD3D.cpp:252-256, FDTD2D.cpp) -> fl(fl(c/DH)*DT).
    Csr<T> derivSparse(int axis, bool forward)
    {
        const int q = d.fd_order;
        std::map<int, vector<T>> fdmap;
        for (int o = 2; o <= q; o += 2)
            fdmap[o] = fdCoef<T>(o);
        const Idx n = axisN(axis), st = axisStride(axis);
        const T dh = DH;
        Csr<T> A = assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx c = axisCoord(row, axis);
            const Idx c0 = c;
            int order = q;
            int k = 0;
            for (int j = 0; j < order; j++) {
                Idx X, Xmin, Xmax;
                if (forward) {
                    X = c + (j - order / 2 + 1);
                    Xmin = c + (-order / 2 + 1);
                    Xmax = c + order / 2;
                } else {
                    X = c + (j - order / 2);
                    Xmin = c + (-order / 2);
                    Xmax = c + (order / 2 - 1);
                }
                if (Xmin < 0) {
                    order += 2 * Xmin;
                    if (!forward && order == 0) { // Derivatives.cpp:745-748
                        order = 2;
                        c += 1;
                    }
                    j--;
                } else if (Xmax >= n) {
                    order -= 2 * (Xmax - n + 1);
                    if (forward && order == 0) { // Derivatives.cpp:171-174
                        order = 2;
                        c -= 1;
                    }
                    j--;
                } else {
                    cols[k] = row + (X - c0) * st;
                    vals[k] = fdmap[order][j] / dh;
                    k++;
                }
            }
            return k;
        });
        scaleCsr(A, DT);
        return A;
    }

    // StencilMatrix path (useStencilMatrix=1): Derivatives.cpp:112-121,194-201,288-295; off-grid taps are dropped;
    // backward = -(forward)^T (FDTD3D.cpp:202-207); values scaled by DT/DH (FDTD3D.cpp:211-216).
    Csr<T> derivStencil(int axis, bool forward)
    {
        const int q = d.fd_order;
        const vector<T> fd = fdCoef<T>(q);
        const Idx n = axisN(axis), st = axisStride(axis);
        const T s = DT / DH;
        return assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            const Idx c = axisCoord(row, axis);
            int k = 0;
            for (int j = 0; j < q; j++) {
                Idx off = forward ? (j - q / 2 + 1) : (j - q / 2);
                Idx X = c + off;
                if (X < 0 || X >= n)
                    continue;
                cols[k] = row + off * st;
                vals[k] = fd[j] * s;
                k++;
            }
            return k;
        });
    }

    // Image-method matrices, Derivatives.cpp:367-440 (calcDyfFreeSurface), :448-526 (calcDybFreeSurface);
    // values (c - c_image)/DH then *= DT (FDTD3D.cpp:312-323). No order reduction; off-grid rows dropped.
    Csr<T> derivFreeSurface(bool forward)
    {
        const int q = d.fd_order;
        const vector<T> fd = fdCoef<T>(q);
        const Idx st = NX * NZ;
        const T dh = DH;
        Csr<T> A = assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            const Idx y = axisCoord(row, 1);
            int k = 0;
            for (int j = 0; j < q; j++) {
                Idx Y = forward ? y + (j - q / 2 + 1) : y + (j - q / 2);
                T fdCoeff = fd[j];
                T diffCoeff = 0;
                if (forward) {
                    if (q >= (2 + 2 * y + j)) {
                        int im = q - 2 - 2 * y - j;
                        diffCoeff = fd[im];
                    }
                } else {
                    if (q >= (1 + 2 * y + j)) {
                        int im = q - 1 - 2 * y - j;
                        diffCoeff = fd[im];
                    }
                }
                if (Y >= 0 && Y < NY) {
                    cols[k] = row + (Y - y) * st;
                    vals[k] = (fdCoeff - diffCoeff) / dh;
                    k++;
                }
            }
            return k;
        });
        scaleCsr(A, DT);
        return A;
    }

    // ------------------------------------------------------------------------------------------------------------
    // Variable grid / variable FD order: the same assembly loops with the layer's spacing and order
    // ------------------------------------------------------------------------------------------------------------
    int orderOfLayer(int layer) const { return vg.fdOrder.empty() ? d.fd_order : vg.fdOrder[layer]; }
    T dhOfLayer(int layer) const { return DH * (T)vg.dhFactor[layer]; } // Coordinates.cpp:240 varDH
    std::map<int, vector<T>> allFdCoef() const
    {
        std::map<int, vector<T>> m;
        for (int o = 2; o <= WS_MAXQ_ORACLE; o += 2)
            m[o] = fdCoef<T>(o);
        return m;
    }

    // x / z derivatives: Derivatives.cpp:129-186 (calcDxf), :706-763 (calcDxb), :304-359 (calcDzf), :1177-1243 (calcDzb).
    // On an interface plane (stored with the fine spacing, operated with the coarse one) the taps are shifted by a third
    // of the coarse spacing towards the staggered position.
    Csr<T> derivVarXZ(int axis, bool forward)
    {
        const auto fdmap = allFdCoef();
        const Idx n = axis == 0 ? NX : NZ;
        Csr<T> A = assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx x, y, z;
            coord(row, x, y, z);
            const int layer = vg.getLayer(y);
            const int f = vg.dhFactor[layer];
            int order = orderOfLayer(layer);
            Idx c = axis == 0 ? x : z;
            const Idx shift = vg.onInterface(y) ? (forward ? -(f / 3) : (f / 3)) : 0;
            int k = 0;
            for (int j = 0; j < order; j++) {
                Idx X, Xmin, Xmax;
                if (forward) {
                    X = c + shift + f * (j - order / 2 + 1);
                    Xmin = c + shift + f * (-order / 2 + 1);
                    Xmax = c + shift + f * order / 2;
                } else {
                    X = c + shift + f * (j - order / 2);
                    Xmin = c + shift + f * (-order / 2);
                    Xmax = c + shift + f * (order / 2 - 1);
                }
                if (d.edge_policy == 0) { // full order, off-grid taps dropped (what reproduces the reference's goldens, see the header of buildDerivatives)
                    if (X >= 0 && X < n) {
                        cols[k] = axis == 0 ? vg.coord2index(X, y, z) : vg.coord2index(x, y, X);
                        vals[k] = fdmap.at(order)[j] / dhOfLayer(layer);
                        k++;
                    }
                    continue;
                }
                if (Xmin < 0) {
                    order += 2 * (int)Xmin;
                    if (!forward && order == 0) {
                        order = 2;
                        c += 1;
                    }
                    j--;
                } else if (Xmax >= n) {
                    order -= 2 * (int)(Xmax - n + 1);
                    if (forward && order == 0) {
                        order = 2;
                        c -= 1;
                    }
                    j--;
                } else {
                    cols[k] = axis == 0 ? vg.coord2index(X, y, z) : vg.coord2index(x, y, X);
                    vals[k] = fdmap.at(order)[j] / dhOfLayer(layer);
                    k++;
                }
            }
            return k;
        });
        scaleCsr(A, DT);
        return A;
    }

    // order at a point of the y operators: reduced towards the interfaces (Derivatives.cpp:237-244, 803-811)
    int orderNearInterface(int layer, Idx y) const
    {
        int order = orderOfLayer(layer);
        if (!vg.variableSpacing) // useVariableFDoperators without useVariableGrid: layers of one spacing, no reduction
            return order;
        const Idx distance = vg.distToInterface(y) / vg.dhFactor[layer];
        if (distance == 0)
            order = 2;
        else if (order > distance * 2)
            order = (int)distance * 2;
        return order;
    }

    // y derivatives: Derivatives.cpp:210-280 (calcDyf), :771-843 (calcDyb), :367-440 (calcDyfFreeSurface; image = true)
    Csr<T> derivVarY(bool forward, bool image)
    {
        const auto fdmap = allFdCoef();
        Csr<T> A = assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx x, y, z;
            coord(row, x, y, z);
            const int layer = vg.getLayer(y);
            int order = orderNearInterface(layer, y);
            const bool onI = vg.onInterface(y);
            const int trans = onI ? vg.getTransition(y) : 0;
            int f = vg.dhFactor[layer];
            T dh = dhOfLayer(layer);
            if (forward && onI && trans == -1) { // coarse -> fine: the forward operator of the interface uses the fine grid below
                f = vg.dhFactor[layer + 1];
                dh = dhOfLayer(layer + 1);
            }
            int k = 0;
            if (image) {
                for (int j = 0; j < order; j++) {
                    const Idx Y = y + f * (j - order / 2 + 1);
                    const T fdCoeff = fdmap.at(order)[j];
                    T diffCoeff = 0;
                    if (order >= (2 + (int)(2 * y / f) + j)) // C++ precedence: 2 * y / dhFactor = (2 * y) / dhFactor
                        diffCoeff = fdmap.at(order)[order - 2 - (int)(2 * y / f) - j];
                    if (Y >= 0 && Y < NY) {
                        cols[k] = vg.coord2index(x, Y, z);
                        vals[k] = (fdCoeff - diffCoeff) / dh;
                        k++;
                    }
                }
                return k;
            }
            Idx yc = y;
            for (int j = 0; j < order; j++) {
                Idx Y, Ymin, Ymax;
                if (forward) {
                    Y = yc + f * (j - order / 2 + 1);
                    Ymin = yc + f * (-order / 2 + 1);
                    Ymax = yc + f * order / 2;
                } else {
                    Y = yc + f * (j - order / 2);
                    Ymin = yc + f * (-order / 2);
                    Ymax = yc + f * (order / 2 - 1);
                    // coordinate correction in the fine staggered grid (:823-830)
                    if (onI) {
                        if (j == 0 && trans == 1)
                            Y += vg.dhFactor[layer - 1];
                        if (j == 1 && trans == -1)
                            Y += vg.dhFactor[layer + 1];
                    }
                }
                if (d.edge_policy == 0) {
                    if (Y >= 0 && Y < NY) {
                        cols[k] = vg.coord2index(x, Y, z);
                        vals[k] = fdmap.at(order)[j] / (forward ? dh : dhOfLayer(layer));
                        k++;
                    }
                    continue;
                }
                if (Ymin < 0) {
                    order += 2 * (int)Ymin;
                    if (!forward && order == 0) {
                        order = 2;
                        yc += 1;
                    }
                    j--;
                } else if (Ymax >= NY) {
                    order -= 2 * (int)(Ymax - NY + 1);
                    if (forward && order == 0) {
                        order = 2;
                        yc -= 1;
                    }
                    j--;
                } else {
                    cols[k] = vg.coord2index(x, Y, z);
                    vals[k] = fdmap.at(order)[j] / (forward ? dh : dhOfLayer(layer));
                    k++;
                }
            }
            return k;
        });
        scaleCsr(A, DT);
        return A;
    }

    // bilinear interpolation of the points of an interface plane that are not points of the coarse grid
    // (Derivatives.cpp:1252-1354 full grid, :1356-1460 staggered in x, :1462-1566 staggered in z); identity elsewhere
    Csr<T> interpolationVar(int mode)
    {
        return assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx x, y, z;
            coord(row, x, y, z);
            if (!vg.onInterface(y)) {
                cols[0] = row;
                vals[0] = (T)1;
                return 1;
            }
            const int layer = vg.getLayer(y);
            const int f = vg.dhFactor[layer];
            int fFine = f;
            const int trans = vg.getTransition(y);
            if (trans == 1)
                fFine = vg.dhFactor[layer - 1];
            else if (trans == -1)
                fFine = vg.dhFactor[layer + 1];
            const T denom = (T)1 / (T)(f * f);
            const int modx = mode == 1 ? (int)((x - f / 2) % f) : (int)(x % f);
            const int modz = mode == 2 ? (int)((z - f / 2) % f) : (int)(z % f);
            const int kx = mode == 1 ? 1 : 2, kz = mode == 2 ? 1 : 2; // reach towards the upper end of the axis
            const bool lowX = mode != 1 || x >= modx, lowZ = mode != 2 || z >= modz;
            int k = 0;
            auto push = [&](Idx X, Idx Z, T v) {
                cols[k] = vg.coord2index(X, y, Z);
                vals[k] = v;
                k++;
            };
            if (lowX && lowZ)
                push(x - modx, z - modz, (T)((f - modx) * (f - modz)) * denom);
            if (x + kx * fFine < NX)
                push(x + f - modx, z - modz, (T)(modx * (f - modz)) * denom);
            if (lowX && z + kz * fFine < NZ)
                push(x - modx, z + f - modz, (T)((f - modx) * modz) * denom);
            if (x + kx * fFine < NX && z + kz * fFine < NZ)
                push(x + f - modx, z + f - modz, (T)(modx * modz) * denom);
            return k;
        });
    }

    // 2-point averaging matrices on the variable grid, Modelparameter.cpp:336-437
    Csr<T> avg2Var(int axis)
    {
        return assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx x, y, z;
            coord(row, x, y, z);
            const int layer = vg.getLayer(y);
            Idx X = x, Y = y, Z = z;
            bool inside;
            if (axis == 0) {
                X += vg.dhFactor[layer];
                inside = X < NX;
            } else if (axis == 2) {
                Z += vg.dhFactor[layer];
                inside = Z < NZ;
            } else {
                if (vg.onInterface(y) && vg.getTransition(y) == 0)
                    Y += vg.dhFactor[layer + 1];
                else
                    Y += vg.dhFactor[layer];
                inside = Y < NY;
            }
            if (!inside) {
                cols[0] = row;
                vals[0] = (T)1.0;
                return 1;
            }
            cols[0] = row;
            vals[0] = (T)(1.0 / 2.0);
            cols[1] = vg.coord2index(X, Y, Z);
            vals[1] = (T)(1.0 / 2.0);
            return 2;
        });
    }

    // Edge policy on the variable grid: like on the regular grid (tests/test_oracle_golden.py), the reference's golden
    // seismograms are reproduced by edge_policy 0 — every operator keeps its full order and the taps outside the grid are
    // dropped — within the reference's own CI gate (Test_CompareSeismogram.cpp:84), while the literal order reduction of
    // the current calc* loops (edge_policy 1; with dhFactor 3 it subtracts fine-grid distances from an order counted in
    // coarse taps, Derivatives.cpp:159-175) differs from them by 7e-3 in relative L2.
    void buildDerivatives()
    {
        if (vg.active) {
            ORACLE_REQUIRE(d.eq == WS_EQ_ACOUSTIC, "variable grid: this oracle restates the acoustic solvers only");
            Dxf = derivVarXZ(0, true);
            Dxb = derivVarXZ(0, false);
            Dyf = derivVarY(true, false);
            Dyb = derivVarY(false, false);
            if (d.dim == 3) {
                Dzf = derivVarXZ(2, true);
                Dzb = derivVarXZ(2, false);
            }
            if (d.free_surface == 1)
                DyfFS = derivVarY(true, true);
            InterFull = interpolationVar(0);
            InterStagX = interpolationVar(1);
            if (d.dim == 3)
                InterStagZ = interpolationVar(2);
            return;
        }
        auto mk = [&](int axis, bool fwd) { return d.edge_policy == 0 ? derivStencil(axis, fwd) : derivSparse(axis, fwd); };
        Dxf = mk(0, true);
        Dxb = mk(0, false);
        Dyf = mk(1, true);
        Dyb = mk(1, false);
        if (d.dim == 3) {
            Dzf = mk(2, true);
            Dzb = mk(2, false);
        }
        if (d.free_surface == 1) {
            DyfFS = derivFreeSurface(true);
            DybFS = derivFreeSurface(false);
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // Model preparation
    // ------------------------------------------------------------------------------------------------------------
    // 2-point averaging matrices, Modelparameter.cpp:336-447
    Csr<T> avg2(int axis)
    {
        if (vg.active)
            return avg2Var(axis);
        const Idx n = axisN(axis), st = axisStride(axis);
        return assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx c = axisCoord(row, axis);
            if (c + 1 < n) {
                cols[0] = row;
                vals[0] = (T)(1.0 / 2.0);
                cols[1] = row + st;
                vals[1] = (T)(1.0 / 2.0);
                return 2;
            }
            cols[0] = row;
            vals[0] = (T)1.0;
            return 1;
        });
    }
    // 4-point averaging matrices, Modelparameter.cpp:449-624 (calc4PointAverageMatrixRow + calcAverageMatrixXY/XZ/YZ)
    Csr<T> avg4(int axA, int axB)
    {
        return assemble<T>(N, [&](Idx row, Idx *cols, T *vals) {
            Idx cx, cy, cz;
            coord(row, cx, cy, cz);
            Idx c[3] = {cx, cy, cz};
            // points 1..4: (0,0) (+A,0) (0,+B) (+A,+B)
            Idx p[5][3];
            for (int k = 1; k <= 4; k++)
                for (int a = 0; a < 3; a++)
                    p[k][a] = c[a];
            p[2][axA] += 1;
            p[3][axB] += 1;
            p[4][axA] += 1;
            p[4][axB] += 1;
            Idx mx[3];
            for (int a = 0; a < 3; a++)
                mx[a] = std::max(std::max(p[1][a], p[2][a]), std::max(p[3][a], p[4][a]));
            const bool inX = mx[0] < NX, inY = mx[1] < NY, inZ = mx[2] < NZ;
            auto id = [&](Idx x, Idx y, Idx z) { return index(x, y, z); };
            int k = 0;
            if (inX && inY && inZ) {
                cols[k] = row; vals[k++] = (T)(1.0 / 4.0);
                cols[k] = id(p[2][0], p[2][1], p[2][2]); vals[k++] = (T)(1.0 / 4.0);
                cols[k] = id(p[3][0], p[3][1], p[3][2]); vals[k++] = (T)(1.0 / 4.0);
                cols[k] = id(p[4][0], p[4][1], p[4][2]); vals[k++] = (T)(1.0 / 4.0);
            }
            if (inX && !inY && inZ) { // bottom side
                cols[k] = id(mx[0], p[1][1], mx[2]); vals[k++] = (T)(1.0 / 2.0);
                cols[k] = row; vals[k++] = (T)(1.0 / 2.0);
            }
            if (!inX && inY && inZ) { // right side
                cols[k] = id(p[1][0], mx[1], mx[2]); vals[k++] = (T)(1.0 / 2.0);
                cols[k] = row; vals[k++] = (T)(1.0 / 2.0);
            }
            if (inX && inY && !inZ) { // back side
                cols[k] = id(mx[0], mx[1], p[1][2]); vals[k++] = (T)(1.0 / 2.0);
                cols[k] = row; vals[k++] = (T)(1.0 / 2.0);
            }
            if (!inX && !inY && inZ) { cols[k] = row; vals[k++] = (T)1.0; }
            if (inX && !inY && !inZ) { cols[k] = row; vals[k++] = (T)1.0; }
            if (!inX && inY && !inZ) { cols[k] = row; vals[k++] = (T)1.0; }
            // a diagonal entry pushed twice cannot happen; duplicates of (row,row) in the half cases are distinct columns
            return k;
        });
    }

    // Modelparameter.cpp:633-639
    void calcInverseAveragedParameter(const vector<T> &par, vector<T> &out, const Csr<T> &A)
    {
        spmv(A, par, out);
        T *p = out.data();
        VFOR(out.size()) p[i] = (T)1 / p[i];
        replaceInvalid(out, (T)0.0);
    }
    // ModelparameterSeismic.cpp:421-432 (NB: clamps the input vector in place)
    void calcAveragedSWaveModulus(vector<T> &mu, vector<T> &out, const Csr<T> &A)
    {
        searchAndReplaceLess(mu, (T)1.0, (T)1.0);
        vector<T> inv(mu.size());
        {
            T *p = inv.data();
            const T *q = mu.data();
            VFOR(mu.size()) p[i] = (T)1 / q[i];
        }
        vector<T> tmp;
        spmv(A, inv, tmp);
        out.resize(mu.size());
        {
            T *p = out.data();
            const T *q = tmp.data();
            VFOR(mu.size()) p[i] = (T)1 / q[i];
        }
        searchAndReplaceLess(out, (T)4.0, (T)0.0);
    }
    // ModelparameterSeismic.cpp:131-136
    void calcModulusFromVelocity(const vector<T> &v, const vector<T> &rho, vector<T> &modulus)
    {
        modulus = rho;
        vmul(modulus, v);
        vmul(modulus, v);
    }
    bool has(const std::string &k) const { return mat.count(k) > 0; }

    // Viscoelastic.cpp:650-713: relaxed modulus scaling
    void viscoScaleModulus(vector<T> &modulus, const vector<T> &tau)
    {
        T w_ref = (T)(2.0 * M_PI * d.fc_cpml);
        T sum = 0;
        for (int l = 0; l < L; l++) {
            T tauSigma = (T)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
            sum += (T)(w_ref * w_ref * tauSigma * tauSigma / (1.0 + w_ref * w_ref * tauSigma * tauSigma));
        }
        T *p = modulus.data();
        const T *q = tau.data();
        VFOR(modulus.size())
        {
            T temp = (T)1.0 + sum * q[i];
            p[i] = p[i] / temp;
        }
    }

    void prepareModelSeismic()
    {
        const bool needP = (d.eq == WS_EQ_ACOUSTIC || d.eq == WS_EQ_ELASTIC || d.eq == WS_EQ_VISCOELASTIC);
        const bool needS = (d.eq != WS_EQ_ACOUSTIC);
        const bool vis = visco();
        if (needP && !has("pWaveModulus")) {
            calcModulusFromVelocity(M("velocityP"), M("density"), mat["pWaveModulus"]);
            if (vis)
                viscoScaleModulus(mat["pWaveModulus"], M("tauP"));
        }
        if (needS && !has("sWaveModulus")) {
            calcModulusFromVelocity(M("velocityS"), M("density"), mat["sWaveModulus"]);
            if (vis)
                viscoScaleModulus(mat["sWaveModulus"], M("tauS"));
        }
        if (d.eq == WS_EQ_SH || d.eq == WS_EQ_VISCOSH) {
            // SH.cpp:424-430: 2-point x / y averaging; un-averaged inverse density (ModelparameterSeismic.cpp:164-172)
            if (!has("inverseDensity")) {
                const vector<T> &rho = M("density");
                vector<T> &o = mat["inverseDensity"];
                o.resize(N);
                VFOR(N) o[i] = (T)1 / rho[i];
            }
            Csr<T> AX = avg2(0), AY = avg2(1);
            if (!has("sWaveModulusAverageXZ"))
                calcAveragedSWaveModulus(M("sWaveModulus"), mat["sWaveModulusAverageXZ"], AX);
            if (!has("sWaveModulusAverageYZ"))
                calcAveragedSWaveModulus(M("sWaveModulus"), mat["sWaveModulusAverageYZ"], AY);
            if (vis) { // ViscoSH: tauS averages with the same matrices
                if (!has("tauSAverageXZ"))
                    spmv(AX, M("tauS"), mat["tauSAverageXZ"]);
                if (!has("tauSAverageYZ"))
                    spmv(AY, M("tauS"), mat["tauSAverageYZ"]);
            }
            return;
        }
        // Acoustic.cpp:381-387, Elastic.cpp:576-588, Viscoelastic.cpp:562-577
        if (!has("inverseDensityAverageX")) {
            Csr<T> A = avg2(0);
            calcInverseAveragedParameter(M("density"), mat["inverseDensityAverageX"], A);
        }
        if (!has("inverseDensityAverageY")) {
            Csr<T> A = avg2(1);
            calcInverseAveragedParameter(M("density"), mat["inverseDensityAverageY"], A);
        }
        if (d.dim == 3 && !has("inverseDensityAverageZ")) {
            Csr<T> A = avg2(2);
            calcInverseAveragedParameter(M("density"), mat["inverseDensityAverageZ"], A);
        }
        if (d.eq == WS_EQ_ACOUSTIC)
            return;
        {
            Csr<T> A = avg4(0, 1);
            if (!has("sWaveModulusAverageXY"))
                calcAveragedSWaveModulus(M("sWaveModulus"), mat["sWaveModulusAverageXY"], A);
            if (vis && !has("tauSAverageXY"))
                spmv(A, M("tauS"), mat["tauSAverageXY"]);
        }
        if (d.dim == 3) {
            Csr<T> A = avg4(0, 2);
            if (!has("sWaveModulusAverageXZ"))
                calcAveragedSWaveModulus(M("sWaveModulus"), mat["sWaveModulusAverageXZ"], A);
            if (vis && !has("tauSAverageXZ"))
                spmv(A, M("tauS"), mat["tauSAverageXZ"]);
            Csr<T> B = avg4(1, 2);
            if (!has("sWaveModulusAverageYZ"))
                calcAveragedSWaveModulus(M("sWaveModulus"), mat["sWaveModulusAverageYZ"], B);
            if (vis && !has("tauSAverageYZ"))
                spmv(B, M("tauS"), mat["tauSAverageYZ"]);
        }
    }

    // ---- EM model preparation + coefficient builders ------------------------------------------------------------
    // ForwardSolverEM.cpp:14-33
    vector<T> getAveragedCinv(const vector<T> &eps, const vector<T> &sig)
    {
        vector<T> c(N);
        VFOR(N)
        {
            T v = (T)0.5 / eps[i];
            v = v * sig[i];
            v = v * DT;
            v = v + (T)1;
            c[i] = (T)1 / v;
        }
        return c;
    }
    // ForwardSolverEM.cpp:35-57
    vector<T> getAveragedCa(const vector<T> &eps, const vector<T> &sig)
    {
        vector<T> cinv = getAveragedCinv(eps, sig);
        vector<T> c(N);
        VFOR(N)
        {
            T v = (T)0.5 / eps[i];
            v = v * sig[i];
            v = v * DT;
            v = (T)1 - v;
            c[i] = v * cinv[i];
        }
        return c;
    }
    // ForwardSolverEM.cpp:59-76
    vector<T> getAveragedCb(const vector<T> &eps, const vector<T> &sig)
    {
        vector<T> cinv = getAveragedCinv(eps, sig);
        vector<T> c(N);
        VFOR(N)
        {
            T v = (T)1 / eps[i];
            c[i] = v * cinv[i];
        }
        return c;
    }
    // ForwardSolverEM.cpp:78-93
    vector<T> getCc()
    {
        vector<T> r;
        for (int l = 0; l < L; l++)
            r.push_back((T)((1 - 0.5 * DT / relaxationTime[l]) / (1 + 0.5 * DT / relaxationTime[l])));
        return r;
    }
    // ForwardSolverEM.cpp:95-117
    vector<T> getAveragedCd(const vector<T> &epsStatic, const vector<T> &tauEps, int l)
    {
        T tempValue = (T)(1 / (1 + 0.5 * DT / relaxationTime[l]));
        tempValue /= (L * relaxationTime[l] * relaxationTime[l]);
        vector<T> c(N);
        VFOR(N)
        {
            T v = tauEps[i] * tempValue;
            v = v * epsStatic[i];
            c[i] = v * (-DT);
        }
        return c;
    }
    // ForwardSolverEM.cpp:122-135
    vector<T> sigmaEffectiveOptical(const vector<T> &eps, const vector<T> &sig, const vector<T> &tauEps)
    {
        T sum = 0;
        for (int l = 0; l < L; l++)
            sum += (T)(1.0 / relaxationTime[l]);
        sum /= L;
        vector<T> c(N);
        VFOR(N)
        {
            T v = tauEps[i] * sum;
            v = v * eps[i];
            c[i] = v + sig[i];
        }
        return c;
    }
    // ForwardSolverEM.cpp:140-154; eps0 from Modelparameter.hpp:359-360
    vector<T> epsEffectiveOptical(const vector<T> &eps, const vector<T> &sig, const vector<T> &tauEps, const vector<T> &tauSig)
    {
        const T eps0 = (T)8.8541878176e-12;
        vector<T> c(N);
        VFOR(N)
        {
            T v = (T)1 - tauEps[i];
            v = v * eps[i];
            T t = sig[i] * tauSig[i];
            c[i] = v + t;
        }
        searchAndReplaceLess(c, eps0, eps0);
        return c;
    }

    // plain average (ModelparameterEM analogue of calcAveragedParameter) and inverse average
    void avgPlain(const vector<T> &in, vector<T> &out, const Csr<T> &A) { spmv(A, in, out); }

    void prepareModelEM()
    {
        // Inputs are absolute SI values: dielectricPermittivity (eps), electricConductivity (sigma),
        // magneticPermeability (mu), and for visco: tauDielectricPermittivity, tauElectricConductivity (absolute, s).
        // TMEM.cpp / ViscoTMEM.cpp:442-450: mu^-1 2-point averaged in x (-> XZ) and y (-> YZ); eps/sigma cell centred.
        // EMEM.cpp:378-392 / ViscoEMEM.cpp:409-423: mu^-1 4-point on YZ/XZ/XY; sigma, eps (tau*) 2-point in x/y/z.
        const bool vis = visco();
        relaxationTime.clear();
        for (int l = 0; l < L; l++)
            relaxationTime.push_back((T)(1.0 / (2.0 * M_PI * d.relax_freq[l])));
        if (vis)
            Cc = getCc();
        if (d.eq == WS_EQ_TMEM || d.eq == WS_EQ_VISCOTMEM) {
            Csr<T> AX = avg2(0), AY = avg2(1);
            if (!has("inverseMagneticPermeabilityAverageXZ"))
                calcInverseAveragedParameter(M("magneticPermeability"), mat["inverseMagneticPermeabilityAverageXZ"], AX);
            if (!has("inverseMagneticPermeabilityAverageYZ"))
                calcInverseAveragedParameter(M("magneticPermeability"), mat["inverseMagneticPermeabilityAverageYZ"], AY);
            const vector<T> &eps = M("dielectricPermittivity");
            const vector<T> &sig = M("electricConductivity");
            if (!vis) { // ForwardSolver2Dtmem.cpp:70-78
                mat["CaAverageZ"] = getAveragedCa(eps, sig);
                mat["CbAverageZ"] = getAveragedCb(eps, sig);
            } else { // ForwardSolver2Dviscotmem.cpp:79-104
                const vector<T> &te = M("tauDielectricPermittivity");
                const vector<T> &ts = M("tauElectricConductivity");
                vector<T> epsO = epsEffectiveOptical(eps, sig, te, ts);
                vector<T> sigO = sigmaEffectiveOptical(eps, sig, te);
                mat["CaAverageZ"] = getAveragedCa(epsO, sigO);
                mat["CbAverageZ"] = getAveragedCb(epsO, sigO);
                for (int l = 0; l < L; l++)
                    mat["CdAverageZ" + std::to_string(l + 1)] = getAveragedCd(eps, te, l);
            }
            return;
        }
        // EMEM / ViscoEMEM (2D TE and 3D)
        static const char *AXN[3] = {"X", "Y", "Z"};
        {
            Csr<T> A = avg4(0, 1);
            if (!has("inverseMagneticPermeabilityAverageXY"))
                calcInverseAveragedParameter(M("magneticPermeability"), mat["inverseMagneticPermeabilityAverageXY"], A);
        }
        if (d.dim == 3) {
            Csr<T> A = avg4(0, 2);
            if (!has("inverseMagneticPermeabilityAverageXZ"))
                calcInverseAveragedParameter(M("magneticPermeability"), mat["inverseMagneticPermeabilityAverageXZ"], A);
            Csr<T> B = avg4(1, 2);
            if (!has("inverseMagneticPermeabilityAverageYZ"))
                calcInverseAveragedParameter(M("magneticPermeability"), mat["inverseMagneticPermeabilityAverageYZ"], B);
        }
        const int nax = d.dim == 3 ? 3 : 2;
        for (int a = 0; a < nax; a++) {
            Csr<T> A = avg2(a);
            vector<T> eps, sig;
            avgPlain(M("dielectricPermittivity"), eps, A);
            avgPlain(M("electricConductivity"), sig, A);
            std::string ax = AXN[a];
            if (!vis) { // ForwardSolver2Demem.cpp / 3Demem prepareForModelling
                mat["CaAverage" + ax] = getAveragedCa(eps, sig);
                mat["CbAverage" + ax] = getAveragedCb(eps, sig);
            } else {
                vector<T> te, ts;
                avgPlain(M("tauDielectricPermittivity"), te, A);
                avgPlain(M("tauElectricConductivity"), ts, A);
                vector<T> epsO = epsEffectiveOptical(eps, sig, te, ts);
                vector<T> sigO = sigmaEffectiveOptical(eps, sig, te);
                mat["CaAverage" + ax] = getAveragedCa(epsO, sigO);
                mat["CbAverage" + ax] = getAveragedCb(epsO, sigO);
                for (int l = 0; l < L; l++)
                    mat["CdAverage" + ax + std::to_string(l + 1)] = getAveragedCd(eps, te, l);
            }
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // Boundary conditions
    // ------------------------------------------------------------------------------------------------------------
    // Coordinates.cpp:734-750 edgeDistance
    static Idx edgeDist(Idx c, Idx n) { return !((n - 1 - c) < c) ? c : (n - 1 - c); }

    // ABS3D.cpp:154-218, ABS2D.cpp:115-178
    void initABS()
    {
        // variable grid: the same rule on the coordinates of the points (index2coordinate + edgeDistance work in units of the finest spacing, so a
        // coarse layer picks every dhFactor-th coefficient)
        const int W = d.boundary_width;
        damping.assign(N, (T)1.0);
        vector<T> coeff(W);
        T amp = (T)(1.0 - d.damping_coeff / 100.0);
        T a = (T)std::sqrt(-std::log(amp) / (T)(W * W));
        for (int j = 0; j < W; j++)
            coeff[j] = (T)std::exp(-(a * a * (W - j) * (W - j)));
        const bool fs = d.free_surface != 0; // ABS*::init takes useFreeSurface (0,1,2) and tests == 0
        for (Idx i = 0; i < N; i++) {
            Idx x, y, z;
            coord(i, x, y, z);
            Idx dx = edgeDist(x, NX), dy = edgeDist(y, NY), dz = edgeDist(z, NZ);
            if (d.dim == 3) {
                Idx mn = dx < dy ? dx : dy;
                if (dz < mn)
                    mn = dz;
                if (!fs) {
                    if (mn < W)
                        damping[i] = coeff[mn];
                } else {
                    Idx xz = !(dx < dz) ? dz : dx;
                    if (y < W) {
                        if (dz < W || dx < W)
                            damping[i] = coeff[xz];
                    } else if (mn < W)
                        damping[i] = coeff[mn];
                }
            } else {
                Idx mn = dx < dy ? dx : dy;
                if (!fs) {
                    if (mn < W)
                        damping[i] = coeff[mn];
                } else {
                    if (y < W) {
                        if (dx < W)
                            damping[i] = coeff[dx];
                    } else if (mn < W)
                        damping[i] = coeff[mn];
                }
            }
        }
    }

    // CPML.cpp:39-68
    void calcCoeffCPML(vector<T> &a, vector<T> &b, bool shiftGrid) { calcCoeffCPML(a, b, shiftGrid, DH); }
    void calcCoeffCPML(vector<T> &a, vector<T> &b, bool shiftGrid, T DH)
    {
        const int W = (int)a.size();
        T shift = shiftGrid ? (T)0.5 : (T)0;
        T RCoef = (T)0.0008;
        T alpha_max = (T)(2.0 * M_PI * (d.fc_cpml / 2.0));
        T NPower = d.npower, VMax = d.vmax_cpml;
        T d0 = (T)(-(NPower + 1) * VMax * std::log(RCoef) / (2.0 * W * DH));
        for (int i = 0; i < W; i++) {
            T pos = (T)(W - i - shift) / W;
            T dd = d0 * (T)std::pow(pos, NPower);
            T alpha_prime = (T)(alpha_max * (1.0 - pos));
            b[i] = (T)std::exp(-(dd + alpha_prime) * DT);
            if (std::abs(dd) > 1.0e-6)
                a[i] = (T)(dd * (b[i] - 1.0) / (dd + alpha_prime));
            else
                a[i] = 0;
        }
    }

    // CPML3D.cpp:222-368, CPML2D.cpp:170-285 (same rules; 2D has no z)
    // CPML2DAcoustic.cpp:99-200, CPML3DAcoustic.cpp init: per layer a profile of ceil(BoundaryWidth / dhFactor) points with the layer's DH
    void initCPMLVar()
    {
        const int fs = d.free_surface;
        vector<vector<T>> a, b, ah, bh;
        for (int l = 0; l < vg.numLayers; l++) {
            const int width = (int)std::ceil((float)d.boundary_width / vg.dhFactor[l]);
            vector<T> al(width), bl(width), ahl(width), bhl(width);
            calcCoeffCPML(al, bl, false, dhOfLayer(l));
            calcCoeffCPML(ahl, bhl, true, dhOfLayer(l));
            a.push_back(al); b.push_back(bl); ah.push_back(ahl); bh.push_back(bhl);
        }
        px = Profile<T>();
        py = Profile<T>();
        pz = Profile<T>();
        for (Idx i = 0; i < N; i++) {
            Idx x, y, z;
            coord(i, x, y, z);
            const int l = vg.getLayer(y), f = vg.dhFactor[l];
            const int width = (int)std::ceil((float)d.boundary_width / f);
            const Idx xDist = edgeDist(x, NX) / f, yDist = edgeDist(y, NY) / f, zDist = edgeDist(z, NZ) / f;
            auto push = [&](Profile<T> &p, Idx dist, bool low) {
                p.idx.push_back(i);
                if (low) {
                    p.a.push_back(a[l][dist]); p.b.push_back(b[l][dist]); p.ah.push_back(ah[l][dist]); p.bh.push_back(bh[l][dist]);
                } else {
                    p.a.push_back(ah[l][dist]); p.b.push_back(bh[l][dist]); p.ah.push_back(a[l][dist]); p.bh.push_back(b[l][dist]);
                }
            };
            if (xDist < width)
                push(px, xDist, x / f < width);
            if (yDist < width) {
                if (y / f < width) {
                    if (fs == 0)
                        push(py, yDist, true);
                } else
                    push(py, yDist, false);
            }
            if (d.dim == 3 && zDist < width)
                push(pz, zDist, z / f < width);
        }
    }
    void initCPML()
    {
        if (vg.active) {
            initCPMLVar();
            return;
        }
        const int W = d.boundary_width;
        vector<T> a(W), b(W), ah(W), bh(W);
        calcCoeffCPML(a, b, false);
        calcCoeffCPML(ah, bh, true);
        const int fs = d.free_surface; // useFreeSurface; only "== 0" enables the top layer
        px = Profile<T>();
        py = Profile<T>();
        pz = Profile<T>();
        for (Idx i = 0; i < N; i++) {
            Idx x, y, z;
            coord(i, x, y, z);
            Idx dx = edgeDist(x, NX), dy = edgeDist(y, NY), dz = edgeDist(z, NZ);
            auto push = [&](Profile<T> &p, Idx dist, bool low) {
                p.idx.push_back(i);
                if (low) {
                    p.a.push_back(a[dist]); p.b.push_back(b[dist]); p.ah.push_back(ah[dist]); p.bh.push_back(bh[dist]);
                } else { // swapped on the high-coordinate side, CPML3D.cpp:304-317
                    p.a.push_back(ah[dist]); p.b.push_back(bh[dist]); p.ah.push_back(a[dist]); p.bh.push_back(b[dist]);
                }
            };
            if (dx < W)
                push(px, dx, x < W);
            if (dy < W) {
                if (y < W) {
                    if (fs == 0)
                        push(py, dy, true);
                } else
                    push(py, dy, false);
            }
            if (d.dim == 3 && dz < W)
                push(pz, dz, z < W);
        }
    }
    vector<T> &psiOf(const std::string &name, const Profile<T> &p)
    {
        vector<T> &v = psi[name];
        if (v.size() != p.idx.size())
            v.assign(p.idx.size(), (T)0);
        return v;
    }
    // CPML.cpp:84-95 applyCPML: temp = a; Psi *= b; temp *= Vec; Psi += temp; Vec += Psi
    void applyCPML(vector<T> &vec, const std::string &psiName, const Profile<T> &p, bool half)
    {
        if (d.damping != 2)
            return;
        vector<T> &ps = psiOf(psiName, p);
        const vector<T> &A = half ? p.ah : p.a;
        const vector<T> &B = half ? p.bh : p.b;
        const Idx n = (Idx)p.idx.size();
        T *v = vec.data();
#pragma omp parallel for schedule(static)
        for (Idx k = 0; k < n; k++) {
            Idx i = p.idx[k];
            T temp = A[k];
            ps[k] = ps[k] * B[k];
            temp = temp * v[i];
            ps[k] = ps[k] + temp;
            v[i] = v[i] + ps[k];
        }
    }

    void initFreeSurface()
    {
        surfIdx.clear();
        for (Idx i = 0; i < N; i++) {
            Idx x, y, z;
            coord(i, x, y, z);
            if (y == 0)
                surfIdx.push_back(i); // Coordinates.cpp:598-607
        }
    }
    // FreeSurface.cpp:13-20
    void setSurfaceZero(vector<T> &v)
    {
        for (Idx i : surfIdx)
            v[i] = v[i] * (T)0;
    }
    // FreeSurfaceElastic.cpp:11-47
    void fsSetModelElastic()
    {
        const vector<T> &pw = M("pWaveModulus");
        const vector<T> &sw = M("sWaveModulus");
        sH.resize(surfIdx.size());
        sV.resize(surfIdx.size());
        for (size_t k = 0; k < surfIdx.size(); k++) {
            Idx i = surfIdx[k];
            ORACLE_REQUIRE(sw[i] > 0, "S wave modulus can't be zero when using image method");
            T temp = pw[i] - (T)2 * sw[i];
            sV[k] = (T)1 * temp;
            temp = temp * temp;
            temp = temp / pw[i];
            temp = temp * (T)-1;
            sH[k] = (T)1 * temp;
        }
    }
    // FreeSurfaceViscoelastic.cpp:12-96
    vector<T> sSH, sSV; // scaleStressHorizontalUpdate / scaleStressVerticalUpdate
    void fsSetModelVisco()
    {
        const vector<T> &pw = M("pWaveModulus");
        const vector<T> &sw = M("sWaveModulus");
        const vector<T> &tauS = M("tauS");
        const vector<T> &tauP = M("tauP");
        const size_t ns = surfIdx.size();
        sSH.resize(ns);
        sSV.resize(ns);
        sRH.assign(L, vector<T>(ns));
        sRV.assign(L, vector<T>(ns));
        for (size_t k = 0; k < ns; k++) {
            Idx i = surfIdx[k];
            ORACLE_REQUIRE(sw[i] > 0, "S wave modulus can't be zero when using image method");
            T temp = ((T)-2 * sw[i]) * onePlusLtauS[i];
            T temp2 = pw[i] * onePlusLtauP[i];
            temp = temp + temp2;
            temp2 = (T)1 / temp2;
            sSV[k] = (T)1 * temp;
            T h = (T)-1 * (T)1;
            h = h * temp;
            h = h * temp;
            h = h * temp2;
            sSH[k] = h;
            for (int l = 0; l < L; l++) {
                T relTime = (T)(1.0 / (2.0 * M_PI * d.relax_freq[l]));
                T vc2 = (T)(1.0 / (1.0 + DT / (2.0 * relTime)));
                T t = (T)2 * sw[i];
                T t2 = pw[i];
                T t3 = t * tauS[i];
                t3 = t3 - t2 * tauP[i];
                T rv = (T)1 * t3;
                rv = rv * vc2;
                rv = rv / relTime;
                sRV[l][k] = rv;
                t = t * onePlusLtauS[i];
                t2 = t2 * onePlusLtauP[i];
                t = t / t2;
                t = t - (T)1;
                T rh = (T)1 * t3;
                rh = rh * t;
                rh = rh * vc2;
                rh = rh / relTime;
                sRH[l][k] = rh;
            }
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // setup
    // ------------------------------------------------------------------------------------------------------------
    std::string Rn(const char *base, int l) const { return std::string(base) + std::to_string(l + 1); }

    void create(const ws_desc &desc)
    {
        d = desc;
        ORACLE_REQUIRE(d.dim == 2 || d.dim == 3, "dimension must be 2 or 3");
        ORACLE_REQUIRE(d.nranks <= 1, "the oracle is single-domain");
        NX = d.nx;
        NY = d.ny;
        NZ = d.dim == 2 ? 1 : d.nz;
        ORACLE_REQUIRE(NX > 0 && NY > 0 && NZ > 0, "invalid grid");
        ORACLE_REQUIRE((int64_t)NX * NY * NZ < (int64_t)1 << 31, "grid too large for int32 indices");
        N = NX * NY * NZ;
        if (vg.active) { // the grid may have been trimmed to fit the coarsest spacing (Coordinates.cpp:153-200)
            NX = vg.NX; NY = vg.NY; NZ = vg.NZ;
            N = vg.n;
        }
        DT = d.dt;
        DH = d.dh;
        L = visco() ? d.n_relax : 0;
        ORACLE_REQUIRE(L <= WS_MAX_RELAX, "numRelaxationMechanisms more than 4 is not available here!");
        ORACLE_REQUIRE(!visco() || L >= 1, "visco solvers need numRelaxationMechanisms >= 1");
        auto mkf = [&](const std::string &n) { fld[n].assign(N, (T)0); };
        switch (d.eq) {
        case WS_EQ_ACOUSTIC:
            mkf("VX"); mkf("VY"); if (d.dim == 3) mkf("VZ"); mkf("P");
            break;
        case WS_EQ_ELASTIC:
        case WS_EQ_VISCOELASTIC:
            mkf("VX"); mkf("VY"); mkf("Sxx"); mkf("Syy"); mkf("Sxy");
            if (d.dim == 3) { mkf("VZ"); mkf("Szz"); mkf("Sxz"); mkf("Syz"); }
            for (int l = 0; l < L; l++) {
                mkf(Rn("Rxx", l)); mkf(Rn("Ryy", l)); mkf(Rn("Rxy", l));
                if (d.dim == 3) { mkf(Rn("Rzz", l)); mkf(Rn("Rxz", l)); mkf(Rn("Ryz", l)); }
            }
            break;
        case WS_EQ_SH:
        case WS_EQ_VISCOSH:
            ORACLE_REQUIRE(d.dim == 2, "sh is 2D only");
            mkf("VZ"); mkf("Sxz"); mkf("Syz");
            for (int l = 0; l < L; l++) { mkf(Rn("Rxz", l)); mkf(Rn("Ryz", l)); }
            break;
        case WS_EQ_TMEM:
        case WS_EQ_VISCOTMEM:
            ORACLE_REQUIRE(d.dim == 2, "tmem is 2D only");
            mkf("HX"); mkf("HY"); mkf("EZ");
            for (int l = 0; l < L; l++) mkf(Rn("RZ", l));
            break;
        case WS_EQ_EMEM:
        case WS_EQ_VISCOEMEM:
            mkf("HZ"); mkf("EX"); mkf("EY");
            if (d.dim == 3) { mkf("HX"); mkf("HY"); mkf("EZ"); }
            for (int l = 0; l < L; l++) {
                mkf(Rn("RX", l)); mkf(Rn("RY", l));
                if (d.dim == 3) mkf(Rn("RZ", l));
            }
            break;
        default:
            throw std::runtime_error("unknown equationType");
        }
    }

    void prepare()
    {
        buildDerivatives();
        if (seismic())
            prepareModelSeismic();
        else
            prepareModelEM();
        if (visco() && seismic()) {
            // ForwardSolver3Dviscoelastic.cpp:55-66, :103-118
            relaxationTime.clear(); inverseRelaxationTime.clear(); viscoCoeff1.clear(); viscoCoeff2.clear();
            for (int l = 0; l < L; l++) {
                relaxationTime.push_back((T)(1.0 / (2.0 * M_PI * d.relax_freq[l])));
                inverseRelaxationTime.push_back((T)(1.0 / relaxationTime[l]));
                viscoCoeff1.push_back((T)(1.0 - DT / (2.0 * relaxationTime[l])));
                viscoCoeff2.push_back((T)(1.0 / (1.0 + DT / (2.0 * relaxationTime[l]))));
            }
            DThalf = (T)(DT / 2.0);
            const vector<T> &tauS = M("tauS");
            onePlusLtauS.resize(N);
            VFOR(N) onePlusLtauS[i] = (T)1.0 + (T)L * tauS[i];
            if (d.eq == WS_EQ_VISCOELASTIC) {
                const vector<T> &tauP = M("tauP");
                onePlusLtauP.resize(N);
                VFOR(N) onePlusLtauP[i] = (T)1.0 + (T)L * tauP[i];
            }
        }
        const bool emSolver = !seismic();
        if (d.free_surface == 1 && !emSolver) {
            initFreeSurface();
            if (d.eq == WS_EQ_ELASTIC)
                fsSetModelElastic();
            if (d.eq == WS_EQ_VISCOELASTIC)
                fsSetModelVisco();
        }
        if (d.damping == 1)
            initABS();
        if (d.damping == 2)
            initCPML();
        psi.clear();
        update.assign(N, 0); update_temp.assign(N, 0); update2.assign(N, 0);
        vxx.assign(N, 0); vyy.assign(N, 0); vzz.assign(N, 0);
        prepared = true;
    }

    void reset()
    {
        for (auto &kv : fld)
            std::fill(kv.second.begin(), kv.second.end(), (T)0);
        for (auto &kv : psi)
            std::fill(kv.second.begin(), kv.second.end(), (T)0);
        std::fill(seis.begin(), seis.end(), (T)0);
    }

    // ------------------------------------------------------------------------------------------------------------
    // sources / receivers (SourceReceiverImpl.cpp:12-37, FDTD3Delastic.cpp:12-53, FDTD2Delastic.cpp, FDTDacoustic.cpp,
    // ForwardSolverEM/SourceReceiverImpl/SourceReceiverImplEM.cpp)
    // ------------------------------------------------------------------------------------------------------------
    void addAt(const char *f, Idx i, T v) { vector<T> &w = F(f); w[i] = w[i] + v; }
    void applySource(Idx t)
    {
        const Idx ns = (Idx)srcIdx.size();
        for (int type = 1; type <= 4; type++) // reference order: P, VX, VY, VZ (EZ, EX, EY, HZ)
            for (Idx s = 0; s < ns; s++) {
                if (srcType[s] != type)
                    continue;
                T v = srcSig[(size_t)s * d.nt + t];
                Idx i = srcIdx[s];
                if (seismic()) {
                    switch (type) {
                    case WS_TYPE_P:
                        if (d.eq == WS_EQ_ACOUSTIC)
                            addAt("P", i, v);
                        else if (d.eq == WS_EQ_ELASTIC || d.eq == WS_EQ_VISCOELASTIC) {
                            addAt("Sxx", i, v); addAt("Syy", i, v);
                            if (d.dim == 3) addAt("Szz", i, v);
                        } else
                            throw std::runtime_error("Pressure sources can not be implemented in SH modeling");
                        break;
                    case WS_TYPE_VX:
                        ORACLE_REQUIRE(d.eq != WS_EQ_SH && d.eq != WS_EQ_VISCOSH, "VX sources can not be implemented in SH modeling");
                        addAt("VX", i, v);
                        break;
                    case WS_TYPE_VY:
                        ORACLE_REQUIRE(d.eq != WS_EQ_SH && d.eq != WS_EQ_VISCOSH, "VY sources can not be implemented in SH modeling");
                        addAt("VY", i, v);
                        break;
                    case WS_TYPE_VZ:
                        ORACLE_REQUIRE(fld.count("VZ"), "no VZ wavefield in this modelling");
                        addAt("VZ", i, v);
                        break;
                    }
                } else {
                    static const char *nm[5] = {"", "EZ", "EX", "EY", "HZ"};
                    ORACLE_REQUIRE(fld.count(nm[type]), std::string("no ") + nm[type] + " wavefield in this modelling");
                    addAt(nm[type], i, v);
                }
            }
    }
    void gatherSeismogram(Idx t)
    {
        const Idx nr = (Idx)recIdx.size();
        for (Idx r = 0; r < nr; r++) {
            Idx i = recIdx[r];
            T v = 0;
            const int type = recType[r];
            if (seismic()) {
                switch (type) {
                case WS_TYPE_P:
                    if (d.eq == WS_EQ_ACOUSTIC)
                        v = F("P")[i] * (T)1;
                    else if (d.eq == WS_EQ_ELASTIC || d.eq == WS_EQ_VISCOELASTIC) {
                        if (d.dim == 3) {
                            v = F("Sxx")[i];
                            v = v + F("Syy")[i];
                            v = v + F("Szz")[i];
                            v = v / (T)3;
                        } else {
                            v = F("Sxx")[i];
                            v = v + F("Syy")[i];
                            v = v * (T)0.5;
                        }
                    } else
                        throw std::runtime_error("Pressure receivers can not be implemented in SH modeling");
                    break;
                case WS_TYPE_VX: v = F("VX")[i]; break;
                case WS_TYPE_VY: v = F("VY")[i]; break;
                case WS_TYPE_VZ: v = F("VZ")[i]; break;
                default: throw std::runtime_error("unknown receiver type");
                }
            } else {
                static const char *nm[5] = {"", "EZ", "EX", "EY", "HZ"};
                ORACLE_REQUIRE(type >= 1 && type <= 4, "unknown receiver type");
                v = F(nm[type])[i];
            }
            seis[(size_t)r * d.nt + t] = v;
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // time steps — one statement per reference statement
    // ------------------------------------------------------------------------------------------------------------
    void absApply(std::initializer_list<const char *> names)
    {
        if (d.damping != 1)
            return;
        for (const char *n : names)
            vmul(F(n), damping);
    }

    // ForwardSolver3Dacoustic.cpp:131-229, ForwardSolver2Dacoustic.cpp:121-193
    void stepAcoustic(Idx t)
    {
        const bool fs = d.free_surface == 1, d3 = d.dim == 3;
        vector<T> &p = F("P"), &vX = F("VX"), &vY = F("VY");
        // variable grid: the points of the interface planes that are not coarse-grid points are interpolated after every
        // update (ForwardSolver2Dacoustic.cpp:127-157,179-183; ForwardSolver3Dacoustic.cpp:141-190,218-222)
        auto interpolate = [&](const Csr<T> &I, vector<T> &v) {
            if (!vg.active)
                return;
            update_temp.swap(v);
            spmv(I, update_temp, v);
        };
        spmv(Dxf, p, update);
        applyCPML(update, "p_x", px, true);
        vmul(update, M("inverseDensityAverageX"));
        vadd(vX, update);
        interpolate(InterStagX, vX);
        spmv(fs ? DyfFS : Dyf, p, update);
        applyCPML(update, "p_y", py, true);
        vmul(update, M("inverseDensityAverageY"));
        vadd(vY, update);
        interpolate(InterFull, vY);
        if (d3) {
            spmv(Dzf, p, update);
            applyCPML(update, "p_z", pz, true);
            vmul(update, M("inverseDensityAverageZ"));
            vadd(F("VZ"), update);
            interpolate(InterStagZ, F("VZ"));
        }
        spmv(Dxb, vX, update);
        applyCPML(update, "vxx", px, false);
        spmv(Dyb, vY, update_temp);
        applyCPML(update_temp, "vyy", py, false);
        vadd(update, update_temp);
        if (d3) {
            spmv(Dzb, F("VZ"), update_temp);
            applyCPML(update_temp, "vzz", pz, false);
            vadd(update, update_temp);
        }
        vmul(update, M("pWaveModulus"));
        vadd(p, update);
        if (d3)
            absApply({"P", "VX", "VY", "VZ"});
        else
            absApply({"P", "VX", "VY"});
        interpolate(InterFull, p);
        if (fs)
            setSurfaceZero(p);
        applySource(t);
        gatherSeismogram(t);
    }

    // velocity half-step shared by elastic and viscoelastic
    // ForwardSolver3Delastic.cpp:181-277, ForwardSolver2Delastic.cpp:163-208, ForwardSolver3Dviscoelastic.cpp:188-262
    void velocityElastic()
    {
        const bool fs = d.free_surface == 1, d3 = d.dim == 3;
        vector<T> &vX = F("VX"), &vY = F("VY");
        vector<T> &Sxx = F("Sxx"), &Syy = F("Syy"), &Sxy = F("Sxy");
        // vx
        spmv(Dxf, Sxx, update);
        applyCPML(update, "sxx_x", px, true);
        spmv(fs ? DybFS : Dyb, Sxy, update_temp);
        applyCPML(update_temp, "sxy_y", py, false);
        vadd(update, update_temp);
        if (d3) {
            spmv(Dzb, F("Sxz"), update_temp);
            applyCPML(update_temp, "sxz_z", pz, false);
            vadd(update, update_temp);
        }
        vmul(update, M("inverseDensityAverageX"));
        vadd(vX, update);
        // vy
        spmv(Dxb, Sxy, update);
        applyCPML(update, "sxy_x", px, false);
        spmv(fs ? DyfFS : Dyf, Syy, update_temp);
        applyCPML(update_temp, "syy_y", py, true);
        vadd(update, update_temp);
        if (d3) {
            spmv(Dzb, F("Syz"), update_temp);
            applyCPML(update_temp, "syz_z", pz, false);
            vadd(update, update_temp);
        }
        vmul(update, M("inverseDensityAverageY"));
        vadd(vY, update);
        // vz
        if (d3) {
            vector<T> &vZ = F("VZ");
            spmv(Dxb, F("Sxz"), update);
            applyCPML(update, "sxz_x", px, false);
            spmv(fs ? DybFS : Dyb, F("Syz"), update_temp);
            applyCPML(update_temp, "syz_y", py, false);
            vadd(update, update_temp);
            spmv(Dzf, F("Szz"), update_temp);
            applyCPML(update_temp, "szz_z", pz, true);
            vadd(update, update_temp);
            vmul(update, M("inverseDensityAverageZ"));
            vadd(vZ, update);
        }
    }

    void normalStrainRates()
    {
        const bool d3 = d.dim == 3;
        spmv(Dxb, F("VX"), vxx);
        spmv(Dyb, F("VY"), vyy);
        if (d3)
            spmv(Dzb, F("VZ"), vzz);
        applyCPML(vxx, "vxx", px, false);
        applyCPML(vyy, "vyy", py, false);
        if (d3)
            applyCPML(vzz, "vzz", pz, false);
    }

    // ForwardSolver3Delastic.cpp:120-414, ForwardSolver2Delastic.cpp:117-294
    void stepElastic(Idx t)
    {
        const bool fs = d.free_surface == 1, d3 = d.dim == 3;
        velocityElastic();
        vector<T> &Sxx = F("Sxx"), &Syy = F("Syy"), &Sxy = F("Sxy");
        const vector<T> &pw = M("pWaveModulus"), &sw = M("sWaveModulus");
        normalStrainRates();
        if (d3) {
            vector<T> &Szz = F("Szz");
            vset(update, vxx);
            vadd(update, vyy);
            vadd(update, vzz);
            vmul(update, pw);
            vadd(Sxx, update);
            vadd(Syy, update);
            vadd(Szz, update);
            vsum(update, vyy, vzz);
            vmul(update, sw);
            vaxmy(Sxx, (T)2.0, update);
            vsum(update, vxx, vzz);
            vmul(update, sw);
            vaxmy(Syy, (T)2.0, update);
            vsum(update, vxx, vyy);
            vmul(update, sw);
            vaxmy(Szz, (T)2.0, update);
        } else {
            vset(update, vxx);
            vadd(update, vyy);
            vmul(update, pw);
            vadd(Sxx, update);
            vadd(Syy, update);
            vset(update, vyy);
            vmul(update, sw);
            vaxmy(Sxx, (T)2.0, update);
            vset(update, vxx);
            vmul(update, sw);
            vaxmy(Syy, (T)2.0, update);
        }
        // shear
        spmv(Dyf, F("VX"), update);
        applyCPML(update, "vxy", py, true);
        spmv(Dxf, F("VY"), update_temp);
        applyCPML(update_temp, "vyx", px, true);
        vadd(update, update_temp);
        vmul(update, M("sWaveModulusAverageXY"));
        vadd(Sxy, update);
        if (d3) {
            spmv(Dzf, F("VX"), update);
            applyCPML(update, "vxz", pz, true);
            spmv(Dxf, F("VZ"), update_temp);
            applyCPML(update_temp, "vzx", px, true);
            vadd(update, update_temp);
            vmul(update, M("sWaveModulusAverageXZ"));
            vadd(F("Sxz"), update);

            spmv(Dzf, F("VY"), update);
            applyCPML(update, "vyz", pz, true);
            spmv(Dyf, F("VZ"), update_temp);
            applyCPML(update_temp, "vzy", py, true);
            vadd(update, update_temp);
            vmul(update, M("sWaveModulusAverageYZ"));
            vadd(F("Syz"), update);
        }
        if (fs) {
            // FreeSurface3Delastic.cpp:15-47 / FreeSurface2Delastic.cpp:14-46
            if (d3) {
                vsum(update, vxx, vzz);
                setSurfaceZero(Syy);
                vector<T> &Szz = F("Szz");
                for (size_t k = 0; k < surfIdx.size(); k++) {
                    Idx i = surfIdx[k];
                    T temp = sH[k] * update[i];
                    Sxx[i] = Sxx[i] + temp;
                    Szz[i] = Szz[i] + temp;
                    temp = sV[k] * vyy[i];
                    Sxx[i] = Sxx[i] - temp;
                    Szz[i] = Szz[i] - temp;
                }
            } else {
                for (size_t k = 0; k < surfIdx.size(); k++) {
                    Idx i = surfIdx[k];
                    T temp = sH[k] * vxx[i];
                    Sxx[i] = Sxx[i] + temp;
                    temp = sV[k] * vyy[i];
                    Sxx[i] = Sxx[i] - temp;
                }
                setSurfaceZero(Syy);
            }
        }
        if (d3)
            absApply({"Sxx", "Syy", "Szz", "Sxy", "Sxz", "Syz", "VX", "VY", "VZ"});
        else
            absApply({"Sxx", "Syy", "Sxy", "VX", "VY"});
        applySource(t);
        gatherSeismogram(t);
    }

    // one shear component of the viscoelastic update, ForwardSolver3Dviscoelastic.cpp:355-416
    void viscoShear(vector<T> &S, const char *Rbase, const vector<T> &muAvg, const vector<T> &tauAvg)
    {
        vmul(update, muAvg);
        for (int l = 0; l < L; l++) {
            vector<T> &R = F(Rn(Rbase, l));
            vaxpy(S, DThalf, R);
            vscale(R, viscoCoeff1[l]);
            vscaled(update2, inverseRelaxationTime[l], update);
            vmul(update2, tauAvg);
            vsub(R, update2);
            vscale(R, viscoCoeff2[l]);
            vaxpy(S, DThalf, R);
        }
        vmul(update, onePlusLtauS);
        vadd(S, update);
    }
    // one normal component, second half: ForwardSolver3Dviscoelastic.cpp:307-352
    void viscoNormalShearPart(vector<T> &S, const char *Rbase)
    {
        vmul(update, M("sWaveModulus"));
        vscale(update, (T)2.0);
        for (int l = 0; l < L; l++) {
            vector<T> &R = F(Rn(Rbase, l));
            vscaled(update2, inverseRelaxationTime[l], update);
            vmul(update2, M("tauS"));
            vadd(R, update2);
            vscale(R, viscoCoeff2[l]);
            vaxpy(S, DThalf, R);
        }
        vmul(update, onePlusLtauS);
        vsub(S, update);
    }

    // ForwardSolver3Dviscoelastic.cpp:132-454, ForwardSolver2Dviscoelastic.cpp:130-318
    void stepViscoelastic(Idx t)
    {
        const bool fs = d.free_surface == 1, d3 = d.dim == 3;
        velocityElastic();
        vector<T> &Sxx = F("Sxx"), &Syy = F("Syy"), &Sxy = F("Sxy");
        normalStrainRates();
        vset(update, vxx);
        vadd(update, vyy);
        if (d3)
            vadd(update, vzz);
        vmul(update, M("pWaveModulus"));
        for (int l = 0; l < L; l++) {
            vscaled(update2, inverseRelaxationTime[l], update);
            vmul(update2, M("tauP"));
            vaxpy(Sxx, DThalf, F(Rn("Rxx", l)));
            vscale(F(Rn("Rxx", l)), viscoCoeff1[l]);
            vsub(F(Rn("Rxx", l)), update2);
            vaxpy(Syy, DThalf, F(Rn("Ryy", l)));
            vscale(F(Rn("Ryy", l)), viscoCoeff1[l]);
            vsub(F(Rn("Ryy", l)), update2);
            if (d3) {
                vaxpy(F("Szz"), DThalf, F(Rn("Rzz", l)));
                vscale(F(Rn("Rzz", l)), viscoCoeff1[l]);
                vsub(F(Rn("Rzz", l)), update2);
            }
        }
        vmul(update, onePlusLtauP);
        vadd(Sxx, update);
        vadd(Syy, update);
        if (d3)
            vadd(F("Szz"), update);
        if (d3) {
            vsum(update, vyy, vzz);
            viscoNormalShearPart(Sxx, "Rxx");
            vsum(update, vxx, vzz);
            viscoNormalShearPart(Syy, "Ryy");
            vsum(update, vxx, vyy);
            viscoNormalShearPart(F("Szz"), "Rzz");
        } else {
            vset(update, vyy);
            viscoNormalShearPart(Sxx, "Rxx");
            vset(update, vxx);
            viscoNormalShearPart(Syy, "Ryy");
        }
        // shear
        spmv(Dyf, F("VX"), update);
        applyCPML(update, "vxy", py, true);
        spmv(Dxf, F("VY"), update_temp);
        applyCPML(update_temp, "vyx", px, true);
        vadd(update, update_temp);
        viscoShear(Sxy, "Rxy", M("sWaveModulusAverageXY"), M("tauSAverageXY"));
        if (d3) {
            spmv(Dzf, F("VX"), update);
            applyCPML(update, "vxz", pz, true);
            spmv(Dxf, F("VZ"), update_temp);
            applyCPML(update_temp, "vzx", px, true);
            vadd(update, update_temp);
            viscoShear(F("Sxz"), "Rxz", M("sWaveModulusAverageXZ"), M("tauSAverageXZ"));
            spmv(Dzf, F("VY"), update);
            applyCPML(update, "vyz", pz, true);
            spmv(Dyf, F("VZ"), update_temp);
            applyCPML(update_temp, "vzy", py, true);
            vadd(update, update_temp);
            viscoShear(F("Syz"), "Ryz", M("sWaveModulusAverageYZ"), M("tauSAverageYZ"));
        }
        if (fs) {
            // FreeSurface3Dviscoelastic.cpp:17-75, FreeSurface2Dviscoelastic.cpp:15-63
            if (d3)
                vsum(update, vxx, vzz);
            const vector<T> &hor = d3 ? update : vxx;
            const int ncomp = d3 ? 2 : 1;
            vector<T> *S[2] = {&Sxx, d3 ? &F("Szz") : nullptr};
            const char *Rb[2] = {"Rxx", "Rzz"};
            for (size_t k = 0; k < surfIdx.size(); k++) {
                Idx i = surfIdx[k];
                for (int l = 0; l < L; l++)
                    for (int c = 0; c < ncomp; c++) {
                        T temp = (T)1 * F(Rn(Rb[c], l))[i];
                        (*S[c])[i] = (*S[c])[i] - DThalf * temp;
                    }
                T temp = sSH[k] * hor[i];
                for (int c = 0; c < ncomp; c++)
                    (*S[c])[i] = (*S[c])[i] + temp;
                temp = sSV[k] * vyy[i];
                for (int c = 0; c < ncomp; c++)
                    (*S[c])[i] = (*S[c])[i] - temp;
                for (int l = 0; l < L; l++) {
                    T th = sRH[l][k] * hor[i];
                    for (int c = 0; c < ncomp; c++) {
                        vector<T> &R = F(Rn(Rb[c], l));
                        R[i] = R[i] + th;
                    }
                    T tv = sRV[l][k] * vyy[i];
                    for (int c = 0; c < ncomp; c++) {
                        vector<T> &R = F(Rn(Rb[c], l));
                        R[i] = R[i] - tv;
                    }
                    for (int c = 0; c < ncomp; c++) {
                        T tt = (T)1 * F(Rn(Rb[c], l))[i];
                        (*S[c])[i] = (*S[c])[i] + DThalf * tt;
                    }
                }
            }
            setSurfaceZero(Syy);
            for (int l = 0; l < L; l++)
                setSurfaceZero(F(Rn("Ryy", l)));
        }
        if (d3)
            absApply({"Sxx", "Syy", "Szz", "Sxy", "Sxz", "Syz", "VX", "VY", "VZ"});
        else
            absApply({"Sxx", "Syy", "Sxy", "VX", "VY"});
        applySource(t);
        gatherSeismogram(t);
    }

    // ForwardSolver2Dsh.cpp:108-193, ForwardSolver2Dviscosh.cpp:126-241
    void stepSH(Idx t)
    {
        const bool fs = d.free_surface == 1;
        vector<T> &vZ = F("VZ"), &Sxz = F("Sxz"), &Syz = F("Syz");
        spmv(Dxb, Sxz, update);
        spmv(fs ? DybFS : Dyb, Syz, update_temp);
        applyCPML(update, "sxz_x", px, false);
        applyCPML(update_temp, "syz_y", py, false);
        vadd(update, update_temp);
        vmul(update, M("inverseDensity"));
        vadd(vZ, update);
        spmv(Dxf, vZ, update);
        applyCPML(update, "vzx", px, true);
        if (d.eq == WS_EQ_SH) {
            vmul(update, M("sWaveModulusAverageXZ"));
            vadd(Sxz, update);
        } else
            viscoShear(Sxz, "Rxz", M("sWaveModulusAverageXZ"), M("tauSAverageXZ"));
        spmv(Dyf, vZ, update);
        applyCPML(update, "vzy", py, true);
        if (d.eq == WS_EQ_SH) {
            vmul(update, M("sWaveModulusAverageYZ"));
            vadd(Syz, update);
        } else
            viscoShear(Syz, "Ryz", M("sWaveModulusAverageYZ"), M("tauSAverageYZ"));
        absApply({"Sxz", "Syz", "VZ"});
        applySource(t);
        gatherSeismogram(t);
    }

    // ForwardSolver2Dtmem.cpp:108-171, ForwardSolver2Dviscotmem.cpp:130-207; CPML map CPMLEM2D.cpp:21-73
    void stepTMEM(Idx t)
    {
        vector<T> &hX = F("HX"), &hY = F("HY"), &eZ = F("EZ");
        spmv(Dyf, eZ, update);
        applyCPML(update, "ezy", py, true);
        vmul(update, M("inverseMagneticPermeabilityAverageYZ"));
        vsub(hX, update);
        spmv(Dxf, eZ, update_temp);
        applyCPML(update_temp, "ezx", px, true);
        vscaled(update, (T)-1, update_temp);
        vmul(update, M("inverseMagneticPermeabilityAverageXZ"));
        vsub(hY, update);
        for (int l = 0; l < L; l++) {
            vector<T> &r = F(Rn("RZ", l));
            vscaled(update, Cc[l], r);
            vset(update_temp, M("CdAverageZ" + std::to_string(l + 1)));
            vmul(update_temp, eZ);
            vsum(r, update_temp, update);
        }
        spmv(Dxb, hY, update);
        spmv(Dyb, hX, update_temp);
        applyCPML(update, "hyx", px, false);
        applyCPML(update_temp, "hxy", py, false);
        vsub(update, update_temp);
        for (int l = 0; l < L; l++)
            vaxmy(update, DT, F(Rn("RZ", l)));
        vmul(update, M("CbAverageZ"));
        vset(update_temp, M("CaAverageZ"));
        vmul(update_temp, eZ);
        vsum(eZ, update_temp, update);
        absApply({"EZ", "HX", "HY"});
        applySource(t);
        gatherSeismogram(t);
    }

    // r_l = Cc_l * r_l + Cd_l (.) e   (ForwardSolver2Dviscoemem.cpp:186-193, ForwardSolver3Dviscoemem.cpp:236-246)
    void relaxEM(const char *Rbase, const char *CdBase, const vector<T> &e)
    {
        for (int l = 0; l < L; l++) {
            vector<T> &r = F(Rn(Rbase, l));
            vscaled(update, Cc[l], r);
            vset(update_temp, M(std::string(CdBase) + std::to_string(l + 1)));
            vmul(update_temp, e);
            vsum(r, update_temp, update);
        }
    }
    // e = Ca (.) e + Cb (.) (update - DT * sum_l r_l)
    void updateE(vector<T> &e, const char *Rbase, const char *Ca, const char *Cb)
    {
        for (int l = 0; l < L; l++)
            vaxmy(update, DT, F(Rn(Rbase, l)));
        vmul(update, M(Cb));
        vset(update_temp, M(Ca));
        vmul(update_temp, e);
        vsum(e, update_temp, update);
    }

    // ForwardSolver2Demem.cpp:112-176, ForwardSolver2Dviscoemem.cpp:142-228; CPML map CPMLEM2D.cpp:21-73
    void stepEMEM2D(Idx t)
    {
        vector<T> &hZ = F("HZ"), &eX = F("EX"), &eY = F("EY");
        spmv(Dxf, eY, update);
        spmv(Dyf, eX, update_temp);
        applyCPML(update, "eyx", px, true);
        applyCPML(update_temp, "exy", py, true);
        vsub(update, update_temp);
        vmul(update, M("inverseMagneticPermeabilityAverageXY"));
        vsub(hZ, update);
        for (int l = 0; l < L; l++) { // interleaved X/Y per mechanism as in the reference
            vector<T> &rx = F(Rn("RX", l)), &ry = F(Rn("RY", l));
            vscaled(update, Cc[l], rx);
            vset(update_temp, M("CdAverageX" + std::to_string(l + 1)));
            vmul(update_temp, eX);
            vsum(rx, update_temp, update);
            vscaled(update, Cc[l], ry);
            vset(update_temp, M("CdAverageY" + std::to_string(l + 1)));
            vmul(update_temp, eY);
            vsum(ry, update_temp, update);
        }
        spmv(Dyb, hZ, update);
        applyCPML(update, "hzy", py, false);
        updateE(eX, "RX", "CaAverageX", "CbAverageX");
        // non-visco: eY = Ca*eY - Cb*(Dxb hZ); visco: update = -(Dxb hZ) - DT*r; eY = Ca*eY + Cb*update  (bit-identical forms)
        spmv(Dxb, hZ, update_temp);
        applyCPML(update_temp, "hzx", px, false);
        vscaled(update, (T)-1, update_temp);
        updateE(eY, "RY", "CaAverageY", "CbAverageY");
        absApply({"EY", "EX", "HZ"});
        applySource(t);
        gatherSeismogram(t);
    }

    // ForwardSolver3Demem.cpp:119-240, ForwardSolver3Dviscoemem.cpp:161-312; CPML map CPMLEM3D.cpp:27-104
    // (all E-derivatives use the half profile; H-derivatives the full profile except hyx, CPMLEM3D.cpp:69)
    void stepEMEM3D(Idx t)
    {
        vector<T> &hX = F("HX"), &hY = F("HY"), &hZ = F("HZ"), &eX = F("EX"), &eY = F("EY"), &eZ = F("EZ");
        spmv(Dyf, eZ, update);
        spmv(Dzf, eY, update_temp);
        applyCPML(update, "ezy", py, true);
        applyCPML(update_temp, "eyz", pz, true);
        vsub(update, update_temp);
        vmul(update, M("inverseMagneticPermeabilityAverageYZ"));
        vsub(hX, update);
        spmv(Dzf, eX, update);
        spmv(Dxf, eZ, update_temp);
        applyCPML(update, "exz", pz, true);
        applyCPML(update_temp, "ezx", px, true);
        vsub(update, update_temp);
        vmul(update, M("inverseMagneticPermeabilityAverageXZ"));
        vsub(hY, update);
        spmv(Dxf, eY, update);
        spmv(Dyf, eX, update_temp);
        applyCPML(update, "eyx", px, true);
        applyCPML(update_temp, "exy", py, true);
        vsub(update, update_temp);
        vmul(update, M("inverseMagneticPermeabilityAverageXY"));
        vsub(hZ, update);
        for (int l = 0; l < L; l++) {
            const char *rb[3] = {"RX", "RY", "RZ"};
            const char *cd[3] = {"CdAverageX", "CdAverageY", "CdAverageZ"};
            vector<T> *e[3] = {&eX, &eY, &eZ};
            for (int c = 0; c < 3; c++) {
                vector<T> &r = F(Rn(rb[c], l));
                vscaled(update, Cc[l], r);
                vset(update_temp, M(std::string(cd[c]) + std::to_string(l + 1)));
                vmul(update_temp, *e[c]);
                vsum(r, update_temp, update);
            }
        }
        spmv(Dyb, hZ, update);
        spmv(Dzb, hY, update_temp);
        applyCPML(update, "hzy", py, false);
        applyCPML(update_temp, "hyz", pz, false);
        vsub(update, update_temp);
        updateE(eX, "RX", "CaAverageX", "CbAverageX");
        spmv(Dzb, hX, update);
        spmv(Dxb, hZ, update_temp);
        applyCPML(update, "hxz", pz, false);
        applyCPML(update_temp, "hzx", px, false);
        vsub(update, update_temp);
        updateE(eY, "RY", "CaAverageY", "CbAverageY");
        spmv(Dxb, hY, update);
        spmv(Dyb, hX, update_temp);
        applyCPML(update, "hyx", px, true);
        applyCPML(update_temp, "hxy", py, false);
        vsub(update, update_temp);
        updateE(eZ, "RZ", "CaAverageZ", "CbAverageZ");
        absApply({"EZ", "EY", "EX", "HX", "HY", "HZ"});
        applySource(t);
        gatherSeismogram(t);
    }

    // ------------------------------------------------------------------------------------------------------------
    // Second CPU baseline ("the fused matrix-free CPU oracle as a second, stronger CPU number"): the 3-D elastic step
    // with the operators applied as 1-D coefficient rows (taken from the assembled matrices: a row of D_x depends on x only, of D_y
    // on y, of D_z on z) and all statements of a half-step fused per grid point.  Same operations in the same order as stepElastic,
    // so the wavefields are bit-identical to the matrix formulation (tests/test_oracle_golden.py::test_fused_backend_*).
    // ------------------------------------------------------------------------------------------------------------
    bool fused = false, fusedReady = false;
    struct Rows1D {
        vector<int> cnt, off;
        vector<T> val;
    };
    Rows1D fXf, fXb, fYf, fYb, fYfS, fYbS, fZf, fZb;
    vector<Idx> fkx, fky, fkz; // grid point -> entry of the CPML pattern of the axis (-1 outside the layer)

    Rows1D rows1d(const Csr<T> &A, Idx count, Idx stride) const
    {
        Rows1D R;
        R.cnt.assign(count, 0);
        R.off.assign((size_t)count * MAXROW, 0);
        R.val.assign((size_t)count * MAXROW, (T)0);
        if (A.empty())
            return R;
        for (Idx c = 0; c < count; c++) {
            const Idx i = c * stride;
            int k = 0;
            for (int64_t e = A.ia[i]; e < A.ia[i + 1]; e++, k++) {
                const Idx delta = A.ja[e] - i;
                ORACLE_REQUIRE(k < MAXROW && delta % stride == 0, "fused back-end: operator row is not one-dimensional");
                R.off[(size_t)c * MAXROW + k] = (int)(delta / stride);
                R.val[(size_t)c * MAXROW + k] = A.va[e];
            }
            R.cnt[c] = k;
        }
        return R;
    }
    void fusedSetup()
    {
        ORACLE_REQUIRE(d.eq == WS_EQ_ELASTIC && d.dim == 3 && !vg.active && d.damping != 1,
                       "the fused back-end serves the 3-D elastic solver on a regular grid without an ABS frame");
        const Idx plane = NX * NZ;
        fXf = rows1d(Dxf, NX, 1);
        fXb = rows1d(Dxb, NX, 1);
        fYf = rows1d(Dyf, NY, plane);
        fYb = rows1d(Dyb, NY, plane);
        fYfS = rows1d(DyfFS, NY, plane);
        fYbS = rows1d(DybFS, NY, plane);
        fZf = rows1d(Dzf, NZ, NX);
        fZb = rows1d(Dzb, NZ, NX);
        auto mapOf = [&](const Profile<T> &p, vector<Idx> &m) {
            m.assign(d.damping == 2 ? N : 0, (Idx)-1);
            if (d.damping == 2)
                for (size_t k = 0; k < p.idx.size(); k++)
                    m[p.idx[k]] = (Idx)k;
        };
        mapOf(px, fkx);
        mapOf(py, fky);
        mapOf(pz, fkz);
        fusedReady = true;
    }
    // out += w * src over a line: the one hot loop of the derivative rows, compiled for the vector units the host has (function
    // multi-versioning; the arithmetic stays one multiplication and one addition per element, no contraction)
    __attribute__((target_clones("avx512f", "avx2", "default"), optimize("O3"))) static void axpyLine(T *__restrict__ out, T w, const T *__restrict__ src, Idx n)
    {
        for (Idx k = 0; k < n; k++)
            out[k] += w * src[k];
    }
    // v += ri * ((a + b) + c)  and  S += mu * (a + b)  over a line (same operations in the same order as the vector statements)
    __attribute__((target_clones("avx512f", "avx2", "default"), optimize("O3"))) static void combine3(T *__restrict__ v, const T *__restrict__ a, const T *__restrict__ b,
                                                                                                     const T *__restrict__ c, const T *__restrict__ ri, Idx n)
    {
        for (Idx k = 0; k < n; k++) {
            T u = a[k] + b[k];
            u = u + c[k];
            u = u * ri[k];
            v[k] = v[k] + u;
        }
    }
    __attribute__((target_clones("avx512f", "avx2", "default"), optimize("O3"))) static void combine2(T *__restrict__ S, const T *__restrict__ a, const T *__restrict__ b,
                                                                                                     const T *__restrict__ mu, Idx n)
    {
        for (Idx k = 0; k < n; k++) {
            T u = a[k] + b[k];
            u = u * mu[k];
            S[k] = S[k] + u;
        }
    }
    // applyCPML over consecutive entries: temp = a; psi *= b; temp *= u; psi += temp; u += psi
    __attribute__((target_clones("avx512f", "avx2", "default"), optimize("O3"))) static void cpRun(T *__restrict__ u, T *__restrict__ q, const T *__restrict__ A, const T *__restrict__ B, Idx n)
    {
        for (Idx k = 0; k < n; k++) {
            T temp = A[k];
            q[k] = q[k] * B[k];
            temp = temp * u[k];
            q[k] = q[k] + temp;
            u[k] = u[k] + q[k];
        }
    }
    // one x-line of a derivative into a line buffer; taps outermost so that the loops over x vectorise, every x accumulating its taps
    // in ascending column order from 0 like the sparse matrix-vector product
    static void lineYZ(const Rows1D &R, Idx c, const T *x, Idx i0, Idx stride, Idx n, T *__restrict__ out)
    {
        const int *o = &R.off[(size_t)c * MAXROW];
        const T *v = &R.val[(size_t)c * MAXROW];
        for (Idx k = 0; k < n; k++)
            out[k] = (T)0;
        for (int t = 0; t < R.cnt[c]; t++)
            axpyLine(out, v[t], x + i0 + (Idx)o[t] * stride, n);
    }
    // rows of D_x: identical between the edge zones [0, fEdge) and [NX - fEdge, NX)
    static inline void lineX(const Rows1D &R, Idx fEdge, const T *x, Idx i0, Idx n, T *out)
    {
        auto scalar = [&](Idx c) {
            const int *o = &R.off[(size_t)c * MAXROW];
            const T *v = &R.val[(size_t)c * MAXROW];
            T sum = 0;
            for (int t = 0; t < R.cnt[c]; t++)
                sum += v[t] * x[i0 + c + o[t]];
            out[c] = sum;
        };
        if (n <= 2 * fEdge) {
            for (Idx c = 0; c < n; c++)
                scalar(c);
            return;
        }
        for (Idx c = 0; c < fEdge; c++)
            scalar(c);
        for (Idx c = n - fEdge; c < n; c++)
            scalar(c);
        const int *o = &R.off[(size_t)fEdge * MAXROW];
        const T *v = &R.val[(size_t)fEdge * MAXROW];
        const int cnt = R.cnt[fEdge];
        for (Idx c = fEdge; c < n - fEdge; c++)
            out[c] = (T)0;
        for (int t = 0; t < cnt; t++)
            axpyLine(out + fEdge, v[t], x + i0 + o[t] + fEdge, n - 2 * fEdge);
    }
    // applyCPML on an x-line.  y / z axis: the whole line lies in the layer or not, its entries are consecutive in the pattern;
    // x axis: the two ends of the line.   temp = a; psi *= b; temp *= u; psi += temp; u += psi
    static inline void cpEntry(T &u, T *ps, const Profile<T> &p, Idx k, bool half)
    {
        T temp = half ? p.ah[k] : p.a[k];
        ps[k] = ps[k] * (half ? p.bh[k] : p.b[k]);
        temp = temp * u;
        ps[k] = ps[k] + temp;
        u = u + ps[k];
    }
    static inline void cpLine(T *u, T *ps, const Profile<T> &p, const vector<Idx> &map, Idx i0, Idx n, bool half)
    {
        if (map.empty())
            return;
        const Idx k0 = map[i0];
        if (k0 < 0)
            return;
        cpRun(u, ps + k0, (half ? p.ah : p.a).data() + k0, (half ? p.bh : p.b).data() + k0, n);
    }
    static inline void cpLineX(T *u, T *ps, const Profile<T> &p, const vector<Idx> &map, Idx i0, Idx n, Idx W, bool half)
    {
        if (map.empty())
            return;
        for (Idx c = 0; c < n; c++) {
            if (c == W && n > 2 * W)
                c = n - W;
            const Idx k = map[i0 + c];
            if (k >= 0)
                cpEntry(u[c], ps, p, k, half);
        }
    }
    Idx fEdge = 0, fW = 0;
    void fusedCheckLayout()
    {
        // (a) rows of the x operators are identical away from the edges; (b) a line of the y / z layer is consecutive in the pattern
        fEdge = std::min<Idx>(NX, (Idx)(d.fd_order / 2 + 1));
        for (const Rows1D *R : {&fXf, &fXb})
            for (Idx c = fEdge; c + fEdge < NX; c++) {
                ORACLE_REQUIRE(R->cnt[c] == R->cnt[fEdge], "fused back-end: interior rows of D_x differ");
                for (int t = 0; t < R->cnt[c]; t++)
                    ORACLE_REQUIRE(R->off[(size_t)c * MAXROW + t] == R->off[(size_t)fEdge * MAXROW + t] && R->val[(size_t)c * MAXROW + t] == R->val[(size_t)fEdge * MAXROW + t],
                                   "fused back-end: interior rows of D_x differ");
            }
        fW = d.damping == 2 ? (Idx)d.boundary_width : 0;
        if (d.damping == 2) {
            for (const vector<Idx> *m : {&fky, &fkz})
                for (Idx i0 = 0; i0 < N; i0 += NX)
                    for (Idx c = 1; c < NX; c++)
                        ORACLE_REQUIRE(((*m)[i0] < 0 && (*m)[i0 + c] < 0) || ((*m)[i0] >= 0 && (*m)[i0 + c] == (*m)[i0] + c), "fused back-end: CPML pattern is not line-wise");
            for (Idx i0 = 0; i0 < N; i0 += NX)
                for (Idx c = fW; c + fW < NX; c++)
                    ORACLE_REQUIRE(fkx[i0 + c] < 0, "fused back-end: x layer wider than BoundaryWidth");
        }
    }
    void stepElasticFused(Idx t)
    {
        if (!fusedReady) {
            fusedSetup();
            fusedCheckLayout();
        }
        const bool fs = d.free_surface == 1;
        const Idx plane = NX * NZ, n = NX;
        T *vX = F("VX").data(), *vY = F("VY").data(), *vZ = F("VZ").data();
        T *Sxx = F("Sxx").data(), *Syy = F("Syy").data(), *Szz = F("Szz").data(), *Sxy = F("Sxy").data(), *Sxz = F("Sxz").data(), *Syz = F("Syz").data();
        const T *rix = M("inverseDensityAverageX").data(), *riy = M("inverseDensityAverageY").data(), *riz = M("inverseDensityAverageZ").data();
        const T *pw = M("pWaveModulus").data(), *sw = M("sWaveModulus").data();
        const T *mxy = M("sWaveModulusAverageXY").data(), *mxz = M("sWaveModulusAverageXZ").data(), *myz = M("sWaveModulusAverageYZ").data();
        auto P = [&](const char *name, const Profile<T> &p) -> T * { return d.damping == 2 ? psiOf(name, p).data() : nullptr; };
        T *p_sxx_x = P("sxx_x", px), *p_sxy_y = P("sxy_y", py), *p_sxz_z = P("sxz_z", pz), *p_sxy_x = P("sxy_x", px), *p_syy_y = P("syy_y", py), *p_syz_z = P("syz_z", pz);
        T *p_sxz_x = P("sxz_x", px), *p_syz_y = P("syz_y", py), *p_szz_z = P("szz_z", pz);
        T *p_vxx = P("vxx", px), *p_vyy = P("vyy", py), *p_vzz = P("vzz", pz), *p_vxy = P("vxy", py), *p_vyx = P("vyx", px), *p_vxz = P("vxz", pz), *p_vzx = P("vzx", px);
        T *p_vyz = P("vyz", pz), *p_vzy = P("vzy", py);
        const Rows1D &Yb1 = fs ? fYbS : fYb, &Yf1 = fs ? fYfS : fYf;
        const T *sHp = sH.data(), *sVp = sV.data();
#pragma omp parallel
        {
            vector<T> bufA(n), bufB(n), bufC(n);
            T *a = bufA.data(), *b = bufB.data(), *c = bufC.data();
            // particle velocities (velocityElastic): v_a += rho_a^-1 ((P(D_x .) + P(D_y .)) + P(D_z .))
            auto velocity = [&](Idx y, Idx z, Idx i0, const Rows1D &RX, const T *fx, T *psx, bool hx, const Rows1D &RY, const T *fy, T *psy, bool hy, const Rows1D &RZ,
                                const T *fz, T *psz, bool hz, const T *ri, T *v) {
                lineX(RX, fEdge, fx, i0, n, a);
                cpLineX(a, psx, px, fkx, i0, n, fW, hx);
                lineYZ(RY, y, fy, i0, plane, n, b);
                cpLine(b, psy, py, fky, i0, n, hy);
                lineYZ(RZ, z, fz, i0, NX, n, c);
                cpLine(c, psz, pz, fkz, i0, n, hz);
                combine3(v + i0, a, b, c, ri + i0, n);
            };
#pragma omp for collapse(2) schedule(static)
            for (Idx y = 0; y < NY; y++)
                for (Idx z = 0; z < NZ; z++) {
                    const Idx i0 = y * plane + z * NX;
                    velocity(y, z, i0, fXf, Sxx, p_sxx_x, true, Yb1, Sxy, p_sxy_y, false, fZb, Sxz, p_sxz_z, false, rix, vX);
                    velocity(y, z, i0, fXb, Sxy, p_sxy_x, false, Yf1, Syy, p_syy_y, true, fZb, Syz, p_syz_z, false, riy, vY);
                    velocity(y, z, i0, fXb, Sxz, p_sxz_x, false, Yb1, Syz, p_syz_y, false, fZf, Szz, p_szz_z, true, riz, vZ);
                }
            // stresses (normalStrainRates, stepElastic incl. the free surface); the implicit barrier of the loop above orders the passes
            auto shear = [&](Idx i0, Idx c1, const Rows1D &R1, const T *f1, Idx stride1, T *ps1, const Profile<T> &pr1, const vector<Idx> &m1, bool x1, Idx c2, const Rows1D &R2,
                             const T *f2, Idx stride2, T *ps2, const Profile<T> &pr2, const vector<Idx> &m2, bool x2, const T *mu, T *S) {
                if (x1) {
                    lineX(R1, fEdge, f1, i0, n, a);
                    cpLineX(a, ps1, pr1, m1, i0, n, fW, true);
                } else {
                    lineYZ(R1, c1, f1, i0, stride1, n, a);
                    cpLine(a, ps1, pr1, m1, i0, n, true);
                }
                if (x2) {
                    lineX(R2, fEdge, f2, i0, n, b);
                    cpLineX(b, ps2, pr2, m2, i0, n, fW, true);
                } else {
                    lineYZ(R2, c2, f2, i0, stride2, n, b);
                    cpLine(b, ps2, pr2, m2, i0, n, true);
                }
                combine2(S + i0, a, b, mu + i0, n);
            };
#pragma omp for collapse(2) schedule(static)
            for (Idx y = 0; y < NY; y++)
                for (Idx z = 0; z < NZ; z++) {
                    const Idx i0 = y * plane + z * NX;
                    lineX(fXb, fEdge, vX, i0, n, a);
                    lineYZ(fYb, y, vY, i0, plane, n, b);
                    lineYZ(fZb, z, vZ, i0, NX, n, c);
                    cpLineX(a, p_vxx, px, fkx, i0, n, fW, false);
                    cpLine(b, p_vyy, py, fky, i0, n, false);
                    cpLine(c, p_vzz, pz, fkz, i0, n, false);
                    const bool surface = fs && y == 0;
                    for (Idx k = 0; k < n; k++) {
                        const Idx i = i0 + k;
                        const T exx = a[k], eyy = b[k], ezz = c[k];
                        T u = exx + eyy;
                        u = u + ezz;
                        u = u * pw[i];
                        T sxx = Sxx[i] + u, syy = Syy[i] + u, szz = Szz[i] + u;
                        u = eyy + ezz;
                        u = u * sw[i];
                        sxx = sxx - (T)2.0 * u;
                        u = exx + ezz;
                        u = u * sw[i];
                        syy = syy - (T)2.0 * u;
                        u = exx + eyy;
                        u = u * sw[i];
                        szz = szz - (T)2.0 * u;
                        if (surface) { // FreeSurface3Delastic.cpp:15-47 (surface point k = x + z NX = i)
                            u = exx + ezz;
                            syy = syy * (T)0;
                            T temp = sHp[i] * u;
                            sxx = sxx + temp;
                            szz = szz + temp;
                            temp = sVp[i] * eyy;
                            sxx = sxx - temp;
                            szz = szz - temp;
                        }
                        Sxx[i] = sxx;
                        Syy[i] = syy;
                        Szz[i] = szz;
                    }
                    shear(i0, y, fYf, vX, plane, p_vxy, py, fky, false, 0, fXf, vY, 1, p_vyx, px, fkx, true, mxy, Sxy);
                    shear(i0, z, fZf, vX, NX, p_vxz, pz, fkz, false, 0, fXf, vZ, 1, p_vzx, px, fkx, true, mxz, Sxz);
                    shear(i0, z, fZf, vY, NX, p_vyz, pz, fkz, false, y, fYf, vZ, plane, p_vzy, py, fky, false, myz, Syz);
                }
        }
        applySource(t);
        gatherSeismogram(t);
    }

    void step(Idx t)
    {
        ORACLE_REQUIRE(prepared, "call prepare before step");
        ORACLE_REQUIRE(t >= 0 && t < d.nt, "time step out of range");
        switch (d.eq) {
        case WS_EQ_ACOUSTIC: stepAcoustic(t); break;
        case WS_EQ_ELASTIC:
            if (fused)
                stepElasticFused(t);
            else
                stepElastic(t);
            break;
        case WS_EQ_VISCOELASTIC: stepViscoelastic(t); break;
        case WS_EQ_SH:
        case WS_EQ_VISCOSH: stepSH(t); break;
        case WS_EQ_TMEM:
        case WS_EQ_VISCOTMEM: stepTMEM(t); break;
        case WS_EQ_EMEM:
        case WS_EQ_VISCOEMEM:
            if (d.dim == 3)
                stepEMEM3D(t);
            else
                stepEMEM2D(t);
            break;
        default: throw std::runtime_error("unknown equationType");
        }
    }
};

struct Handle {
    int precision = 32;
    std::unique_ptr<Oracle<float>> f;
    std::unique_ptr<Oracle<double>> dd;
};

template <typename F>
int guard(F fn)
{
    try {
        fn();
        return WS_OK;
    } catch (const std::exception &e) {
        g_err = e.what();
        return WS_EINVAL;
    }
}

template <typename T>
void setVec(vector<T> &dst, const float *src, size_t n)
{
    dst.resize(n);
    for (size_t i = 0; i < n; i++)
        dst[i] = (T)src[i];
}
template <typename T>
void getVec(const vector<T> &src, float *dst, size_t n)
{
    ORACLE_REQUIRE(src.size() == n, "size mismatch");
    for (size_t i = 0; i < n; i++)
        dst[i] = (float)src[i];
}

} // namespace

#define DISPATCH(h, expr)                                                                                              \
    do {                                                                                                               \
        if ((h)->precision == 64) {                                                                                    \
            auto &o = *(h)->dd;                                                                                        \
            expr;                                                                                                      \
        } else {                                                                                                       \
            auto &o = *(h)->f;                                                                                         \
            expr;                                                                                                      \
        }                                                                                                              \
    } while (0)

extern "C" {

struct wso_solver;

const char *wso_last_error(void) { return g_err.c_str(); }

// precision: 32 (reference build, ValueType = float, Configuration/ValueType.hpp:5) or 64 (accuracy yardstick)
int wso_create(const ws_desc *desc, int precision, wso_solver **out)
{
    return guard([&] {
        ORACLE_REQUIRE(desc && out, "null argument");
        ORACLE_REQUIRE(precision == 32 || precision == 64, "precision must be 32 or 64");
        auto h = std::make_unique<Handle>();
        h->precision = precision;
        if (precision == 64) {
            h->dd = std::make_unique<Oracle<double>>();
            h->dd->create(*desc);
        } else {
            h->f = std::make_unique<Oracle<float>>();
            h->f->create(*desc);
        }
        *out = reinterpret_cast<wso_solver *>(h.release());
    });
}
// Variable grid / variable FD order (Coordinates.cpp:44-72: gridConfig columns interface, dhFactor, FDorder; the first interface must
// be 0).  desc->nx/ny/nz is the fine regular grid; the model vectors then hold wso_grid_size() values in the layered order and
// acquisition indices come from wso_coordinate2index().  fd_orders may be null (desc->fd_order everywhere).
int wso_create_vargrid(const ws_desc *desc, int precision, int32_t nlayers, const int32_t *interfaces, const int32_t *dh_factors, const int32_t *fd_orders, wso_solver **out)
{
    return guard([&] {
        ORACLE_REQUIRE(desc && out && nlayers >= 1 && interfaces && dh_factors, "null argument");
        ORACLE_REQUIRE(precision == 32 || precision == 64, "precision must be 32 or 64");
        ORACLE_REQUIRE(interfaces[0] == 0, "First interface must by at y=0 ");
        for (int l = 1; l < nlayers; l++) {
            ORACLE_REQUIRE(l == 1 || interfaces[l] > interfaces[l - 1], "interface coordinates must increase.");
            ORACLE_REQUIRE(dh_factors[l] == dh_factors[l - 1] * 3 || dh_factors[l] * 3 == dh_factors[l - 1] || dh_factors[l] == dh_factors[l - 1],
                           "Only gridspacing changes with factor 3 eg: 1<->3 or 9<->3 are alowed");
        }
        VarGrid vg;
        vg.init(desc->nx, desc->ny, desc->dim == 2 ? 1 : desc->nz, vector<int>(dh_factors, dh_factors + nlayers), vector<int>(interfaces + 1, interfaces + nlayers));
        if (fd_orders)
            vg.fdOrder.assign(fd_orders, fd_orders + nlayers);
        for (int o : vg.fdOrder)
            ORACLE_REQUIRE(o >= 2 && o <= WS_MAXQ_ORACLE && o % 2 == 0, "Unsupported spatialFDorder value.");
        auto h = std::make_unique<Handle>();
        h->precision = precision;
        if (precision == 64) {
            h->dd = std::make_unique<Oracle<double>>();
            h->dd->vg = vg;
            h->dd->create(*desc);
        } else {
            h->f = std::make_unique<Oracle<float>>();
            h->f->vg = vg;
            h->f->create(*desc);
        }
        *out = reinterpret_cast<wso_solver *>(h.release());
    });
}
int wso_grid_size(wso_solver *s, int32_t *nx, int32_t *ny, int32_t *nz, int32_t *n)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, { *nx = o.NX; *ny = o.NY; *nz = o.NZ; *n = o.N; }); });
}
int wso_coordinate2index(wso_solver *s, int32_t x, int32_t y, int32_t z, int32_t *index)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, *index = o.index(x, y, z)); });
}
void wso_destroy(wso_solver *s) { delete reinterpret_cast<Handle *>(s); }

int wso_set_material(wso_solver *s, const char *name, const float *host, size_t n)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, { ORACLE_REQUIRE(n == (size_t)o.N, "material size mismatch"); setVec(o.mat[name], host, n); o.prepared = false; }); });
}
int wso_get_material(wso_solver *s, const char *name, float *host, size_t n)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, getVec(o.M(name), host, n)); });
}
int wso_prepare(wso_solver *s)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, o.prepare()); });
}
int wso_set_sources(wso_solver *s, int32_t n, const int32_t *type, const int32_t *idx, const float *sig)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] {
        DISPATCH(h, {
            o.srcType.assign(type, type + n);
            o.srcIdx.assign(idx, idx + n);
            for (int32_t k = 0; k < n; k++)
                ORACLE_REQUIRE(idx[k] >= 0 && idx[k] < o.N, "source index out of range");
            setVec(o.srcSig, sig, (size_t)n * o.d.nt);
        });
    });
}
int wso_set_receivers(wso_solver *s, int32_t n, const int32_t *type, const int32_t *idx)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] {
        DISPATCH(h, {
            o.recType.assign(type, type + n);
            o.recIdx.assign(idx, idx + n);
            for (int32_t k = 0; k < n; k++)
                ORACLE_REQUIRE(idx[k] >= 0 && idx[k] < o.N, "receiver index out of range");
            o.seis.assign((size_t)n * o.d.nt, 0);
        });
    });
}
int wso_reset(wso_solver *s)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, o.reset()); });
}
/* 1 = fused matrix-free back-end (3-D elastic, regular grid, no ABS frame): the second CPU baseline of bench.py */
int wso_set_fused(wso_solver *s, int32_t on)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, o.fused = on != 0); });
}
int wso_step(wso_solver *s, int32_t t)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, o.step(t)); });
}
int wso_run(wso_solver *s, int32_t t0, int32_t t1)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] {
        for (int32_t t = t0; t < t1; t++)
            DISPATCH(h, o.step(t));
    });
}
int wso_get_seismogram(wso_solver *s, float *host)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, getVec(o.seis, host, o.seis.size())); });
}
int wso_get_wavefield(wso_solver *s, const char *comp, float *host, size_t n)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, getVec(o.F(comp), host, n)); });
}
int wso_set_wavefield(wso_solver *s, const char *comp, const float *host, size_t n)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    return guard([&] { DISPATCH(h, { ORACLE_REQUIRE(n == (size_t)o.N, "size mismatch"); setVec(o.F(comp), host, n); }); });
}

// Row `row` of a derivative matrix as dense taps (for checking the product's coefficient tables):
// which: 0 Dxf 1 Dxb 2 Dyf 3 Dyb 4 Dzf 5 Dzb 6 DyfFreeSurface 7 DybFreeSurface.  Returns nnz, cols/vals need >= 16 entries.
int wso_deriv_row(wso_solver *s, int which, int32_t row, int32_t *cols, float *vals)
{
    Handle *h = reinterpret_cast<Handle *>(s);
    int nnz = -1;
    int rc = guard([&] {
        DISPATCH(h, {
            ORACLE_REQUIRE(o.prepared, "call prepare first");
            auto *A = &o.Dxf;
            switch (which) {
            case 0: A = &o.Dxf; break;
            case 1: A = &o.Dxb; break;
            case 2: A = &o.Dyf; break;
            case 3: A = &o.Dyb; break;
            case 4: A = &o.Dzf; break;
            case 5: A = &o.Dzb; break;
            case 6: A = &o.DyfFS; break;
            case 7: A = &o.DybFS; break;
            default: throw std::runtime_error("bad matrix id");
            }
            ORACLE_REQUIRE(!A->empty(), "matrix not built for this configuration");
            ORACLE_REQUIRE(row >= 0 && row < A->n, "row out of range");
            nnz = (int)(A->ia[row + 1] - A->ia[row]);
            for (int k = 0; k < nnz; k++) {
                cols[k] = A->ja[A->ia[row] + k];
                vals[k] = (float)A->va[A->ia[row] + k];
            }
        });
    });
    return rc == WS_OK ? nnz : rc;
}

// Wavelets (Acquisition/SourceSignal/*.cpp), evaluated like the reference in ValueType = float.
// shape: 1 Ricker (Ricker.cpp:29-53), 7 Ricker_GprMax
int wso_wavelet(int shape, int32_t nt, float dt, float fc, float amp, float tshift, float *out)
{
    return guard([&] {
        ORACLE_REQUIRE(nt > 0 && dt > 0 && fc > 0, "NT, DT, FC must be positive");
        if (shape == 1) {
            float help = (float)(1.5 / fc + tshift);
            float w = (float)(M_PI * fc);
            for (int32_t k = 0; k < nt; k++) {
                float t = 0.0f + (float)k * dt;
                float tau = t - help;
                tau = tau * w;
                float h2 = tau * tau;
                float e = std::exp(-1.0f * h2);
                float hh = 1.0f - 2.0f * h2;
                out[k] = (amp * hh) * e;
            }
        } else
            throw std::runtime_error("Unknown wavelet shape ");
    });
}

int wso_num_threads(void)
{
    int n = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
#endif
    return n;
}

// bench.py only: number of OpenMP threads of the following runs (the launcher's OMP_NUM_THREADS must not decide it)
void wso_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// bench.py only: flush-to-zero / denormals-are-zero in every OpenMP thread.  The timed sample runs on a wavefield whose
// front is a shell of fp32 denormals; x86 handles those in microcode, which would make the CPU baseline depend on how far
// the front has travelled.  Never set by the parity tests (it changes roundings near zero).
void wso_set_flush_denormals(int on)
{
#if defined(__x86_64__) || defined(__i386__)
#pragma omp parallel
    {
        unsigned csr = __builtin_ia32_stmxcsr();
        if (on)
            csr |= 0x8040u; // FTZ | DAZ
        else
            csr &= ~0x8040u;
        __builtin_ia32_ldmxcsr(csr);
    }
#else
    (void)on;
#endif
}
}

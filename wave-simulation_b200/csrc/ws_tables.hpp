// ws_tables.hpp — host-side operator definition: what the reference expresses as LAMA matrices becomes a few small
// weight tables consumed by the matrix-free kernels.
//
// Reference: src/ForwardSolver/Derivatives/Derivatives.cpp (setFDCoef :2001-2042, calcDxf :112-186, calcDxb :706-763,
// calcDyfFreeSurface :367-440, calcDybFreeSurface :448-526), FDTD3D.cpp:183-323 (DT/DH scaling),
// BoundaryCondition/CPML.cpp:39-68, CPML3D.cpp:222-368, ABS3D.cpp:154-218.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace wstab {

// Taylor coefficients of the staggered first derivative, stored as float like the reference's ValueType
inline std::vector<float> fdCoef(int q)
{
    // exact rationals; float(double(p/q)) equals the reference's float constants for every order
    static const double num[6][12] = {
        {-1, 1},
        {1, -9, 9, -1},
        {-3, 25, -75, 75, -25, 3},
        {5, -49, 245, -1225, 1225, -245, 49, -5},
        {-35, 405, -567, 735, -19845, 19845, -735, 567, -405, 35},
        {63, -847, 5445, -22869, 12705, -160083, 160083, -12705, 22869, -5445, 847, -63}};
    static const double den[6][12] = {
        {1, 1},
        {24, 8, 8, 24},
        {640, 384, 64, 64, 384, 640},
        {7168, 5120, 3072, 1024, 1024, 3072, 5120, 7168},
        {294912, 229376, 40960, 8192, 16384, 16384, 8192, 40960, 229376, 294912},
        {2883584, 2359296, 1835008, 1310720, 131072, 131072, 131072, 131072, 1310720, 1835008, 2359296, 2883584}};
    if (q < 2 || q > 12 || (q & 1))
        throw std::invalid_argument("spatialFDorder = " + std::to_string(q) + " Unsupported spatialFDorder value.");
    const int k = q / 2 - 1;
    std::vector<float> c(q);
    for (int j = 0; j < q; j++)
        c[j] = (float)(num[k][j] / den[k][j]);
    return c;
}

// Weights of one matrix row as a dense window w[0..q] over offsets -h..+h (h = q/2) around `pos`.
//   forward : true = D?f (taps pos-h+1 .. pos+h), false = D?b (taps pos-h .. pos+h-1)
//   policy 0: StencilMatrix — taps outside [0,n) dropped, value c * (DT/DH)
//   policy 1: sparse assembly — order reduced symmetrically until the stencil fits, one-sided 2-point row at the far
//             end, value (c / DH) * DT
inline void rowWeights(bool forward, int policy, int q, int pos, int n, float DH, float DT, float *w)
{
    const int h = q / 2;
    for (int j = 0; j <= q; j++)
        w[j] = 0.0f;
    if (policy == 0) {
        const std::vector<float> c = fdCoef(q);
        const float s = DT / DH;
        for (int j = 0; j < q; j++) {
            const int off = forward ? j - h + 1 : j - h;
            const int X = pos + off;
            if (X >= 0 && X < n)
                w[off + h] = c[j] * s;
        }
        return;
    }
    int order = q;
    int c0 = pos; // row centre; shifted by one for the one-sided rows
    for (;;) {
        const int lo = forward ? c0 - order / 2 + 1 : c0 - order / 2;
        const int hi = forward ? c0 + order / 2 : c0 + order / 2 - 1;
        if (lo < 0) {
            order += 2 * lo;
            if (!forward && order == 0) {
                order = 2;
                c0 += 1;
            }
            continue;
        }
        if (hi >= n) {
            order -= 2 * (hi - n + 1);
            if (forward && order == 0) {
                order = 2;
                c0 -= 1;
            }
            continue;
        }
        break;
    }
    if (order < 2)
        throw std::invalid_argument("grid too small for the requested spatialFDorder");
    const std::vector<float> c = fdCoef(order);
    for (int j = 0; j < order; j++) {
        const int X = forward ? c0 + (j - order / 2 + 1) : c0 + (j - order / 2);
        const int off = X - pos;
        if (off < -h || off > h)
            throw std::logic_error("derivative tap outside the table window");
        float v = c[j] / DH;
        w[off + h] = v * DT;
    }
}

// Image-method rows (free surface at y = 0): coefficient of column Y is c_j - c_image, rows outside [0,n) dropped,
// value ((c_j - c_image) / DH) * DT; no order reduction at the bottom.
inline void rowWeightsFreeSurface(bool forward, int q, int pos, int n, float DH, float DT, float *w)
{
    const int h = q / 2;
    const std::vector<float> c = fdCoef(q);
    for (int j = 0; j <= q; j++)
        w[j] = 0.0f;
    for (int j = 0; j < q; j++) {
        const int off = forward ? j - h + 1 : j - h;
        const int Y = pos + off;
        float coeff = c[j];
        float image = 0.0f;
        if (forward) {
            if (q >= 2 + 2 * pos + j)
                image = c[q - 2 - 2 * pos - j];
        } else {
            if (q >= 1 + 2 * pos + j)
                image = c[q - 1 - 2 * pos - j];
        }
        if (Y >= 0 && Y < n) {
            float v = (coeff - image) / DH;
            w[off + h] = v * DT;
        }
    }
}

// representative coordinate of a row class (inverse of wsRowClass)
inline int classPos(int r, int n, int h) { return r < h ? r : (r == h ? h : n - h + (r - h - 1)); }

// tab[op][r][j], op in WS_NOPS order: xf xb yf yb zf zb yfFreeSurface ybFreeSurface
inline std::vector<float> buildTables(int q, int policy, bool freeSurface, int nx, int ny, int nz, int dim, float DH, float DT)
{
    const int h = q / 2, rows = 2 * h + 1, taps = q + 1;
    std::vector<float> tab((size_t)8 * rows * taps, 0.0f);
    auto need = [&](int n, const char *name) {
        if (n < taps)
            throw std::invalid_argument(std::string(name) + " must be >= spatialFDorder + 1");
    };
    need(nx, "NX");
    need(ny, "NY");
    if (dim == 3)
        need(nz, "NZ");
    const int axisN[3] = {nx, ny, nz};
    for (int axis = 0; axis < 3; axis++) {
        if (axis == 2 && dim != 3)
            continue;
        for (int dir = 0; dir < 2; dir++) {
            const int op = axis * 2 + dir;
            for (int r = 0; r < rows; r++)
                rowWeights(dir == 0, policy, q, classPos(r, axisN[axis], h), axisN[axis], DH, DT, &tab[((size_t)op * rows + r) * taps]);
        }
    }
    for (int dir = 0; dir < 2; dir++) {
        const int op = 6 + dir;
        for (int r = 0; r < rows; r++) {
            float *w = &tab[((size_t)op * rows + r) * taps];
            if (freeSurface)
                rowWeightsFreeSurface(dir == 0, q, classPos(r, ny, h), ny, DH, DT, w);
            else // getDyfFreeSurface is never consulted without a free surface; alias the plain operator
                rowWeights(dir == 0, policy, q, classPos(r, ny, h), ny, DH, DT, w);
        }
    }
    return tab;
}

// CPML.cpp:39-68 calcCoeffCPML (ValueType = float, double intermediates where the reference has double literals)
inline void calcCoeffCPML(std::vector<float> &a, std::vector<float> &b, float NPower, float fc, float vmax, float DT, float DH, bool shiftGrid)
{
    const int W = (int)a.size();
    const float shift = shiftGrid ? 0.5f : 0.0f;
    const float RCoef = 0.0008f;
    const float alpha_max = (float)(2.0 * M_PI * (fc / 2.0));
    const float d0 = (float)(-(NPower + 1) * vmax * std::log(RCoef) / (2.0 * W * DH));
    for (int i = 0; i < W; i++) {
        const float pos = (float)(W - i - shift) / W;
        const float d = d0 * (float)std::pow(pos, NPower);
        const float alpha_prime = (float)(alpha_max * (1.0 - pos));
        b[i] = (float)std::exp(-(d + alpha_prime) * DT);
        if (std::abs(d) > 1.0e-6)
            a[i] = (float)(d * (b[i] - 1.0) / (d + alpha_prime));
        else
            a[i] = 0.0f;
    }
}

// Per-axis CPML coefficient arrays of 2W entries: k < W low-coordinate side (distance k), k >= W high-coordinate
// side (distance 2W-1-k) where the full-grid and half-grid profiles are swapped (CPML3D.cpp:297-317).
struct CpmlAxis {
    std::vector<float> a, b, ah, bh;
};
inline CpmlAxis buildCpmlAxis(int W, float NPower, float fc, float vmax, float DT, float DH)
{
    std::vector<float> a(W), b(W), ah(W), bh(W);
    calcCoeffCPML(a, b, NPower, fc, vmax, DT, DH, false);
    calcCoeffCPML(ah, bh, NPower, fc, vmax, DT, DH, true);
    CpmlAxis c;
    c.a.resize(2 * W);
    c.b.resize(2 * W);
    c.ah.resize(2 * W);
    c.bh.resize(2 * W);
    for (int k = 0; k < W; k++) {
        c.a[k] = a[k];
        c.b[k] = b[k];
        c.ah[k] = ah[k];
        c.bh[k] = bh[k];
        const int dist = W - 1 - k; // k-th point of the high side has edge distance W-1-k
        c.a[W + k] = ah[dist];
        c.b[W + k] = bh[dist];
        c.ah[W + k] = a[dist];
        c.bh[W + k] = b[dist];
    }
    return c;
}

// ABS3D.cpp:174-179 / ABS2D.cpp: Cerjan damping profile
inline std::vector<float> buildAbsCoeff(int W, float dampingCoeff)
{
    std::vector<float> coeff(W);
    const float amp = (float)(1.0 - dampingCoeff / 100.0);
    const float a = (float)std::sqrt(-std::log(amp) / (float)(W * W));
    for (int j = 0; j < W; j++)
        coeff[j] = (float)std::exp(-(a * a * (W - j) * (W - j)));
    return coeff;
}

} // namespace wstab

#include "IrregularGrid.hpp"
#include <algorithm>
#include <cmath>
#include <map>

using namespace KITGPI;
using namespace KITGPI::ForwardSolver;

namespace
{
    // one row under construction: entries in assembly order, written out sorted by column (like lama::MatrixAssembly -> CSR)
    struct RowBuilder {
        std::vector<std::pair<int32_t, float>> e;
        void push(IndexType col, float v) { e.emplace_back((int32_t)col, v); }
    };
    void writeRows(EllRows &m, std::vector<RowBuilder> &rows)
    {
        m.taps = 1;
        for (auto const &r : rows)
            m.taps = std::max<IndexType>(m.taps, (IndexType)r.e.size());
        m.cols.assign(rows.size() * (size_t)m.taps, -1);
        m.vals.assign(rows.size() * (size_t)m.taps, 0.0f);
        for (size_t i = 0; i < rows.size(); i++) {
            std::stable_sort(rows[i].e.begin(), rows[i].e.end(), [](auto const &a, auto const &b) { return a.first < b.first; });
            for (size_t k = 0; k < rows[i].e.size(); k++) {
                m.cols[i * m.taps + k] = rows[i].e[k].first;
                m.vals[i * m.taps + k] = rows[i].e[k].second;
            }
        }
    }
}

// FD order of a row: the layer's order (useVariableFDoperators) and, on a grid with variable spacing, the reduction of the
// y operators towards the interfaces (Derivatives.cpp:237-244, 803-811)
template <typename ValueType> IndexType IrregularOperators<ValueType>::orderAt(IndexType layer, IndexType y, bool yOperator) const
{
    IndexType order = deriv.getSpatialFDorder(layer);
    if (yOperator && mc.hasVariableSpacing()) {
        const IndexType distance = mc.distToInterface(y) / mc.getDHFactor(layer);
        if (distance == 0)
            order = 2;
        else if (order > distance * 2)
            order = distance * 2;
    }
    return order;
}

template <typename ValueType> EllRows IrregularOperators<ValueType>::derivative(IndexType axis, bool forward, bool imageMethod) const
{
    const IndexType N = mc.getNGridpoints();
    const bool truncate = deriv.getEdgePolicy() == 0;
    std::map<IndexType, std::vector<ValueType>> fd;
    for (IndexType o = 2; o <= 12; o += 2)
        fd[o] = Derivatives::Derivatives<ValueType>::calcFDCoef(o);
    const IndexType nAxis = axis == 0 ? mc.getNX() : (axis == 1 ? mc.getNY() : mc.getNZ());
    std::vector<RowBuilder> rows((size_t)N);
    for (IndexType row = 0; row < N; row++) {
        Acquisition::coordinate3D c = mc.index2coordinate(row);
        const IndexType layer = mc.getLayer(c.y);
        const bool onInterface = mc.locatedOnInterface(c.y);
        const int trans = onInterface ? mc.getTransition(c.y) : 0;
        IndexType order = orderAt(layer, c.y, axis == 1);
        IndexType f = mc.getDHFactor(layer);
        ValueType dh = mc.getDH(layer);
        IndexType shift = 0;
        if (axis != 1 && onInterface) // interface planes are stored fine and operated coarse: taps move a third of the coarse spacing
            shift = forward ? -(f / 3) : f / 3;
        if (axis == 1 && forward && onInterface && trans == -1) { // coarse -> fine: Dyf of the interface works on the fine grid below
            f = mc.getDHFactor(layer + 1);
            dh = mc.getDH(layer + 1);
        }
        auto column = [&](IndexType pos) { return axis == 0 ? mc.coordinate2index(pos, c.y, c.z) : (axis == 1 ? mc.coordinate2index(c.x, pos, c.z) : mc.coordinate2index(c.x, c.y, pos)); };
        const ValueType dhBack = mc.getDH(layer); // Dyb divides by the layer's own spacing (Derivatives.cpp:839)
        if (imageMethod) {
            // Derivatives.cpp:367-440: coefficient minus the coefficient of the tap mirrored at the free surface
            for (IndexType j = 0; j < order; j++) {
                const IndexType Y = c.y + f * (j - order / 2 + 1);
                ValueType coeff = fd[order][j];
                if (order >= 2 + 2 * c.y / f + j)
                    coeff -= fd[order][order - 2 - 2 * c.y / f - j];
                if (Y >= 0 && Y < nAxis)
                    rows[row].push(column(Y), (float)((coeff / dh) * DT));
            }
            continue;
        }
        IndexType pos = axis == 0 ? c.x : (axis == 1 ? c.y : c.z);
        for (IndexType j = 0; j < order; j++) {
            const IndexType first = forward ? -order / 2 + 1 : -order / 2;
            IndexType X = pos + shift + f * (j + first);
            const IndexType Xmin = pos + shift + f * first;
            const IndexType Xmax = forward ? pos + shift + f * order / 2 : pos + shift + f * (order / 2 - 1);
            if (axis == 1 && !forward && onInterface) { // staggered positions next to the interface (Derivatives.cpp:823-830)
                if (j == 0 && trans == 1)
                    X += mc.getDHFactor(layer - 1);
                if (j == 1 && trans == -1)
                    X += mc.getDHFactor(layer + 1);
            }
            const ValueType value = (ValueType)((fd[order][j] / (axis == 1 && !forward ? dhBack : dh)) * DT);
            if (truncate) { // full order, taps outside the grid dropped (LAMA StencilMatrix behaviour; what the reference's goldens contain)
                if (X >= 0 && X < nAxis)
                    rows[row].push(column(X), value);
                continue;
            }
            // order reduction towards the grid edges (Derivatives.cpp:159-175, 263-276, 744-757, 832-843)
            if (Xmin < 0) {
                order += 2 * Xmin;
                if (!forward && order == 0) {
                    order = 2;
                    pos += 1;
                }
                j--;
            } else if (Xmax >= nAxis) {
                order -= 2 * (Xmax - nAxis + 1);
                if (forward && order == 0) {
                    order = 2;
                    pos -= 1;
                }
                j--;
            } else
                rows[row].push(column(X), value);
        }
    }
    EllRows m;
    writeRows(m, rows);
    return m;
}

template <typename ValueType> EllRows IrregularOperators<ValueType>::interpolation(IndexType mode) const
{
    const IndexType N = mc.getNGridpoints(), NX = mc.getNX(), NZ = mc.getNZ();
    EllRows m;
    std::vector<RowBuilder> rows;
    for (IndexType row = 0; row < N; row++) {
        const Acquisition::coordinate3D c = mc.index2coordinate(row);
        if (!mc.locatedOnInterface(c.y))
            continue; // identity row
        const IndexType layer = mc.getLayer(c.y), f = mc.getDHFactor(layer);
        const int trans = mc.getTransition(c.y);
        const IndexType fFine = trans == 1 ? mc.getDHFactor(layer - 1) : (trans == -1 ? mc.getDHFactor(layer + 1) : f);
        const ValueType denom = ValueType(1) / ValueType(f * f);
        // bilinear weights between the four surrounding points of the coarse (or coarse staggered) grid
        const IndexType modx = mode == 1 ? (c.x - f / 2) % f : c.x % f, modz = mode == 2 ? (c.z - f / 2) % f : c.z % f;
        const IndexType reachX = mode == 1 ? 1 : 2, reachZ = mode == 2 ? 1 : 2;
        const bool lowX = mode != 1 || c.x >= modx, lowZ = mode != 2 || c.z >= modz;
        const bool highX = c.x + reachX * fFine < NX, highZ = c.z + reachZ * fFine < NZ;
        RowBuilder r;
        if (lowX && lowZ)
            r.push(mc.coordinate2index(c.x - modx, c.y, c.z - modz), (float)((f - modx) * (f - modz) * denom));
        if (highX)
            r.push(mc.coordinate2index(c.x + f - modx, c.y, c.z - modz), (float)(modx * (f - modz) * denom));
        if (lowX && highZ)
            r.push(mc.coordinate2index(c.x - modx, c.y, c.z + f - modz), (float)((f - modx) * modz * denom));
        if (highX && highZ)
            r.push(mc.coordinate2index(c.x + f - modx, c.y, c.z + f - modz), (float)(modx * modz * denom));
        m.rows.push_back((int32_t)row);
        rows.push_back(r);
    }
    writeRows(m, rows);
    return m;
}

template <typename ValueType> std::vector<ValueType> IrregularOperators<ValueType>::inverseAverage(std::vector<ValueType> const &par, IndexType axis) const
{
    const IndexType N = mc.getNGridpoints();
    SCAI_ASSERT_ERROR((IndexType)par.size() == N, "model vector must hold one value per grid point")
    std::vector<ValueType> out((size_t)N);
    for (IndexType row = 0; row < N; row++) {
        const Acquisition::coordinate3D c = mc.index2coordinate(row);
        const IndexType layer = mc.getLayer(c.y);
        IndexType X = c.x, Y = c.y, Z = c.z;
        bool inside;
        if (axis == 0) {
            X += mc.getDHFactor(layer);
            inside = X < mc.getNX();
        } else if (axis == 2) {
            Z += mc.getDHFactor(layer);
            inside = Z < mc.getNZ();
        } else {
            Y += (mc.locatedOnInterface(c.y) && mc.getTransition(c.y) == 0) ? mc.getDHFactor(layer + 1) : mc.getDHFactor(layer);
            inside = Y < mc.getNY();
        }
        // row of the averaging matrix times the vector, ascending column order (the second point always has the larger index)
        ValueType s = 0;
        if (inside) {
            const IndexType other = mc.coordinate2index(X, Y, Z);
            const ValueType w = ValueType(1.0 / 2.0);
            if (other > row) {
                s = s + w * par[row];
                s = s + w * par[other];
            } else {
                s = s + w * par[other];
                s = s + w * par[row];
            }
        } else
            s = s + ValueType(1.0) * par[row];
        ValueType r = ValueType(1) / s;
        if (std::isnan(r) || std::isinf(r))
            r = 0;
        out[row] = r;
    }
    return out;
}

template <typename ValueType>
CpmlProfile IrregularOperators<ValueType>::cpml(IndexType axis, IndexType boundaryWidth, ValueType NPower, ValueType centerFrequency, ValueType vMax, bool freeSurface) const
{
    // per layer a profile of ceil(BoundaryWidth / dhFactor) points with the layer's spacing (CPML.cpp:39-68 calcCoeffCPML)
    const IndexType numLayers = mc.getNumLayers();
    std::vector<std::vector<ValueType>> a(numLayers), b(numLayers), ah(numLayers), bh(numLayers);
    for (IndexType l = 0; l < numLayers; l++) {
        const IndexType width = (IndexType)std::ceil((float)boundaryWidth / mc.getDHFactor(l));
        for (int half = 0; half < 2; half++) {
            std::vector<ValueType> &A = half ? ah[l] : a[l], &B = half ? bh[l] : b[l];
            A.assign(width, 0);
            B.assign(width, 0);
            const ValueType shift = half ? 0.5 : 0;
            const ValueType RCoef = 0.0008;
            const ValueType alphaMax = 2.0 * M_PI * (centerFrequency / 2.0);
            const ValueType d0 = -(NPower + 1) * vMax * std::log(RCoef) / (2.0 * width * mc.getDH(l));
            for (IndexType i = 0; i < width; i++) {
                const ValueType positionNorm = (ValueType)(width - i - shift) / width;
                const ValueType d = d0 * std::pow(positionNorm, NPower);
                const ValueType alphaPrime = alphaMax * (1.0 - positionNorm);
                B[i] = std::exp(-(d + alphaPrime) * DT);
                A[i] = std::abs(d) > 1.0e-6 ? d * (B[i] - 1.0) / (d + alphaPrime) : 0.0;
            }
        }
    }
    CpmlProfile p;
    const IndexType N = mc.getNGridpoints();
    for (IndexType row = 0; row < N; row++) {
        const Acquisition::coordinate3D c = mc.index2coordinate(row);
        const IndexType l = mc.getLayer(c.y), f = mc.getDHFactor(l);
        const IndexType width = (IndexType)std::ceil((float)boundaryWidth / f);
        const Acquisition::coordinate3D g = mc.edgeDistance(c);
        const IndexType dist = (axis == 0 ? g.x : (axis == 1 ? g.y : g.z)) / f;
        const IndexType coord = (axis == 0 ? c.x : (axis == 1 ? c.y : c.z)) / f;
        if (dist >= width)
            continue;
        const bool low = coord < width;
        if (axis == 1 && low && freeSurface)
            continue; // no CPML below a free surface
        p.idx.push_back((int32_t)row);
        // on the high-coordinate side the full- and half-grid profiles swap (CPML2DAcoustic.cpp:171-181)
        p.a.push_back((float)(low ? a[l][dist] : ah[l][dist]));
        p.b.push_back((float)(low ? b[l][dist] : bh[l][dist]));
        p.aHalf.push_back((float)(low ? ah[l][dist] : a[l][dist]));
        p.bHalf.push_back((float)(low ? bh[l][dist] : b[l][dist]));
    }
    return p;
}

template <typename ValueType> AbsProfile IrregularOperators<ValueType>::abs(IndexType boundaryWidth, ValueType dampingCoeff, IndexType useFreeSurface, bool threeD) const
{
    const IndexType W = boundaryWidth;
    std::vector<float> coeff(W);
    const float amp = (float)(1.0 - dampingCoeff / 100.0);
    const float a = (float)std::sqrt(-std::log(amp) / (float)(W * W));
    for (IndexType j = 0; j < W; j++)
        coeff[j] = (float)std::exp(-(a * a * (W - j) * (W - j)));
    AbsProfile p;
    for (IndexType row = 0; row < mc.getNGridpoints(); row++) {
        const Acquisition::coordinate3D c = mc.index2coordinate(row), g = mc.edgeDistance(c);
        IndexType mn = g.x < g.y ? g.x : g.y;
        if (threeD && g.z < mn)
            mn = g.z;
        IndexType k = -1;
        if (useFreeSurface == 0) {
            if (mn < W)
                k = mn;
        } else if (c.y < W) { // below a free surface only the side frames damp
            const IndexType side = threeD ? (!(g.x < g.z) ? g.z : g.x) : g.x;
            if (side < W)
                k = side;
        } else if (mn < W)
            k = mn;
        if (k >= 0) {
            p.idx.push_back((int32_t)row);
            p.damping.push_back(coeff[k]);
        }
    }
    return p;
}

template <typename ValueType> std::vector<int32_t> IrregularOperators<ValueType>::surfacePoints() const
{
    std::vector<int32_t> s;
    for (IndexType row = 0; row < mc.getNGridpoints(); row++)
        if (mc.locatedOnSurface(row))
            s.push_back((int32_t)row);
    return s;
}

template class KITGPI::ForwardSolver::IrregularOperators<float>;

#include "Derivatives.hpp"
#include "Coordinates.hpp"
#include <algorithm>

using namespace KITGPI;
using namespace KITGPI::ForwardSolver::Derivatives;

template <typename ValueType> std::vector<ValueType> Derivatives<ValueType>::calcFDCoef(IndexType spFDo)
{
    // c_k of D f(i) = sum_k c_k (f(i+k) - f(i-k+1)) / DH, stored as the 2*(order/2) taps in ascending offset
    std::vector<double> half;
    switch (spFDo) {
    case 2: half = {1.0}; break;
    case 4: half = {9.0 / 8.0, -1.0 / 24.0}; break;
    case 6: half = {75.0 / 64.0, -25.0 / 384.0, 3.0 / 640.0}; break;
    case 8: half = {1225.0 / 1024.0, -245.0 / 3072.0, 49.0 / 5120.0, -5.0 / 7168.0}; break;
    case 10: half = {19845.0 / 16384.0, -735.0 / 8192.0, 567.0 / 40960.0, -405.0 / 229376.0, 35.0 / 294912.0}; break;
    case 12: half = {160083.0 / 131072.0, -12705.0 / 131072.0, 22869.0 / 1310720.0, -5445.0 / 1835008.0, 847.0 / 2359296.0, -63.0 / 2883584.0}; break;
    default: COMMON_THROWEXCEPTION("spatialFDorder = " << spFDo << " Unsupported spatialFDorder value.")
    }
    const size_t h = half.size();
    std::vector<ValueType> c(2 * h);
    for (size_t k = 0; k < h; k++) {
        c[h + k] = (ValueType)half[k];
        c[h - 1 - k] = (ValueType)(-half[k]);
    }
    return c;
}

template <typename ValueType> void Derivatives<ValueType>::init(Configuration::Configuration const &config)
{
    spatialFDorder = config.get<IndexType>("spatialFDorder");
    FDCoef = calcFDCoef(spatialFDorder);
    useStencilMatrix = config.getAndCatch("useStencilMatrix", 0) != 0;
    // Derivatives.cpp:35-44: the hybrid free-surface operator (stencil + sparse image corrections) is the image-method operator the
    // stencil kernels apply in one pass; the key is accepted with the reference's precondition
    if (config.getAndCatch("useHybridFreeSurface", 0) != 0)
        SCAI_ASSERT_ERROR(useStencilMatrix, "It is not possible to use the hybrid matrix without stencil matrix!")
    spatialFDorderVec.clear();
    if (config.getAndCatch("useVariableFDoperators", 0) != 0) { // Derivatives.cpp:47-53
        SCAI_ASSERT_ERROR(!useStencilMatrix, "Variable FD operators are not available for stencil matrices")
        spatialFDorderVec = Acquisition::readColumnFromFile(config.get<std::string>("gridConfigurationFilename"), 2);
        for (IndexType o : spatialFDorderVec)
            calcFDCoef(o); // throws "Unsupported spatialFDorder value."
    }
    if (config.getAndCatch("useVariableGrid", 0) != 0)
        SCAI_ASSERT_ERROR(!useStencilMatrix, "It is not possible to use the stencil matrix on a variable grid")
    edgePolicy = config.getAndCatch("edgePolicy", useStencilMatrix ? 0 : 1);
    SCAI_ASSERT_ERROR(edgePolicy == 0 || edgePolicy == 1, "edgePolicy must be 0 or 1")
    useFreeSurface = config.get<IndexType>("FreeSurface");
    SCAI_ASSERT_ERROR(useFreeSurface >= 0 && useFreeSurface <= 2, "FreeSurface must be 0, 1 (image method) or 2 (improved vacuum formulation)")
}

template <typename ValueType> typename Derivatives<ValueType>::DerivativesPtr Factory<ValueType>::Create(std::string dimension)
{
    std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
    if (dimension == "2d")
        return std::make_shared<Derivatives<ValueType>>(2);
    if (dimension == "3d")
        return std::make_shared<Derivatives<ValueType>>(3);
    COMMON_THROWEXCEPTION("Unkown dimension")
}

template class KITGPI::ForwardSolver::Derivatives::Derivatives<float>;
template class KITGPI::ForwardSolver::Derivatives::Factory<float>;

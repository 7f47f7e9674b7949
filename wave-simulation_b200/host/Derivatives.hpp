// Derivatives.hpp — FD operator definition (mirror of src/ForwardSolver/Derivatives/Derivatives.hpp:115-176).
// In the reference this class owns the D?f / D?b sparse or stencil matrices.  Here the operators are matrix-free
// kernels, so the class is a descriptor: spatial FD order, Taylor coefficients (Derivatives.cpp:2001-2042), the edge
// policy implied by useStencilMatrix (Derivatives.cpp:112-121 truncation vs :129-186 order reduction) and the
// free-surface flag; the variable-grid / variable-order operators (gridConfig.txt) are rejected.
#pragma once
#include "Common.hpp"
#include "Configuration.hpp"
#include <memory>

namespace KITGPI
{
    namespace ForwardSolver
    {
        namespace Derivatives
        {
            template <typename ValueType> class Derivatives
            {
              public:
                typedef std::shared_ptr<Derivatives<ValueType>> DerivativesPtr;
                explicit Derivatives(IndexType dim) : numDimension(dim) {}

                //! = setup(config) + init(dist, ctx, modelCoordinates, comm) of the reference (FDTD3D.cpp:51-61)
                void init(Configuration::Configuration const &config);
                IndexType getSpatialFDorder() const { return spatialFDorder; }
                //! FD order of a layer of the grid (useVariableFDoperators: third column of gridConfigurationFilename, Derivatives.cpp:1663-1673)
                IndexType getSpatialFDorder(IndexType layer) const { return spatialFDorderVec.empty() ? spatialFDorder : spatialFDorderVec.at(layer); }
                bool getUseVarFDorder() const { return !spatialFDorderVec.empty(); }
                bool getUseStencilMatrix() const { return useStencilMatrix; }
                bool getUseFreeSurface() const { return useFreeSurface == 1; }
                //! 0 = off-grid taps dropped (StencilMatrix), 1 = order reduced towards the edges (sparse assembly)
                //! 0: off-grid taps dropped (StencilMatrix), 1: order reduced towards the grid edges (sparse assembly, Derivatives.cpp:159-175).
                //! Follows useStencilMatrix; the key edgePolicy (B200 extension) overrides it, e.g. to run a variable grid — where the
                //! reference has sparse matrices only — with the dropped taps its golden seismograms were produced with.
                IndexType getEdgePolicy() const { return edgePolicy; }
                IndexType getNumDimension() const { return numDimension; }
                //! Taylor coefficients of the staggered first derivative, spatialFDorder entries (setFDCoef)
                std::vector<ValueType> const &getFDCoef() const { return FDCoef; }
                static std::vector<ValueType> calcFDCoef(IndexType spFDo);
                //! sparse: N*q*(4+4)+2N*4 B per matrix, stencil: N*4 B (Derivatives.cpp:1704-1725); matrix-free: 0
                ValueType estimateMemory() const { return 0; }

              private:
                IndexType numDimension;
                IndexType spatialFDorder = 0, useFreeSurface = 0, edgePolicy = 1;
                std::vector<IndexType> spatialFDorderVec;
                bool useStencilMatrix = false;
                std::vector<ValueType> FDCoef;
            };

            template <typename ValueType> class Factory
            {
              public:
                static typename Derivatives<ValueType>::DerivativesPtr Create(std::string dimension); // "2D" | "3D" (DerivativesFactory.cpp:5-23)
            };
        }
    }
}

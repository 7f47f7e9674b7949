// IrregularGrid.hpp — operators of a layered (variable) grid, assembled point by point by the reference's rules and handed to
// the CUDA library in ELL form (include/wavesim.h, operator-given mode).  Mirrors, for the forward path of the acoustic
// solvers: Derivatives::calcDxf / calcDxb / calcDyf / calcDyb / calcDzf / calcDzb / calcDyfFreeSurface (Derivatives.cpp:129-843,
// 1177-1243), calcInterpolationFull / StaggeredX / StaggeredZ (:1252-1566), Modelparameter::calcAverageMatrixX/Y/Z
// (Modelparameter.cpp:336-437) and CPML{2D,3D}Acoustic::init (CPML2DAcoustic.cpp:99-200).
#pragma once
#include "Common.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include <cstdint>

namespace KITGPI
{
    namespace ForwardSolver
    {
        //! n rows x taps entries, row-major, columns ascending, unused entries cols = -1
        struct EllRows {
            IndexType taps = 0;
            std::vector<int32_t> rows; // empty = every row of the model vector
            std::vector<int32_t> cols;
            std::vector<float> vals;
        };
        struct CpmlProfile {
            std::vector<int32_t> idx;
            std::vector<float> a, b, aHalf, bHalf;
        };

        struct AbsProfile {
            std::vector<int32_t> idx;
            std::vector<float> damping;
        };

        template <typename ValueType> class IrregularOperators
        {
          public:
            IrregularOperators(Acquisition::Coordinates<ValueType> const &coordinates, Derivatives::Derivatives<ValueType> const &derivatives, ValueType DT)
                : mc(coordinates), deriv(derivatives), DT(DT)
            {
            }
            //! axis 0 x, 1 y, 2 z; values carry DT / DH(layer); imageMethod: DyfFreeSurface (forward y only)
            EllRows derivative(IndexType axis, bool forward, bool imageMethod = false) const;
            //! mode 0 full grid points, 1 staggered in x, 2 staggered in z; only the rows of the interface planes
            EllRows interpolation(IndexType mode) const;
            //! 1 / (average of `par` over the two points of the staggered position), Inf / NaN -> 0 (Modelparameter.cpp:633-639)
            std::vector<ValueType> inverseAverage(std::vector<ValueType> const &par, IndexType axis) const;
            CpmlProfile cpml(IndexType axis, IndexType boundaryWidth, ValueType NPower, ValueType centerFrequency, ValueType vMax, bool freeSurface) const;
            //! the sparse vector `damping` of ABS2D::init / ABS3D::init (ABS2D.cpp:110-178, ABS3D.cpp:154-218): the Cerjan coefficient of the distance
            //! to the nearest edge, measured on the coordinates of the points (units of the finest spacing); useFreeSurface != 0: no frame at the top
            AbsProfile abs(IndexType boundaryWidth, ValueType dampingCoeff, IndexType useFreeSurface, bool threeD) const;
            std::vector<int32_t> surfacePoints() const;

          private:
            IndexType orderAt(IndexType layer, IndexType y, bool yOperator) const;
            Acquisition::Coordinates<ValueType> const &mc;
            Derivatives::Derivatives<ValueType> const &deriv;
            ValueType DT;
        };
    }
}

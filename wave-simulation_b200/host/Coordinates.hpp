// Coordinates.hpp — linear index <-> (x, y, z) on the regular grid (mirror of src/Acquisition/Coordinates.hpp:58-154).
// index = x + z*NX + y*NX*NZ: x fastest, y = depth is the slowest axis (Coordinates.cpp:615-694).  The variable grid
// (dhFactor 3^n layers, Coordinates.cpp:115-247) is not part of the B200 path: useVariableGrid=1 throws.
#pragma once
#include "Common.hpp"
#include "Configuration.hpp"

namespace KITGPI
{
    namespace Acquisition
    {
        struct coordinate3D {
            IndexType x, y, z;
            IndexType min() const { return std::min(x, std::min(y, z)); }
        };

        template <typename ValueType> class Coordinates
        {
          public:
            Coordinates() : NX(0), NY(0), NZ(0), DH(0) {}
            Coordinates(IndexType nx, IndexType ny, IndexType nz, ValueType dh) { init(nx, ny, nz, dh); }
            explicit Coordinates(Configuration::Configuration const &config) { init(config); }

            void init(Configuration::Configuration const &config);
            void init(IndexType nx, IndexType ny, IndexType nz, ValueType dh);

            IndexType getNX() const { return NX; }
            IndexType getNY() const { return NY; }
            IndexType getNZ() const { return NZ; }
            ValueType getDH() const { return DH; }
            IndexType getNGridpoints() const { return NX * NY * NZ; }
            bool isVariable() const { return false; }

            coordinate3D index2coordinate(IndexType index) const;
            IndexType coordinate2index(coordinate3D coordinate) const { return coordinate2index(coordinate.x, coordinate.y, coordinate.z); }
            IndexType coordinate2index(IndexType X, IndexType Y, IndexType Z) const;
            coordinate3D edgeDistance(coordinate3D coordinate) const;
            bool locatedOnSurface(IndexType index) const { return index2coordinate(index).y == 0; }

          private:
            void check(IndexType X, IndexType Y, IndexType Z) const;
            IndexType NX, NY, NZ;
            ValueType DH;
        };
    }
}

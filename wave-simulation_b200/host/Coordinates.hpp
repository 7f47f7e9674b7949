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
            //! variable grid / variable FD order: layers of spacing dhFactor * DH below the interfaces (Coordinates.cpp:115-247);
            //! `interfaces` without the leading 0 of the gridConfig file
            void init(IndexType nx, IndexType ny, IndexType nz, ValueType dh, std::vector<IndexType> const &dhFactors, std::vector<IndexType> const &interfaces);

            std::vector<IndexType> const &getInterfaceVec() const { return interface; } // [-1, interfaces as moved to fit the grid, NY - 1]
            IndexType getNX() const { return NX; }
            IndexType getNY() const { return NY; }
            IndexType getNZ() const { return NZ; }
            ValueType getDH() const { return DH; }
            ValueType getX0() const { return x0; } // origin of the model in metres (key x0, Coordinates.cpp:40): places a sub-model in the big one
            IndexType getNGridpoints() const { return layered ? nGridpoints : NX * NY * NZ; }
            //! <filename>X / Y / Z: the grid coordinates of every point as three vectors (Coordinates.cpp:545-590; key writeCoordinate)
            void writeCoordinates(std::string const &filename, IndexType fileFormat) const;
            //! true when the model vector is layered (useVariableGrid or useVariableFDoperators): the operators are assembled point by point
            bool isVariable() const { return layered; }
            bool hasVariableSpacing() const { return variableSpacing; }
            IndexType getNGridpoints(IndexType layer) const { return layered ? nGridpointsPerLayer.at(layer) : NX * NY * NZ; }
            IndexType getNumLayers() const { return layered ? (IndexType)dhFactor.size() : 1; }
            IndexType getLayer(IndexType y) const;                    // Coordinates.cpp:383-398
            IndexType getDHFactor(IndexType layer) const { return layered ? dhFactor[layer] : 1; }
            ValueType getDH(IndexType layer) const { return DH * getDHFactor(layer); }
            bool locatedOnInterface(IndexType y) const;               // :474-483
            IndexType distToInterface(IndexType y) const;             // :490-499
            int getTransition(IndexType y) const;                     // :517-530

            coordinate3D index2coordinate(IndexType index) const;
            IndexType coordinate2index(coordinate3D coordinate) const { return coordinate2index(coordinate.x, coordinate.y, coordinate.z); }
            IndexType coordinate2index(IndexType X, IndexType Y, IndexType Z) const;
            coordinate3D edgeDistance(coordinate3D coordinate) const;
            bool locatedOnSurface(IndexType index) const { return index2coordinate(index).y == 0; }

          private:
            void check(IndexType X, IndexType Y, IndexType Z) const;
            IndexType NX, NY, NZ;
            ValueType DH;
            ValueType x0 = 0;
            bool layered = false, variableSpacing = false;
            IndexType nGridpoints = 0;
            std::vector<IndexType> dhFactor, interface, transition, layerStart, layerEnd, varNX, varNY, varNZ, nGridpointsPerLayer;
        };
        //! one column of a whitespace-separated text file, '#' comments skipped (Common::readColumnFromFile)
        std::vector<IndexType> readColumnFromFile(std::string const &filename, unsigned column);
    }
}

#include "Coordinates.hpp"
#include "IO.hpp"
#include <fstream>
#include <sstream>

using namespace KITGPI;

std::vector<IndexType> Acquisition::readColumnFromFile(std::string const &filename, unsigned column)
{
    std::ifstream in(filename);
    SCAI_ASSERT_ERROR(in.good(), "Could not open " << filename)
    std::vector<IndexType> out;
    std::string line;
    while (std::getline(in, line)) {
        const size_t first = line.find_first_not_of(" \t\r");
        if (first == std::string::npos || line[first] == '#')
            continue;
        std::istringstream ss(line);
        std::string tok;
        for (unsigned c = 0; c <= column; c++)
            SCAI_ASSERT_ERROR(static_cast<bool>(ss >> tok), filename << ": line '" << line << "' has no column " << column)
        out.push_back((IndexType)std::stol(tok));
    }
    return out;
}

// Coordinates.cpp:32-90: regular grid, variable grid (useVariableGrid) or layers of one spacing with their own FD order
// (useVariableFDoperators); the first interface of gridConfigurationFilename must be 0
template <typename ValueType> void Acquisition::Coordinates<ValueType>::init(Configuration::Configuration const &config)
{
    const IndexType nx = config.get<IndexType>("NX"), ny = config.get<IndexType>("NY"), nz = config.get<IndexType>("NZ");
    const ValueType dh = config.get<ValueType>("DH");
    x0 = config.getAndCatch("x0", ValueType(0));
    const bool varGrid = config.getAndCatch("useVariableGrid", 0) != 0, varFD = config.getAndCatch("useVariableFDoperators", 0) != 0;
    if (!varGrid && !varFD) {
        init(nx, ny, nz, dh);
        return;
    }
    const std::string file = config.get<std::string>("gridConfigurationFilename");
    std::vector<IndexType> ifc = readColumnFromFile(file, 0);
    SCAI_ASSERT_ERROR(!ifc.empty() && ifc[0] == 0, "First interface must by at y=0 ")
    ifc.erase(ifc.begin());
    std::vector<IndexType> factors(ifc.size() + 1, 1);
    if (varGrid) {
        factors = readColumnFromFile(file, 1);
        for (size_t i = 1; i < ifc.size(); i++)
            SCAI_ASSERT_ERROR(ifc[i] > ifc[i - 1], "interface coordinates must increase. Interface " << i << " value: " << ifc[i] << " is smaller than  Interface " << i - 1 << " value: " << ifc[i - 1])
        for (size_t i = 1; i < factors.size(); i++)
            if (factors[i] != factors[i - 1] * 3 && factors[i] != factors[i - 1] / 3 && factors[i] != factors[i - 1])
                COMMON_THROWEXCEPTION("Only gridspacing changes with factor 3 eg: 1<->3 or 9<->3 are alowed")
    }
    init(nx, ny, nz, dh, factors, ifc);
}

template <typename ValueType>
void Acquisition::Coordinates<ValueType>::init(IndexType nx, IndexType ny, IndexType nz, ValueType dh, std::vector<IndexType> const &dhFactors, std::vector<IndexType> const &interfaces)
{
    init(nx, ny, nz, dh);
    dhFactor = dhFactors;
    const IndexType numLayers = (IndexType)dhFactor.size();
    SCAI_ASSERT_ERROR(numLayers > 0, "vector of different grid spacings: dhFactor is emty")
    SCAI_ASSERT_ERROR((IndexType)interfaces.size() == numLayers - 1, "number of interfaces doesn't match to the number of different grid spacings")
    interface = interfaces;
    interface.push_back(NY - 1);
    interface.insert(interface.begin(), -1);
    IndexType dhMax = 0;
    for (IndexType f : dhFactor) {
        IndexType r = f;
        while (r > 1 && r % 3 == 0)
            r /= 3;
        SCAI_ASSERT_ERROR(f >= 1 && r == 1, "incompatible dhFactor, dhFactor must be 3^n")
        dhMax = std::max(dhMax, f);
    }
    variableSpacing = dhMax > 1;
    IndexType NXmax = NX, NZmax = NZ;
    if (dhMax != 1) {
        // NX, NZ shrink until the coarsest grid ends on a grid point; interfaces move up until every layer is a whole number of its cells
        while (NXmax != (NXmax / dhMax) * dhMax + 1 + dhMax / 2)
            NXmax--;
        while (NZ != 1 && NZmax != (NZmax / dhMax) * dhMax + 1 + dhMax / 2)
            NZmax--;
        NX = NXmax;
        NZ = NZmax;
        IndexType layer = 0;
        while (layer < 1) {
            layer++;
            if ((interface[layer] - interface[layer - 1] - 1) % dhFactor[layer - 1] != 0) {
                interface[layer]--;
                layer--;
            }
        }
        layer = 1;
        while (layer < numLayers) {
            layer++;
            if ((interface[layer] - interface[layer - 1]) % dhFactor[layer - 1] != 0) {
                interface[layer]--;
                layer--;
            }
        }
    }
    NY = interface[numLayers] + 1;
    transition.assign(numLayers, 0);
    layerStart.assign(numLayers, 0);
    layerEnd.assign(numLayers, 0);
    // an interface plane is stored with the spacing of the finer of its two layers
    for (IndexType l = 0; l + 1 < numLayers; l++) {
        if (dhFactor[l] < dhFactor[l + 1]) {
            transition[l] = 1;
            layerEnd[l] = interface[l + 1];
            layerStart[l + 1] = interface[l + 1] + dhFactor[l + 1];
        } else {
            transition[l] = dhFactor[l] > dhFactor[l + 1] ? -1 : 0;
            layerEnd[l] = interface[l + 1] - dhFactor[l];
            layerStart[l + 1] = interface[l + 1];
        }
    }
    layerEnd[numLayers - 1] = interface[numLayers];
    varNX.assign(numLayers, 0);
    varNY.assign(numLayers, 0);
    varNZ.assign(numLayers, 0);
    nGridpointsPerLayer.assign(numLayers, 0);
    nGridpoints = 0;
    for (IndexType l = 0; l < numLayers; l++) {
        varNY[l] = (layerEnd[l] - layerStart[l]) / dhFactor[l] + 1;
        varNX[l] = NXmax / dhFactor[l] + (dhFactor[l] > 1 ? 1 : 0);
        varNZ[l] = NZmax / dhFactor[l] + (dhFactor[l] > 1 ? 1 : 0);
        nGridpointsPerLayer[l] = varNX[l] * varNY[l] * varNZ[l];
        nGridpoints += nGridpointsPerLayer[l];
    }
    layered = true;
}

template <typename ValueType> IndexType Acquisition::Coordinates<ValueType>::getLayer(IndexType y) const
{
    if (!layered)
        return 0;
    const IndexType numLayers = (IndexType)dhFactor.size();
    IndexType layer = 0;
    for (layer = 0; layer < numLayers; layer++) {
        if (y < interface[layer + 1] && y > interface[layer])
            break;
        if (y == interface[layer + 1]) {
            if (transition[layer] > 0) // the coarse grid owns the interface of a fine -> coarse transition
                layer++;
            break;
        }
    }
    return layer;
}

template <typename ValueType> bool Acquisition::Coordinates<ValueType>::locatedOnInterface(IndexType y) const
{
    if (!layered)
        return false;
    for (size_t l = 0; l + 1 < interface.size(); l++)
        if (y == interface[l])
            return true;
    return false;
}

template <typename ValueType> IndexType Acquisition::Coordinates<ValueType>::distToInterface(IndexType y) const
{
    IndexType dist = NY;
    for (size_t k = 1; k + 1 < interface.size(); k++)
        dist = std::min(dist, (IndexType)std::abs((int)y - (int)interface[k]));
    return dist;
}

template <typename ValueType> int Acquisition::Coordinates<ValueType>::getTransition(IndexType y) const
{
    SCAI_ASSERT_ERROR(locatedOnInterface(y), "Y Coordinate Y=" << y << " is not located on an variable grid interface")
    int t = 0;
    for (size_t l = 0; l + 1 < interface.size(); l++)
        if (y == interface[l + 1])
            t = (int)transition[l];
    return t;
}

template <typename ValueType> void Acquisition::Coordinates<ValueType>::init(IndexType nx, IndexType ny, IndexType nz, ValueType dh)
{
    SCAI_ASSERT_ERROR(nx > 0 && ny > 0 && nz > 0, "NX, NY and NZ must be positive")
    NX = nx;
    NY = ny;
    NZ = nz;
    DH = dh;
}

template <typename ValueType> void Acquisition::Coordinates<ValueType>::check(IndexType X, IndexType Y, IndexType Z) const
{
    SCAI_ASSERT_ERROR(X < NX && Y < NY && Z < NZ && X >= 0 && Y >= 0 && Z >= 0,
                      "X=" << X << " Y=" << Y << " Z=" << Z << " NX=" << NX << " NY=" << NY << " NZ=" << NZ << " Could not map from coordinate to index!")
}

template <typename ValueType> Acquisition::coordinate3D Acquisition::Coordinates<ValueType>::index2coordinate(IndexType index) const
{
    coordinate3D r;
    if (layered) { // Coordinates.cpp:623-651: position inside the layer's own regular grid, scaled to fine-grid coordinates
        IndexType layer = 0;
        for (; layer < (IndexType)dhFactor.size(); layer++) {
            if (index >= nGridpointsPerLayer[layer])
                index -= nGridpointsPerLayer[layer];
            else
                break;
        }
        SCAI_ASSERT_ERROR(layer < (IndexType)dhFactor.size(), "index outside the model vector")
        const IndexType pl = varNX[layer] * varNZ[layer];
        r.y = index / pl;
        index -= r.y * pl;
        r.z = index / varNX[layer];
        r.x = index - r.z * varNX[layer];
        r.x *= dhFactor[layer];
        r.z *= dhFactor[layer];
        r.y = r.y * dhFactor[layer] + layerStart[layer];
        return r;
    }
    const IndexType plane = NX * NZ;
    r.y = index / plane;
    index -= r.y * plane;
    r.z = index / NX;
    r.x = index - r.z * NX;
    return r;
}

template <typename ValueType> IndexType Acquisition::Coordinates<ValueType>::coordinate2index(IndexType X, IndexType Y, IndexType Z) const
{
    check(X, Y, Z);
    if (layered) { // Coordinates.cpp:668-694 (integer divisions: a fine-grid coordinate inside a coarse layer maps to the cell it lies in)
        IndexType layer = 0;
        for (; layer < (IndexType)dhFactor.size(); layer++)
            if (Y <= layerEnd[layer] && Y >= layerStart[layer]) {
                Y -= layerStart[layer];
                break;
            }
        SCAI_ASSERT_ERROR(layer < (IndexType)dhFactor.size(), "X=" << X << " Y=" << Y << " Z=" << Z << " Could not map from coordinate to index!")
        IndexType index = X / dhFactor[layer] + (Z / dhFactor[layer]) * varNX[layer] + (Y / dhFactor[layer]) * varNX[layer] * varNZ[layer];
        for (IndexType l = 0; l < layer; l++)
            index += nGridpointsPerLayer[l];
        return index;
    }
    return X + Z * NX + Y * NX * NZ;
}

template <typename ValueType> Acquisition::coordinate3D Acquisition::Coordinates<ValueType>::edgeDistance(coordinate3D c) const
{
    check(c.x, c.y, c.z);
    coordinate3D d;
    d.x = std::min(c.x, NX - 1 - c.x);
    d.y = std::min(c.y, NY - 1 - c.y);
    d.z = std::min(c.z, NZ - 1 - c.z);
    return d;
}

template <typename ValueType> void Acquisition::Coordinates<ValueType>::writeCoordinates(std::string const &filename, IndexType fileFormat) const
{
    const IndexType n = getNGridpoints();
    std::vector<float> x((size_t)n), y((size_t)n), z((size_t)n);
    for (IndexType i = 0; i < n; i++) {
        const coordinate3D c = index2coordinate(i);
        x[i] = (float)c.x;
        y[i] = (float)c.y;
        z[i] = (float)c.z;
    }
    IO::writeVector(x, filename + "X", fileFormat);
    IO::writeVector(y, filename + "Y", fileFormat);
    IO::writeVector(z, filename + "Z", fileFormat);
}

template class KITGPI::Acquisition::Coordinates<float>;

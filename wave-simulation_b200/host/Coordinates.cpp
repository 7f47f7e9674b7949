#include "Coordinates.hpp"

using namespace KITGPI;

template <typename ValueType> void Acquisition::Coordinates<ValueType>::init(Configuration::Configuration const &config)
{
    if (config.getAndCatch("useVariableGrid", 0) != 0)
        COMMON_THROWEXCEPTION("useVariableGrid=1 is not available in the B200 path (regular grids only)")
    init(config.get<IndexType>("NX"), config.get<IndexType>("NY"), config.get<IndexType>("NZ"), config.get<ValueType>("DH"));
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

template class KITGPI::Acquisition::Coordinates<float>;

#include "Wavefields.hpp"
#include "DeviceGroup.hpp"
#include "IO.hpp"
#include <algorithm>

using namespace KITGPI;

template <typename ValueType> Wavefields::Wavefields<ValueType>::Wavefields(std::string const &dimension, std::string const &type) : equationType(type)
{
    SCAI_ASSERT_ERROR(dimension == "2d" || dimension == "3d", "Unkown dimension")
    numDimension = dimension == "3d" ? 3 : 2;
    const bool d3 = numDimension == 3;
    if (type == "acoustic") {
        first = d3 ? std::vector<std::string>{"VX", "VY", "VZ"} : std::vector<std::string>{"VX", "VY"};
        second = {"P"};
    } else if (type == "elastic" || type == "viscoelastic") {
        first = d3 ? std::vector<std::string>{"VX", "VY", "VZ"} : std::vector<std::string>{"VX", "VY"};
        second = d3 ? std::vector<std::string>{"Sxx", "Syy", "Szz", "Sxy", "Sxz", "Syz"} : std::vector<std::string>{"Sxx", "Syy", "Sxy"};
    } else if (type == "sh" || type == "viscosh") {
        SCAI_ASSERT_ERROR(!d3, "Unkown type") // WavefieldsFactory.cpp: SH exists in 2D only
        first = {"VZ"};
        second = {"Sxz", "Syz"};
    } else if (type == "tmem" || type == "viscotmem") {
        SCAI_ASSERT_ERROR(!d3, "Unkown type")
        first = {"HX", "HY"};
        second = {"EZ"};
    } else if (type == "emem" || type == "viscoemem") {
        first = d3 ? std::vector<std::string>{"HX", "HY", "HZ"} : std::vector<std::string>{"HZ"};
        second = d3 ? std::vector<std::string>{"EX", "EY", "EZ"} : std::vector<std::string>{"EX", "EY"};
    } else
        COMMON_THROWEXCEPTION("Unkown type")
    init(0);
}

template <typename ValueType> void Wavefields::Wavefields<ValueType>::init(IndexType L)
{
    memory.clear();
    const bool d3 = numDimension == 3;
    for (IndexType l = 1; l <= L; l++) {
        std::vector<std::string> base;
        if (equationType == "viscoelastic")
            base = d3 ? std::vector<std::string>{"Rxx", "Ryy", "Rzz", "Rxy", "Rxz", "Ryz"} : std::vector<std::string>{"Rxx", "Ryy", "Rxy"};
        else if (equationType == "viscosh")
            base = {"Rxz", "Ryz"};
        else if (equationType == "viscotmem")
            base = {"RZ"};
        else if (equationType == "viscoemem")
            base = d3 ? std::vector<std::string>{"RX", "RY", "RZ"} : std::vector<std::string>{"RX", "RY"};
        for (auto const &b : base)
            memory.push_back(b + std::to_string(l));
    }
    all = first;
    all.insert(all.end(), second.begin(), second.end());
    all.insert(all.end(), memory.begin(), memory.end());
}

template <typename ValueType> void Wavefields::Wavefields<ValueType>::init(Wavefields<ValueType> const &like)
{
    SCAI_ASSERT_ERROR(like.h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    SCAI_ASSERT_ERROR(like.equationType == equationType && like.numDimension == numDimension, "wavefield objects of different type")
    if (stored && h && groupAlive && *groupAlive)
        h->destroyFieldSet(own);
    h = like.h;
    groupAlive = h->aliveToken();
    memory = like.memory;
    all = like.all;
    own = h->createFieldSet();
    stored = true;
}

template <typename ValueType> Wavefields::Wavefields<ValueType>::~Wavefields()
{
    if (stored && h && groupAlive && *groupAlive) // (a solver that went first has released the components itself)
        h->destroyFieldSet(own);
}

namespace
{
    typedef KITGPI::ForwardSolver::DeviceGroup::FieldSet FieldSet;
}
#define WS_WF_CHECK(rhs)                                                                                                                  \
    SCAI_ASSERT_ERROR(!stored || (groupAlive && *groupAlive), "the forward solver of this wavefield object is gone")                              \
    SCAI_ASSERT_ERROR(h && (rhs).h == h, "wavefield operators need two objects on the same forward solver (Wavefields::init(like))")

template <typename ValueType> Wavefields::Wavefields<ValueType> &Wavefields::Wavefields<ValueType>::operator=(Wavefields<ValueType> &rhs)
{
    WS_WF_CHECK(rhs)
    h->fieldSetBinary(stored ? &own : nullptr, rhs.stored ? &rhs.own : nullptr, 0);
    return *this;
}
template <typename ValueType> Wavefields::Wavefields<ValueType> &Wavefields::Wavefields<ValueType>::operator+=(Wavefields<ValueType> &rhs)
{
    WS_WF_CHECK(rhs)
    h->fieldSetBinary(stored ? &own : nullptr, rhs.stored ? &rhs.own : nullptr, 1);
    return *this;
}
template <typename ValueType> Wavefields::Wavefields<ValueType> &Wavefields::Wavefields<ValueType>::operator-=(Wavefields<ValueType> &rhs)
{
    WS_WF_CHECK(rhs)
    h->fieldSetBinary(stored ? &own : nullptr, rhs.stored ? &rhs.own : nullptr, 2);
    return *this;
}
template <typename ValueType> Wavefields::Wavefields<ValueType> &Wavefields::Wavefields<ValueType>::operator*=(ValueType rhs)
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    h->fieldSetScale(stored ? &own : nullptr, rhs);
    return *this;
}
template <typename ValueType> Wavefields::Wavefields<ValueType> &Wavefields::Wavefields<ValueType>::operator*=(std::vector<ValueType> const &rhs)
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    h->fieldSetScale(stored ? &own : nullptr, rhs);
    return *this;
}

template <typename ValueType> void Wavefields::Wavefields<ValueType>::resetWavefields()
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    if (stored) { // a stored object: all components back to zero
        h->fieldSetScale(&own, ValueType(0));
        return;
    }
    h->forEach([&](IndexType r) {
        if (ws_reset(h->handle(r)) != WS_OK)
            COMMON_THROWEXCEPTION(ws_last_error())
    });
}

template <typename ValueType> bool Wavefields::Wavefields<ValueType>::isFinite() const
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    return h->isFinite();
}

template <typename ValueType> std::vector<ValueType> Wavefields::Wavefields<ValueType>::get(std::string const &component) const
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    if (stored)
        return h->getWavefield(own, component);
    return h->getWavefield(component);
}

template <typename ValueType> void Wavefields::Wavefields<ValueType>::set(std::string const &component, std::vector<ValueType> const &values)
{
    SCAI_ASSERT_ERROR(h, "The wavefields are not bound to a forward solver yet (initForwardSolver)")
    SCAI_ASSERT_ERROR(!stored, "set() addresses the solver's own wavefields")
    h->setWavefield(component, values);
}

template <typename ValueType> void Wavefields::Wavefields<ValueType>::write(IndexType snapType, std::string baseName, IndexType t, IndexType fileFormat) const
{
    const std::string timeStep = std::to_string(static_cast<long long>(t));
    switch (snapType) {
    case 1:
        for (auto const &c : first)
            IO::writeVector(get(c), baseName + "." + c + "." + timeStep, fileFormat);
        break;
    case 2:
        for (auto const &c : second)
            IO::writeVector(get(c), baseName + "." + c + "." + timeStep, fileFormat);
        break;
    case 3: // Wavefields3Delastic.cpp:82-92, Wavefields2Delastic.cpp:75-85: energy of the S- and P-wave parts
        if (equationType == "elastic" || equationType == "viscoelastic") {
            IO::writeVector(get("CURL"), baseName + ".CURL." + timeStep, fileFormat);
            IO::writeVector(get("DIV"), baseName + ".DIV." + timeStep, fileFormat);
        } else if (((equationType == "tmem" || equationType == "viscotmem") && numDimension == 2) || ((equationType == "emem" || equationType == "viscoemem") && numDimension == 3)) {
            // WavefieldsEM/Wavefields2Dtmem.cpp:75-87, Wavefields3Demem.cpp:70-85: lower-case component names
            IO::writeVector(get("CURL"), baseName + ".curl." + timeStep, fileFormat);
            IO::writeVector(get("DIV"), baseName + ".div." + timeStep, fileFormat);
        } else
            COMMON_THROWEXCEPTION("There is no curl or div of wavefield in the " << numDimension << "D " << equationType << " case.")
        break;
    default: COMMON_THROWEXCEPTION("Invalid snapType.")
    }
}

template <typename ValueType> typename Wavefields::Wavefields<ValueType>::WavefieldPtr Wavefields::Factory<ValueType>::Create(std::string dimension, std::string type)
{
    std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
    std::transform(type.begin(), type.end(), type.begin(), ::tolower);
    return std::make_shared<Wavefields<ValueType>>(dimension, type);
}

template class Wavefields::Wavefields<float>;
template class Wavefields::Factory<float>;

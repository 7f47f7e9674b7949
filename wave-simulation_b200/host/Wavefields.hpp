// Wavefields.hpp — wavefield state (mirror of src/Wavefields/Wavefields.hpp:25-215 and src/WavefieldsEM for the forward path).
// The state itself lives in HBM inside the solver handle (padded arenas); this class keeps the reference's
// component names (VX..Sxy, P, Rxx1.., HX..EZ, RX1..), resetWavefields, getRef*-style access (copies through
// ws_get_wavefield / ws_set_wavefield) and the snapshot writer `<base>.<COMP>.<tStep>.<mtx|lmf>`
// (Wavefields3Delastic.cpp:62-96; snapType 1 = particle velocities / H, 2 = stresses / pressure / E).
#pragma once
#include "Common.hpp"
#include "../../include/wavesim.h"
#include <memory>

namespace KITGPI { namespace ForwardSolver { class DeviceGroup; } }

namespace KITGPI
{
    namespace Wavefields
    {
        template <typename ValueType> class Wavefields
        {
          public:
            typedef std::shared_ptr<Wavefields<ValueType>> WavefieldPtr;
            Wavefields(std::string const &dimension, std::string const &type);

            //! number of relaxation mechanisms fixes the memory-variable components (Wavefields3Dviscoelastic.cpp init)
            void init(IndexType numRelaxationMechanisms);
            void resetWavefields();
            bool isFinite() const;
            std::vector<std::string> const &getComponents() const { return all; }
            std::vector<ValueType> get(std::string const &component) const;       // e.g. "VX", "Sxy", "P", "EZ", "Rxx1"
            void set(std::string const &component, std::vector<ValueType> const &values);
            //! snapType 1 = first half-step fields, 2 = second half-step fields, 3 = curl / div energy measures (elastic, viscoelastic)
            void write(IndexType snapType, std::string baseName, IndexType t, IndexType fileFormat) const;
            std::string getEquationType() const { return equationType; }
            IndexType getNumDimension() const { return numDimension; }

            //! the GPUs that hold the state (set by ForwardSolver::initForwardSolver)
            void bind(ForwardSolver::DeviceGroup *group) { h = group; }

            //! a second wavefield object next to the solver's own (e.g. `wavefieldsTemp`, Simulation.cpp:327): same components, zero,
            //! resident on the GPUs of `like` (takes the place of Wavefields::init(ctx, dist, numRelaxationMechanisms))
            void init(Wavefields<ValueType> const &like);
            ~Wavefields();
            Wavefields(Wavefields const &) = delete;

            //! Operator overloading (Wavefields.hpp:62-80): applied to every component incl. the memory variables, on the GPUs
            Wavefields<ValueType> &operator=(Wavefields<ValueType> &rhs);
            Wavefields<ValueType> &operator-=(Wavefields<ValueType> &rhs);
            Wavefields<ValueType> &operator+=(Wavefields<ValueType> &rhs);
            Wavefields<ValueType> &operator*=(ValueType rhs);
            Wavefields<ValueType> &operator*=(std::vector<ValueType> const &rhs);

          private:
            std::string equationType;
            IndexType numDimension;
            std::vector<std::string> first, second, memory, all; // first half-step fields, second half-step fields, memory variables
            ForwardSolver::DeviceGroup *h = nullptr;
            //! empty: this object IS the solver's wavefields; else its own component set per GPU of the group
            std::vector<ws_wavefields *> own;
            bool stored = false;
            std::shared_ptr<bool> groupAlive; // a stored object may outlive the solver it was created on
        };

        template <typename ValueType> class Factory
        {
          public:
            //! (2D|3D) x (acoustic, elastic, viscoelastic, sh, viscosh, tmem, emem, viscotmem, viscoemem): WavefieldsFactory.cpp:6
            static typename Wavefields<ValueType>::WavefieldPtr Create(std::string dimension, std::string type);
        };
    }
}

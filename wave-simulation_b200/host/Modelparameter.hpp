// Modelparameter.hpp — material parameters (mirror of src/Modelparameter/*.hpp and src/ModelparameterEM/*.hpp for the
// forward-modelling path).  The host object holds the RAW parameter vectors in the reference's linear-index order
// (homogeneous from the configuration when ModelRead=0, read from `<ModelFilename>.<vp|vs|density|tauP|tauS|mur|sigma|
// epsilonr|tauSigmar|tauEpsilon>.<mtx|lmf>` when ModelRead=1).  The prepareForModelling products (moduli, staggered
// averages, EM coefficients: Elastic.cpp:28-45, Modelparameter.cpp:336-651, ForwardSolverEM.cpp:14-154) are computed on
// the GPU by ws_prepare and can be read back through getParameter() with the reference getter names.
#pragma once
#include "Common.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include <map>
#include <memory>

namespace KITGPI { namespace ForwardSolver { class DeviceGroup; } }

namespace KITGPI
{
    namespace Modelparameter
    {
        template <typename ValueType> class Modelparameter
        {
          public:
            typedef std::shared_ptr<Modelparameter<ValueType>> ModelparameterPtr;
            explicit Modelparameter(std::string const &type);

            void init(Configuration::Configuration const &config, Acquisition::Coordinates<ValueType> const &modelCoordinates);
            void init(std::string const &name, std::vector<ValueType> const &values); // set one raw parameter (C-ABI name)
            void write(std::string filename, IndexType fileFormat) const;
            //! marks the model as ready; the averaging itself runs in ForwardSolver::prepareForModelling (ws_prepare)
            void prepareForModelling() { dirtyFlag = false; }

            std::string getEquationType() const { return equationType; }
            bool isSeismic() const { return seismic; }
            IndexType getNumRelaxationMechanisms() const { return (IndexType)relaxationFrequency.size(); }
            std::vector<ValueType> const &getRelaxationFrequency() const { return relaxationFrequency; }

            //! raw parameters by their C-ABI / reference getter names ("velocityP", "density", "magneticPermeability", ...)
            std::map<std::string, std::vector<ValueType>> const &getRawParameters() const { return *raw; }
            std::vector<ValueType> const &getVelocityP() const { return at("velocityP"); }
            std::vector<ValueType> const &getVelocityS() const { return at("velocityS"); }
            std::vector<ValueType> const &getDensity() const { return at("density"); }
            std::vector<ValueType> const &getTauP() const { return at("tauP"); }
            std::vector<ValueType> const &getTauS() const { return at("tauS"); }
            std::vector<ValueType> const &getMagneticPermeability() const { return at("magneticPermeability"); }
            std::vector<ValueType> const &getElectricConductivity() const { return at("electricConductivity"); }
            std::vector<ValueType> const &getDielectricPermittivity() const { return at("dielectricPermittivity"); }
            //! derived parameter read back from the device ("pWaveModulus", "inverseDensityAverageX", "CaAverageZ", ...)
            std::vector<ValueType> getParameter(std::string const &name) const;
            //! largest propagation velocity (vp, vs for SH, c0/sqrt(eps_r mu_r) for EM): CheckParameter.hpp:183-200
            ValueType getMaxVelocity() const;
            ValueType getMinVelocity() const;
            //! the part of this (big) model a shot works on: the box of `modelCoordinates` starting at cutCoordinate (useStreamConfig;
            //! Elastic.cpp:104-140 getModelPerShot: shrink matrix = one 1 per row)
            void getModelPerShot(Modelparameter<ValueType> &modelPerShot, Acquisition::Coordinates<ValueType> const &modelCoordinates,
                                 Acquisition::Coordinates<ValueType> const &modelCoordinatesBig, Acquisition::coordinate3D const &cutCoordinate) const;
            //! exp(sigma / eps * tStep * DT), the amplitude compensation of an EM modelling (Modelparameter.cpp:127-143)
            std::vector<ValueType> getCompensation(ValueType DT, IndexType tStep) const;

            void bind(ForwardSolver::DeviceGroup *group) { h = group; }

            static constexpr ValueType MagneticPermeabilityVacuum = 1.2566370614e-6f;    // Modelparameter.hpp:359
            static constexpr ValueType DielectricPermittivityVacuum = 8.8541878176e-12f; // Modelparameter.hpp:360

          private:
            std::vector<ValueType> const &at(std::string const &name) const;
            std::string equationType;
            bool seismic = true;
            bool dirtyFlag = true;
            //! shared between the copies the shot domains work on (one copy of a 768^3 model is 9 GB of host memory); a copy that
            //! changes a parameter gets its own map (init)
            std::shared_ptr<std::map<std::string, std::vector<ValueType>>> raw = std::make_shared<std::map<std::string, std::vector<ValueType>>>();
            std::vector<ValueType> relaxationFrequency;
            ValueType centerFrequencyCPML = 0; // visco-EM: reference frequency of the effective <-> static conversion
            void relaxationAverages(ValueType &aAverage, ValueType &bAverage) const;
            ForwardSolver::DeviceGroup *h = nullptr;
        };

        template <typename ValueType> class Factory
        {
          public:
            //! acoustic, elastic, viscoelastic, sh, viscosh, tmem, emem, viscotmem, viscoemem (ModelparameterFactory.cpp:4-43)
            static typename Modelparameter<ValueType>::ModelparameterPtr Create(std::string type);
        };
    }
}

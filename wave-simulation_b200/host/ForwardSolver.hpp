// ForwardSolver.hpp — the time-stepping interface of the reference on top of the C ABI (include/wavesim.h).
// Mirror of src/ForwardSolver/ForwardSolver.hpp:38-52 and ForwardSolverFactory.cpp:4-66: one class serves all 14
// (dimension, equationType) pairs because the physics lives in the CUDA library; the reference's LAMA arguments
// (DistributionPtr, ContextPtr) have no counterpart.  Errors of the library surface as KITGPI::Exception with the
// library's message, i.e. the reference's abort-with-message behaviour.
#pragma once
#include "Acquisition.hpp"
#include "Common.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include "DeviceGroup.hpp"
#include "Modelparameter.hpp"
#include "Wavefields.hpp"

struct ws_solver;
struct ws_desc;

namespace KITGPI
{
    namespace ForwardSolver
    {
        template <typename ValueType> class ForwardSolver
        {
          public:
            typedef std::shared_ptr<ForwardSolver<ValueType>> ForwardSolverPtr;
            ForwardSolver(std::string const &dimension, std::string const &type);
            ~ForwardSolver();
            ForwardSolver(ForwardSolver const &) = delete;
            ForwardSolver &operator=(ForwardSolver const &) = delete;

            //! CUDA device used by this solver (one solver per shot domain); call before initForwardSolver
            void setDevice(IndexType device) { deviceIds.assign(1, device); }
            //! several GPUs for one shot domain: the grid is cut into y-slabs, one per device (replaces the reference's
            //! spatial Partitioning over the processes of commShot, Simulation.cpp:126-149)
            void setDevices(std::vector<IndexType> const &devices) { deviceIds = devices; }
            std::vector<IndexType> const &getDevices() const { return deviceIds; }

            //! memory in MB the solver will allocate in HBM (wavefields, model, CPML slabs): estimateMemory of the reference
            ValueType estimateMemory(Configuration::Configuration const &config, Acquisition::Coordinates<ValueType> const &modelCoordinates);

            //! creates the device solver (Factory::Create + Wavefields::init + Derivatives::init), uploads the raw model and
            //! binds `wavefield` / `model` to it
            void initForwardSolver(Configuration::Configuration const &config, Derivatives::Derivatives<ValueType> &derivatives, Wavefields::Wavefields<ValueType> &wavefield,
                                   Modelparameter::Modelparameter<ValueType> &model, Acquisition::Coordinates<ValueType> const &modelCoordinates, ValueType DT);
            //! another model on the same grid (the cut-out of the next shot, useStreamConfig): re-uploads the raw parameters into the
            //! existing device solver; prepareForModelling must follow
            void updateModel(Modelparameter::Modelparameter<ValueType> &model);
            //! Modelparameter::prepareForModelling products + boundary coefficients (CPML / ABS / free surface) on the GPU
            void prepareForModelling(Modelparameter::Modelparameter<ValueType> const &model, ValueType DT);
            void prepareBoundaryConditions(Configuration::Configuration const &, Acquisition::Coordinates<ValueType> const &, Derivatives::Derivatives<ValueType> &) {}
            void resetCPML();

            //! one time step t (velocity/H half-step, stress/E half-step, free surface, damping, sources, receivers);
            //! asynchronous; the seismograms are copied into `receiver` after the last step (t = NT-1)
            void run(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources, Modelparameter::Modelparameter<ValueType> const &model,
                     Wavefields::Wavefields<ValueType> &wavefield, Derivatives::Derivatives<ValueType> const &derivatives, IndexType t);
            //! steps t0..t1-1 as one CUDA-graph batched call (extension; same result as the loop over run)
            void run(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources, IndexType t0, IndexType t1);
            void sync();

            ws_solver *handle() const { return group ? group->handle(0) : nullptr; }
            DeviceGroup *getGroup() const { return group.get(); }
            IndexType getNT() const { return NT; }

          private:
            void initIrregular(Configuration::Configuration const &config, ws_desc d, Derivatives::Derivatives<ValueType> &derivatives, Wavefields::Wavefields<ValueType> &wavefield,
                               Modelparameter::Modelparameter<ValueType> &model, Acquisition::Coordinates<ValueType> const &modelCoordinates);
            void bindAcquisition(Acquisition::Receivers<ValueType> &receiver, Acquisition::Sources<ValueType> const &sources);
            void fetchSeismograms(Acquisition::Receivers<ValueType> &receiver);
            std::string dimension, equationType;
            std::unique_ptr<DeviceGroup> group;
            std::vector<IndexType> deviceIds{0};
            IndexType NT = 0;
            unsigned long srcVersion = ~0ul, recVersion = ~0ul;
            const void *srcObj = nullptr, *recObj = nullptr;
        };

        template <typename ValueType> class Factory
        {
          public:
            static typename ForwardSolver<ValueType>::ForwardSolverPtr Create(std::string dimension, std::string type);
        };
    }
}

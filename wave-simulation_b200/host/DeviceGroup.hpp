// DeviceGroup.hpp — the GPUs of one shot domain.  Replaces the role of the reference's commShot / dist pair
// (src/Simulation.cpp:116-149, src/Partitioning/Partitioning.hpp:43-81): the grid of a shot domain is cut into y-slabs, one
// per GPU (ws_desc.rank / nranks), and the slabs exchange halo planes over NCCL inside the CUDA library.  Every rank is
// driven by its own persistent host thread (NCCL needs the ranks of a communicator to call concurrently); forEach()
// is the fork-join the host classes use for every collective call.  With one GPU the calls run inline on the caller's
// thread.  Vectors that cross this interface are GLOBAL vectors in the reference's linear-index order (y slowest), so
// a rank's slab is the contiguous range [y0, y0 + nyl) * NX * NZ.
#pragma once
#include "Common.hpp"
#include "../../include/wavesim.h"
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace KITGPI
{
    namespace ForwardSolver
    {
        class DeviceGroup
        {
          public:
            explicit DeviceGroup(std::vector<IndexType> const &devices);
            ~DeviceGroup();
            DeviceGroup(DeviceGroup const &) = delete;
            DeviceGroup &operator=(DeviceGroup const &) = delete;

            IndexType size() const { return (IndexType)devices.size(); }
            ws_solver *handle(IndexType rank = 0) const { return handles[rank]; }
            //! runs fn(rank) for every rank concurrently and rethrows the first failure as KITGPI::Exception
            void forEach(std::function<void(IndexType)> const &fn);
            //! ws_create for every rank of the group (desc.rank / nranks / device are filled in) + NCCL communicator
            void create(ws_desc desc);
            //! operator-given mode (irregular grids): one GPU, model vectors of nPoints values (ws_create_sparse)
            void createSparse(ws_desc desc, size_t nPoints);
            void destroy();

            size_t getNGlobal() const { return nGlobal; }
            //! wavefield component / derived model parameter as a global vector (every rank contributes its slab)
            std::vector<float> getWavefield(std::string const &component);
            void setWavefield(std::string const &component, std::vector<float> const &values);
            std::vector<float> getMaterial(std::string const &name);
            //! seismogram rows of all receivers (n_rec x NT): every receiver is recorded by the rank that owns its grid point
            void getSeismogram(std::vector<float> &all);
            bool isFinite();

            //! a second set of wavefield components on the GPUs of the group (ws_wavefields, one per rank); the operators below take
            //! nullptr for the solvers' own wavefields (Wavefields/Wavefields.hpp:62-80)
            typedef std::vector<ws_wavefields *> FieldSet;
            //! false once the group (the forward solver) is gone: a stored wavefield object that outlives its solver must not touch it
            std::shared_ptr<bool> const &aliveToken() const { return alive; }
            FieldSet createFieldSet();
            void destroyFieldSet(FieldSet &set);
            //! op 0: dst = src, 1: dst += src, 2: dst -= src
            void fieldSetBinary(FieldSet const *dst, FieldSet const *src, int op);
            void fieldSetScale(FieldSet const *dst, float rhs);
            void fieldSetScale(FieldSet const *dst, std::vector<float> const &rhs); // global vector, NX*NY*NZ values
            std::vector<float> getWavefield(FieldSet const &set, std::string const &component);
            //! `*wavefields *= vector` after every time step inside the library (Simulation.cpp:455-456); empty vector = off
            void setStepScaling(std::vector<float> const &vec);

          private:
            void workerLoop(IndexType rank);
            std::vector<IndexType> devices;
            std::vector<ws_solver *> handles;
            std::vector<IndexType> y0, nyl;
            size_t nGlobal = 0, planeSize = 0;
            // fork-join over the rank threads
            std::vector<std::thread> workers;
            std::mutex m;
            std::condition_variable cvStart, cvDone;
            std::function<void(IndexType)> const *task = nullptr;
            unsigned long generation = 0;
            IndexType pending = 0;
            bool stop = false;
            std::vector<std::string> errors;
            std::shared_ptr<bool> alive = std::make_shared<bool>(true);
            std::vector<ws_wavefields *> liveSets; // released with the group if their owner has not done it
        };
    }
}

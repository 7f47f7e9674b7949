// Host-layer mirror of the wavefield operators (Wavefields/Wavefields.hpp:62-80) the reference's time loop and WAVE-Inversion
// use (Simulation.cpp:327, 450-461): `*wavefieldsTemp = *wavefields; run; *wavefieldsTemp -= *wavefields; *wavefieldsTemp *= -DTinv`,
// `+=`, `*= vector`, and Modelparameter::getCompensation.  TEST INFRASTRUCTURE: linked against the host emulation build of the
// library by tests/test_host_layer.py (the same program runs against the CUDA library in the -m gpu test).
// usage: test_wavefield_ops <configuration file>
#include "Acquisition.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include "ForwardSolver.hpp"
#include "Modelparameter.hpp"
#include "Wavefields.hpp"
#include <cmath>
#include <cstdio>

using namespace KITGPI;

static int failures = 0;
#define EXPECT(cond)                                                                                                   \
    if (!(cond)) {                                                                                                     \
        std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                                  \
        failures++;                                                                                                    \
    }

int main(int argc, char **argv)
{
    if (argc != 2)
        return 2;
    try {
        Configuration::Configuration config(argv[1]);
        std::string dimension = config.get<std::string>("dimension"), equationType = config.get<std::string>("equationType");
        std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
        std::transform(equationType.begin(), equationType.end(), equationType.begin(), ::tolower);
        const ValueType DT = config.get<ValueType>("DT");
        Acquisition::Coordinates<ValueType> modelCoordinates(config);
        auto model = Modelparameter::Factory<ValueType>::Create(equationType);
        model->init(config, modelCoordinates);
        auto derivatives = ForwardSolver::Derivatives::Factory<ValueType>::Create(dimension);
        auto wavefields = Wavefields::Factory<ValueType>::Create(dimension, equationType);
        auto wavefieldsTemp = Wavefields::Factory<ValueType>::Create(dimension, equationType);
        auto sum = Wavefields::Factory<ValueType>::Create(dimension, equationType);
        auto solver = ForwardSolver::Factory<ValueType>::Create(dimension, equationType);
        derivatives->init(config);
        solver->initForwardSolver(config, *derivatives, *wavefields, *model, modelCoordinates, DT);
        model->prepareForModelling();
        solver->prepareForModelling(*model, DT);
        Acquisition::Sources<ValueType> sources;
        sources.getAcquisitionSettings(config);
        sources.init(sources.getSourceSettings(), config, modelCoordinates);
        Acquisition::Receivers<ValueType> receivers;
        receivers.init(config, modelCoordinates);
        wavefieldsTemp->init(*wavefields); // Wavefields::init(ctx, dist, numRelaxationMechanisms) of a second object
        sum->init(*wavefields);
        wavefields->resetWavefields();
        for (IndexType t = 0; t < 8; t++)
            solver->run(receivers, sources, *model, *wavefields, *derivatives, t);
        solver->sync();
        std::vector<std::vector<ValueType>> before, after;
        for (auto const &c : wavefields->getComponents())
            before.push_back(wavefields->get(c));
        *wavefieldsTemp = *wavefields; // Simulation.cpp:450
        solver->run(receivers, sources, *model, *wavefields, *derivatives, 8);
        solver->sync();
        for (auto const &c : wavefields->getComponents())
            after.push_back(wavefields->get(c));
        *wavefieldsTemp -= *wavefields;
        const ValueType DTinv = 1 / DT;
        *wavefieldsTemp *= -DTinv; // :458-459: time derivative of the wavefields
        *sum += *wavefieldsTemp;
        *sum += *wavefields;
        bool moved = false;
        size_t k = 0;
        for (auto const &c : wavefields->getComponents()) {
            auto d = wavefieldsTemp->get(c), s = sum->get(c);
            bool okD = true, okS = true;
            for (size_t i = 0; i < d.size(); i++) {
                const ValueType want = (before[k][i] - after[k][i]) * -DTinv;
                okD = okD && d[i] == want;
                okS = okS && s[i] == want + after[k][i];
                moved = moved || before[k][i] != after[k][i];
            }
            EXPECT(okD);
            EXPECT(okS);
            k++;
        }
        EXPECT(moved);
        // *= vector on the solver's own wavefields, then a stored object back into the solver
        std::vector<ValueType> vec(before[0].size());
        for (size_t i = 0; i < vec.size(); i++)
            vec[i] = 0.5f + 0.001f * (ValueType)(i % 977);
        *wavefields *= vec;
        k = 0;
        for (auto const &c : wavefields->getComponents()) {
            auto v = wavefields->get(c);
            bool ok = true;
            for (size_t i = 0; i < v.size(); i++)
                ok = ok && v[i] == after[k][i] * vec[i];
            EXPECT(ok);
            k++;
        }
        *wavefields = *sum;
        EXPECT(wavefields->get(wavefields->getComponents()[0]) == sum->get(wavefields->getComponents()[0]));
        sum->resetWavefields();
        for (auto x : sum->get(wavefields->getComponents()[0]))
            if (x != 0) {
                EXPECT(false);
                break;
            }
        if (model->isSeismic()) {
            bool thrown = false;
            try {
                model->getCompensation(DT, 1);
            } catch (std::exception const &) {
                thrown = true;
            }
            EXPECT(thrown); // "There is no compensation in an Seismic modelling"
        } else {
            auto comp = model->getCompensation(DT, 3);
            auto const &sg = model->getElectricConductivity();
            auto const &ep = model->getDielectricPermittivity();
            EXPECT(comp.size() == sg.size() && std::abs(comp[5] - std::exp(sg[5] / ep[5] * 3 * DT)) <= 1e-6f * comp[5] && comp[5] > 1.0f);
        }
    } catch (std::exception const &e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "%d FAILURES\n" : "wavefield operator tests OK\n", failures);
    return failures ? 1 : 0;
}

// Simulation.cpp — forward-modelling driver with the flow of the reference's src/Simulation.cpp:36-574:
// configuration -> factories -> acquisition -> model -> forward solver -> loop over shots -> loop over time steps ->
// seismograms.  Runs the `par/` configurations unchanged on the CUDA library (include/wavesim.h).  The GPUs of the box
// take the place of the reference's MPI processes: they are split into NumShotDomains groups (Simulation.cpp:116-121),
// every group works on its block of the shots (:369) and cuts the grid into y-slabs over its GPUs (the reference's
// spatial partitioning over commShot, :126-149; the `partitioning` key is accepted and always means y-slabs here).
#include "Acquisition.hpp"
#include "CheckParameter.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include "ForwardSolver.hpp"
#include "Modelparameter.hpp"
#include "Wavefields.hpp"
#include "../../include/wavesim.h"
#include <chrono>
#include <ctime>
#include <mutex>
#include <thread>

using namespace KITGPI;

namespace
{
    double now()
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    std::mutex printMutex;
}

// what the shot domains share: the acquisition of the survey as the main thread has set it up (Simulation.cpp:233-282)
struct Survey {
    std::vector<Acquisition::sourceSettings<ValueType>> sourceSettings; // plain or encoded (useSourceEncode): the list the shots are taken from
    std::vector<IndexType> uniqueShotNos;                               // ... and its shot numbers
    std::vector<Acquisition::sourceSettings<ValueType>> sourceSettingsEncode; // encoded list (empty without useSourceEncode)
    bool useStreamConfig = false;                                       // per-shot model cut-outs of a big model
    std::vector<Acquisition::coordinate3D> cutCoordinates;
    Acquisition::Coordinates<ValueType> modelCoordinatesBig;
    IndexType useSourceEncode = 0;
    IndexType numshotsAll = 0;            // shots of the source file before the shotIncr selection (rows of the receiver mark matrix)
    std::vector<IndexType> shotIndsIncr;  // rows of the selected shots
    // common-offset profiles (Simulation.cpp:356-361): every shot has ONE source and there are several shots -> the source signals
    // (writeSource) and, when a shot records one trace, the seismograms are gathered into one matrix of numshots traces
    bool copCondition = false, copSources = false, copReceivers = false;
};

// what a shot domain hands back for the reduction over the domains (sumShotDomain, Simulation.cpp:535-560)
struct DomainResult {
    Acquisition::SeismogramHandler<ValueType> sources, receivers;
};

// one shot domain = one group of GPUs: the shots `shotInds` (indices into the unique shot list; block distribution of the shots or,
// with useRandomSource, one shot of every pass: dmemo::blockDistribution(..., commInterShot), Simulation.cpp:339-344, 362-369)
static void runShotDomain(Configuration::Configuration const &config, IndexType shotDomain, std::vector<IndexType> devices, std::vector<IndexType> shotInds, Survey const &survey,
                          Modelparameter::Modelparameter<ValueType>::ModelparameterPtr model, Acquisition::Coordinates<ValueType> const &modelCoordinates, double globalStart_t,
                          std::string *error, DomainResult *result)
{
    auto const &sourceSettings = survey.sourceSettings;
    auto const &uniqueShotNos = survey.uniqueShotNos;
    try {
        std::string dimension = config.get<std::string>("dimension"), equationType = config.get<std::string>("equationType");
        std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
        std::transform(equationType.begin(), equationType.end(), equationType.begin(), ::tolower);
        const ValueType DT = config.get<ValueType>("DT");
        const IndexType tStepEnd = Common::time2index(config.get<ValueType>("T"), DT);
        const IndexType numshots = (IndexType)uniqueShotNos.size();

        auto derivatives = ForwardSolver::Derivatives::Factory<ValueType>::Create(dimension);
        auto wavefields = Wavefields::Factory<ValueType>::Create(dimension, equationType);
        auto solver = ForwardSolver::Factory<ValueType>::Create(dimension, equationType);
        solver->setDevices(devices);
        derivatives->init(config);

        // every domain works on its own copy of the model object (binding to its solver); the raw vectors are shared data.  With
        // useStreamConfig the model of the domain is the cut-out of the current shot (Simulation.cpp:387-398).
        Modelparameter::Modelparameter<ValueType> modelLocal(*model);
        double start_t = now();
        if (survey.useStreamConfig && !shotInds.empty()) {
            IndexType first = shotInds[0];
            if (survey.useSourceEncode == 3)
                Acquisition::getuniqueShotInd(first, survey.sourceSettingsEncode, uniqueShotNos[shotInds[0]]);
            model->getModelPerShot(modelLocal, modelCoordinates, survey.modelCoordinatesBig, survey.cutCoordinates.at(first));
        }
        solver->initForwardSolver(config, *derivatives, *wavefields, modelLocal, modelCoordinates, DT);
        modelLocal.prepareForModelling();
        solver->prepareForModelling(modelLocal, DT);
        HOST_PRINT("", "Finished initializing forward solver of shot domain " << shotDomain << " in " << now() - start_t << " sec.\n\n")

        Acquisition::Sources<ValueType> sources;
        Acquisition::Receivers<ValueType> receivers;
        if (config.get<IndexType>("useReceiversPerShot") == 0)
            receivers.init(config, modelCoordinates);
        if (survey.copSources)
            sources.getSeismogramHandler().allocateCOP(numshots, tStepEnd);
        if (survey.copReceivers)
            receivers.getSeismogramHandler().allocateCOP(numshots, tStepEnd);

        const IndexType snapType = config.get<IndexType>("snapType");
        const double tInit = now() - globalStart_t;
        bool firstShot = true;
        for (IndexType shotInd : shotInds) {
            const IndexType shotNumber = uniqueShotNos[shotInd];
            std::vector<Acquisition::sourceSettings<ValueType>> sourceSettingsShot;
            Acquisition::createSettingsForShot(sourceSettingsShot, sourceSettings, shotNumber);
            sources.init(sourceSettingsShot, config, modelCoordinates);
            const IndexType shotIndIncr = survey.useSourceEncode == 0 && shotInd < (IndexType)survey.shotIndsIncr.size() ? survey.shotIndsIncr[shotInd] : shotInd;
            if (survey.copSources)
                sources.getSeismogramHandler().setShotInd(shotInd, shotIndIncr); // Simulation.cpp:384-386
            IndexType shotIndPerShot = shotInd;
            if (survey.useStreamConfig) {
                // Simulation.cpp:387-398: switch to the model subset of this shot
                if (survey.useSourceEncode == 3)
                    Acquisition::getuniqueShotInd(shotIndPerShot, survey.sourceSettingsEncode, shotNumber);
                if (!firstShot) {
                    model->getModelPerShot(modelLocal, modelCoordinates, survey.modelCoordinatesBig, survey.cutCoordinates.at(shotIndPerShot));
                    solver->updateModel(modelLocal);
                    modelLocal.prepareForModelling();
                    solver->prepareForModelling(modelLocal, DT);
                }
                modelLocal.write(config.get<std::string>("ModelFilename") + ".shot_" + std::to_string(shotNumber), config.get<IndexType>("FileFormat"));
            }
            firstShot = false;
            CheckParameter::checkNumericalArtefactsAndInstabilities<ValueType>(config, sourceSettingsShot, modelLocal, modelCoordinates, shotNumber);
            if (config.getAndCatch("writeSource", false))
                sources.getSeismogramHandler().write(config.get<IndexType>("SeismogramFormat"), config.get<std::string>("writeSourceFilename") + ".shot_" + std::to_string(shotNumber), &modelCoordinates);
            if (survey.useStreamConfig) {
                // Receivers.cpp:86-121: the receivers of the shot are given in the big model and moved into its cut-out
                std::vector<Acquisition::receiverSettings> bigSettings, shotSettings;
                if (config.get<IndexType>("useReceiversPerShot") == 2)
                    receivers.getAcquisitionSettings(config, bigSettings, shotNumber, survey.numshotsAll, survey.shotIndsIncr, survey.sourceSettingsEncode);
                else
                    Acquisition::readAllSettings(bigSettings, config.get<std::string>("ReceiverFilename") + ".shot_" + std::to_string(shotNumber) + ".txt");
                Acquisition::getSettingsPerShot<ValueType>(shotSettings, bigSettings, survey.cutCoordinates.at(shotIndPerShot), modelCoordinates, config.get<IndexType>("BoundaryWidth"));
                receivers.init(shotSettings, config, modelCoordinates);
            } else if (config.get<IndexType>("useReceiversPerShot") == 2)
                receivers.init(config, modelCoordinates, shotNumber, survey.numshotsAll, survey.shotIndsIncr, survey.sourceSettingsEncode);
            else if (config.get<IndexType>("useReceiversPerShot") != 0)
                receivers.init(config, modelCoordinates, shotNumber);
            receivers.getSeismogramHandler().resetData();
            if (survey.copCondition && receivers.getNumTracesGlobal() == 1)
                receivers.getSeismogramHandler().setShotInd(shotInd, shotIndIncr); // Simulation.cpp:421-423

            {
                std::lock_guard<std::mutex> lock(printMutex);
                HOST_PRINT("Start time stepping for shot number " << shotNumber << " (domain " << shotDomain << ", index " << shotInd + 1 << " of " << numshots << ")\n",
                           "\nTotal Number of time steps: " << tStepEnd << "\n")
            }
            start_t = now();
            wavefields->resetWavefields();
            // Simulation.cpp:441-456: `*wavefields *= compensation` after every step; the vector goes to the GPUs once and the
            // multiplication rides in the captured step graph
            if (config.getAndCatch("compensation", 0))
                solver->getGroup()->setStepScaling(modelLocal.getCompensation(DT, 1));
            double start_t2 = start_t;
            for (IndexType tStep = 0; tStep < tStepEnd; tStep++) {
                if ((tStep - 1) % 100 == 0)
                    start_t2 = now();
                solver->run(receivers, sources, modelLocal, *wavefields, *derivatives, tStep);
                if (tStep % 100 == 0 && tStep != 0 && verbose) {
                    solver->sync(); // the steps are enqueued asynchronously: time what has actually run
                    const double end_t2 = now();
                    std::lock_guard<std::mutex> lock(printMutex);
                    HOST_PRINT("", "Calculated " << tStep << " time steps in shot  " << shotNumber << " at t = " << end_t2 - globalStart_t << "\nLast 100 timesteps calculated in "
                                                 << end_t2 - start_t2 << " sec. - Estimated runtime (Simulation/total): " << (int)((tStepEnd / 100) * (end_t2 - start_t2)) << " / "
                                                 << (int)((tStepEnd / 100) * (end_t2 - start_t2) + tInit) << " sec.\n\n")
                }
                if (snapType > 0) {
                    const IndexType tFirst = Common::time2index(config.get<ValueType>("tFirstSnapshot"), DT), tLast = Common::time2index(config.get<ValueType>("tlastSnapshot"), DT),
                                    tInc = Common::time2index(config.get<ValueType>("tincSnapshot"), DT);
                    if (tStep >= tFirst && tStep <= tLast && tInc > 0 && (tStep - tFirst) % tInc == 0)
                        wavefields->write(snapType, config.get<std::string>("WavefieldFileName") + ".shot_" + std::to_string(shotNumber), tStep, config.get<IndexType>("FileFormat"));
                }
            }
            solver->sync();
            // Simulation.cpp:519: every value of the wavefields and of the seismograms must be finite
            SCAI_ASSERT_ERROR(wavefields->isFinite() && receivers.getSeismogramHandler().isFinite(), "Infinite or NaN value in seismogram or/and velocity wavefield!")
            solver->resetCPML();
            {
                std::lock_guard<std::mutex> lock(printMutex);
                HOST_PRINT("Finished time stepping for shot number: " << shotNumber << " in " << now() - start_t << " sec.\n")
            }
            // the SU trace headers refer to the source position when the shot has a single source (Simulation.cpp, Seismogram.cpp:939)
            receivers.getSeismogramHandler().setSourceCoordinate(sources.get1DCoordinates().size() == 1 ? sources.get1DCoordinates()[0] : 0);
            if (config.get<IndexType>("normalizeTraces") == 3) { // Simulation.cpp:520-525: automatic gain control; the gain function is written too
                receivers.getSeismogramHandler().setFrequencyAGC(config.get<ValueType>("CenterFrequencyCPML"));
                receivers.getSeismogramHandler().calcInverseAGC();
                receivers.getSeismogramHandler().write(5, config.get<std::string>("SeismogramFilename") + ".shot_" + std::to_string(shotNumber), &modelCoordinates);
            }
            receivers.getSeismogramHandler().normalize(config.get<IndexType>("normalizeTraces"));
            receivers.getSeismogramHandler().write(config.get<IndexType>("SeismogramFormat"), config.get<std::string>("SeismogramFilename") + ".shot_" + std::to_string(shotNumber),
                                                   &modelCoordinates);
            // Simulation.cpp:531-533: a supershot is split into the seismograms of its shots (files <SeismogramFilename>.shot_<n>.<type>), its marks are written
            receivers.decode(config, config.get<std::string>("SeismogramFilename"), shotNumber, survey.sourceSettingsEncode, 1);
            receivers.writeReceiverMark(config, shotNumber);
        }
        if (result) {
            result->sources = sources.getSeismogramHandler();
            result->receivers = receivers.getSeismogramHandler();
        }
    } catch (std::exception const &e) {
        *error = e.what();
    }
}

int main(int argc, const char *argv[])
{
    const double globalStart_t = now();
    if (argc != 2) {
        std::cout << "\n\nNo configuration file given!\n\n" << std::endl;
        return 2;
    }
    try {
        Configuration::Configuration config(argv[1]);
        verbose = config.getAndCatch("verbose", 0);
        std::string dimension = config.get<std::string>("dimension"), equationType = config.get<std::string>("equationType");
        std::transform(dimension.begin(), dimension.end(), dimension.begin(), ::tolower);
        std::transform(equationType.begin(), equationType.end(), equationType.begin(), ::tolower);
        Survey survey;
        survey.useStreamConfig = config.getAndCatch("useStreamConfig", false);
        Configuration::Configuration configBig;
        if (survey.useStreamConfig) { // Simulation.cpp:60-65: the big model the shots cut their sub-models from
            configBig.readFromFile(config.get<std::string>("streamConfigFilename"));
            survey.modelCoordinatesBig.init(configBig);
            SCAI_ASSERT_ERROR(config.get<IndexType>("useReceiversPerShot") != 0, "useStreamConfig = 1 is not possible when useReceiversPerShot = 0!") // Receivers.cpp:104-105
        }

        HOST_PRINT("\nWAVE-Simulation " << dimension << " " << equationType << " - LAMA-free host layer on " << ws_version() << "\n\n")
        if (verbose)
            config.print();

        IndexType nDevices = ws_device_count();
        SCAI_ASSERT_ERROR(nDevices > 0, "no CUDA device available (there is no CPU fallback)")
        if (const char *e = std::getenv("WS_NUM_GPUS"))
            nDevices = std::max<IndexType>(1, std::min<IndexType>(nDevices, std::atoi(e)));

        Acquisition::Coordinates<ValueType> modelCoordinates(config);
        if (config.getAndCatch("useVariableGrid", 0) != 0) { // Simulation.cpp:100-108
            CheckParameter::checkVariableGrid(config, modelCoordinates);
            for (IndexType layer = 0; layer < modelCoordinates.getNumLayers(); layer++)
                HOST_PRINT("\n Number of gridpoints in layer: " << layer << " = " << modelCoordinates.getNGridpoints(layer))
            const double numGridpointsRegular = (double)config.get<IndexType>("NX") * config.get<IndexType>("NY") * config.get<IndexType>("NZ");
            HOST_PRINT("\n Number of gripoints total: " << modelCoordinates.getNGridpoints())
            HOST_PRINT("\n Percentage of gridpoints of the underlying regular grid given by NX*NY*NZ: " << (float)(modelCoordinates.getNGridpoints() / numGridpointsRegular * 100) << "% \n\n")
        }
        if (config.getAndCatch("writeCoordinate", false)) // Simulation.cpp:151-154
            modelCoordinates.writeCoordinates(config.get<std::string>("coordinateFilename"), config.get<IndexType>("FileFormat"));
        const IndexType numRelaxationMechanisms = config.getAndCatch("numRelaxationMechanisms", 0);
        (void)numRelaxationMechanisms;

        /* memory estimation (Simulation.cpp:168-185) */
        {
            auto solver = ForwardSolver::Factory<ValueType>::Create(dimension, equationType);
            HOST_PRINT(" ========== " << dimension << " " << equationType << " Memory Estimation: ===========\n\n")
            HOST_PRINT(" Wavefields, model and boundary slabs in HBM: " << solver->estimateMemory(config, modelCoordinates) << " MB per shot domain "
                                                                        << (modelCoordinates.isVariable() ? "(variable grid: incl. the derivative operators in ELL form)" : "(derivatives are matrix-free: 0 MB)")
                                                                        << "\n\n")
        }

        /* acquisition geometry (Simulation.cpp:233-282): shot selection (shotIncr), per-shot cut-outs (useStreamConfig), source encoding */
        IndexType seedtime = config.getAndCatch("seedtime", (IndexType)time(nullptr)); // Simulation.cpp:46 takes the clock; the key makes a run repeatable
        Acquisition::Sources<ValueType> sources;
        {   // Receivers.cpp:128-138: the receiver mark matrix has one row per shot of the source file
            Acquisition::Sources<ValueType> all;
            all.getAcquisitionSettings(config, ValueType(0));
            std::vector<IndexType> nos;
            Acquisition::calcuniqueShotNo(nos, all.getSourceSettings());
            survey.numshotsAll = (IndexType)nos.size();
        }
        sources.getAcquisitionSettings(config, config.getAndCatch("shotIncr", ValueType(0)));
        survey.shotIndsIncr = sources.getShotIndsIncr();
        std::vector<Acquisition::sourceSettings<ValueType>> sourceSettings;
        if (survey.useStreamConfig) {
            std::vector<Acquisition::sourceSettings<ValueType>> sourceSettingsBig = sources.getSourceSettings();
            Acquisition::getCutCoord(config, survey.cutCoordinates, sourceSettingsBig, modelCoordinates, survey.modelCoordinatesBig);
            Acquisition::getSettingsPerShot(sourceSettings, sourceSettingsBig, survey.cutCoordinates, modelCoordinates, config.get<IndexType>("BoundaryWidth"));
            sources.setSourceSettings(sourceSettings); // for useSourceEncode
        } else
            sourceSettings = sources.getSourceSettings();
        CheckParameter::checkAcquisition<ValueType>(sourceSettings, modelCoordinates, "source");
        std::vector<IndexType> uniqueShotNos;
        Acquisition::calcuniqueShotNo(uniqueShotNos, sourceSettings);
        survey.useSourceEncode = config.getAndCatch("useSourceEncode", 0);
        const IndexType useRandomSource = config.getAndCatch("useRandomSource", 0);
        IndexType wantedDomains = std::max<IndexType>(1, config.getAndCatch("NumShotDomains", 1));
        {   // Partitioning.hpp:43-81 getShotDomain: 0 = NumShotDomains groups of equal size, 1 = the processors of a node form one domain (this
            // process drives the GPUs of ONE node: one domain over all of them), 2 = per-process environment variable DOMAIN (an MPI notion)
            const IndexType shotDomainDefinition = config.getAndCatch("ShotDomainDefinition", 0);
            SCAI_ASSERT_ERROR(shotDomainDefinition == 0 || shotDomainDefinition == 1,
                              "ShotDomainDefinition = " << shotDomainDefinition << ": the DOMAIN environment variable of an MPI process has no counterpart in the single-process driver")
            if (shotDomainDefinition == 1)
                wantedDomains = 1;
        }
        sources.calcSourceSettingsEncode(config, seedtime);
        IndexType numshots;
        if (survey.useSourceEncode == 0) {
            numshots = (IndexType)uniqueShotNos.size();
            survey.sourceSettings = sourceSettings;
            survey.uniqueShotNos = uniqueShotNos;
        } else { // the supershots take the place of the shots (Simulation.cpp:260-266)
            survey.sourceSettingsEncode = sources.getSourceSettingsEncode();
            survey.sourceSettings = survey.sourceSettingsEncode;
            Acquisition::calcuniqueShotNo(survey.uniqueShotNos, survey.sourceSettingsEncode);
            numshots = (IndexType)survey.uniqueShotNos.size();
            SCAI_ASSERT_ERROR(numshots <= wantedDomains, "more supershots than NumShotDomains")
        }
        // common-offset profiles (Simulation.cpp:356-361).  The receiver count the reference looks at before the shot loop is that of
        // <ReceiverFilename>.txt or, with receivers per shot, of the last shot Receivers::getModelPerShotSize samples (Receivers.cpp:627-662).
        // (With useSourceEncode the reference compares it with the shots per supershot and would write a profile of zeros: not done.)
        survey.copCondition = survey.useSourceEncode == 0 && uniqueShotNos.size() == sourceSettings.size() && uniqueShotNos.size() > 1;
        if (survey.copCondition) {
            survey.copSources = config.getAndCatch("writeSource", false);
            Acquisition::Receivers<ValueType> probe;
            const IndexType rps = config.get<IndexType>("useReceiversPerShot");
            std::vector<Acquisition::receiverSettings> probeSettings;
            const size_t shotSkip = std::max<size_t>(1, sourceSettings.size() / 10), last = (sourceSettings.size() - 1) / shotSkip * shotSkip;
            const IndexType lastNo = std::abs(sourceSettings[last].sourceNo);
            if (rps == 0 && !config.getAndCatch("initReceiverFromSU", false))
                Acquisition::readAllSettings(probeSettings, config.get<std::string>("ReceiverFilename") + ".txt");
            else if (rps == 1 && !config.getAndCatch("initReceiverFromSU", false))
                Acquisition::readAllSettings(probeSettings, config.get<std::string>("ReceiverFilename") + ".shot_" + std::to_string(lastNo) + ".txt");
            else if (rps == 2)
                probe.getAcquisitionSettings(config, probeSettings, lastNo, survey.numshotsAll, survey.shotIndsIncr, survey.sourceSettingsEncode);
            survey.copReceivers = probeSettings.size() == 1;
        }
        SCAI_ASSERT_ERROR(numshots >= wantedDomains || survey.useSourceEncode != 0, "numshots = " << numshots << ", numShotDomains = " << wantedDomains)
        // (Simulation.cpp:270-272 insists on numshots % NumShotDomains == 0; the block distribution below copes with a remainder)
        sources.writeShotIndsIncr(config, uniqueShotNos);
        sources.writeSourceFC(config); // Simulation.cpp:273
        sources.writeSourceEncode(config);
        if (survey.useStreamConfig)
            Acquisition::writeCutCoordToFile(config, configBig.get<std::string>("SourceFilename"), survey.cutCoordinates, uniqueShotNos, config.get<IndexType>("NX"));

        /* model (Simulation.cpp:284-295): with useStreamConfig the big model */
        double start_t = now();
        auto model = Modelparameter::Factory<ValueType>::Create(equationType);
        if (survey.useStreamConfig)
            model->init(configBig, survey.modelCoordinatesBig);
        else
            model->init(config, modelCoordinates);
        HOST_PRINT("", "Finished initializing model in " << now() - start_t << " sec.\n\n")

        /* shot domains: block distribution of the shots over min(NumShotDomains, GPUs) domains; the GPUs of a domain share
           one shot as y-slabs.  A slab should keep enough planes to hide the halo exchange behind its interior, so by
           default a domain uses at most NY / 64 GPUs (key GPUsPerShotDomain overrides, WS_NUM_GPUS limits the box). */
        IndexType numShotDomains = std::min(wantedDomains, std::min(numshots, nDevices));
        IndexType gpusPerDomain = std::max<IndexType>(1, nDevices / numShotDomains);
        {
            const IndexType NY = config.get<IndexType>("NY");
            const IndexType wanted = config.getAndCatch("GPUsPerShotDomain", 0);
            if (wanted > 0)
                gpusPerDomain = std::min(gpusPerDomain, wanted);
            else
                gpusPerDomain = std::min(gpusPerDomain, std::max<IndexType>(1, NY / 64));
        }
        HOST_PRINT(" " << numShotDomains << " shot domain(s) x " << gpusPerDomain << " GPU(s) per domain (y-slabs), " << nDevices << " GPU(s) visible\n\n")
        // which shots a domain works on.  Plain / encoded: its block of the shot list (the reference walks that block numshots /
        // NumShotDomains times, Simulation.cpp:345-362, and writes the same files every time: once is enough).  useRandomSource:
        // numshots / NumShotDomains passes, every pass draws NumShotDomains shots, one per configured domain (Simulation.cpp:339-369);
        // with fewer GPUs than domains a GPU group takes the shots of several configured domains.
        std::vector<std::vector<IndexType>> shotsOfDomain(numShotDomains);
        if (useRandomSource != 0) {
            std::vector<IndexType> shotHistory(numshots, 0);
            const IndexType numRand = numshots / wantedDomains, maxcount = 1;
            for (IndexType randInd = 0; randInd < numRand; randInd++) {
                sources.calcUniqueShotInds(config, shotHistory, maxcount, seedtime);
                auto const &inds = sources.getUniqueShotInds();
                for (IndexType k = 0; k < (IndexType)inds.size(); k++)
                    shotsOfDomain[k % numShotDomains].push_back(inds[k]);
            }
        } else {
            for (IndexType dom = 0; dom < numShotDomains; dom++) {
                const IndexType base = numshots / numShotDomains, rem = numshots % numShotDomains;
                const IndexType lb = dom * base + std::min(dom, rem), ub = lb + base + (dom < rem ? 1 : 0);
                for (IndexType k = lb; k < ub; k++)
                    shotsOfDomain[dom].push_back(k);
            }
        }
        std::vector<std::thread> threads;
        std::vector<std::string> errors(numShotDomains);
        std::vector<DomainResult> results(numShotDomains);
        for (IndexType dom = 0; dom < numShotDomains; dom++) {
            std::vector<IndexType> devices;
            for (IndexType r = 0; r < gpusPerDomain; r++)
                devices.push_back(dom * gpusPerDomain + r);
            if (numShotDomains == 1)
                runShotDomain(config, dom, devices, shotsOfDomain[dom], survey, model, modelCoordinates, globalStart_t, &errors[dom], &results[dom]);
            else
                threads.emplace_back(runShotDomain, std::cref(config), dom, devices, shotsOfDomain[dom], std::cref(survey), model, std::cref(modelCoordinates), globalStart_t, &errors[dom], &results[dom]);
        }
        for (auto &t : threads)
            t.join();
        for (auto const &e : errors)
            if (!e.empty())
                COMMON_THROWEXCEPTION(e)
        // Simulation.cpp:535-560: the common-offset profiles of the shot domains are summed and written by the first domain
        if (survey.copSources) {
            for (IndexType dom = 1; dom < numShotDomains; dom++)
                results[0].sources.sumShotDomain(results[dom].sources);
            results[0].sources.assignCOP();
            results[0].sources.write(config.get<IndexType>("SeismogramFormat"), config.get<std::string>("writeSourceFilename"), &modelCoordinates);
        }
        if (survey.copReceivers && results[0].receivers.getNumTracesTotal() == 1) {
            for (IndexType dom = 1; dom < numShotDomains; dom++)
                results[0].receivers.sumShotDomain(results[dom].receivers);
            results[0].receivers.assignCOP();
            results[0].receivers.write(config.get<IndexType>("SeismogramFormat"), config.get<std::string>("SeismogramFilename"), &modelCoordinates);
        }
        HOST_PRINT("\nTotal runtime of WAVE-Simulation: " << now() - globalStart_t << " sec.\nWAVE-Simulation finished!\n\n")
    } catch (std::exception const &e) {
        std::cerr << "\nERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

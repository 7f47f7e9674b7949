#include "Modelparameter.hpp"
#include <cmath>
#include "DeviceGroup.hpp"
#include "IO.hpp"
#include <algorithm>

using namespace KITGPI;

namespace
{
    struct ParDef {
        const char *name;   // C-ABI / reference getter name
        const char *suffix; // file suffix after ModelFilename
        const char *key;    // configuration key of the homogeneous model
    };
    const ParDef kVp = {"velocityP", "vp", "velocityP"}, kVs = {"velocityS", "vs", "velocityS"}, kRho = {"density", "density", "rho"},
                 kTauP = {"tauP", "tauP", "tauP"}, kTauS = {"tauS", "tauS", "tauS"}, kMu = {"magneticPermeability", "mur", "mur"},
                 kSig = {"electricConductivity", "sigma", "sigma"}, kEps = {"dielectricPermittivity", "epsilonr", "epsilonr"},
                 kTSig = {"tauElectricConductivity", "tauSigmar", "tauSigmar"}, kTEps = {"tauDielectricPermittivity", "tauEpsilon", "tauEpsilon"};

    std::vector<ParDef> parsOf(std::string const &t)
    {
        if (t == "acoustic") return {kVp, kRho};
        if (t == "elastic") return {kVp, kVs, kRho};
        if (t == "viscoelastic") return {kVp, kVs, kRho, kTauP, kTauS};
        if (t == "sh") return {kVs, kRho};
        if (t == "viscosh") return {kVs, kRho, kTauS};
        if (t == "tmem" || t == "emem") return {kMu, kSig, kEps};
        if (t == "viscotmem" || t == "viscoemem") return {kMu, kSig, kEps, kTSig, kTEps};
        COMMON_THROWEXCEPTION("Unkown type: " << t)
    }
}

template <typename ValueType> Modelparameter::Modelparameter<ValueType>::Modelparameter(std::string const &type) : equationType(type)
{
    seismic = Common::checkEquationType(type);
    parsOf(type);
}

template <typename ValueType>
void Modelparameter::Modelparameter<ValueType>::init(Configuration::Configuration const &config, Acquisition::Coordinates<ValueType> const &modelCoordinates)
{
    const IndexType modelRead = config.get<IndexType>("ModelRead");
    const bool variableGrid = config.getAndCatch("UseVariableGrid", 0) == 1;
    SCAI_ASSERT_ERROR(modelRead >= 0 && modelRead <= 2, "ModelRead=" << modelRead)
    SCAI_ASSERT_ERROR(modelRead != 2 || variableGrid, "Read variable model (ModelRead=2) not available if regular grid is chosen!") // Elastic.cpp:164
    // ModelRead = 1 on a variable grid (Elastic.cpp:174-184, Acoustic.cpp likewise): the file holds the model on the REGULAR grid NX x NY x NZ;
    // every point of the variable grid takes the value of the regular point at its coordinate.  ModelRead = 2 reads the variable grid itself.
    std::vector<IndexType> regularIndex;
    if (modelRead == 1 && variableGrid) {
        Acquisition::Coordinates<ValueType> regularCoordinates(config.get<IndexType>("NX"), config.get<IndexType>("NY"), config.get<IndexType>("NZ"), config.get<ValueType>("DH"));
        regularIndex.resize((size_t)modelCoordinates.getNGridpoints());
        for (IndexType i = 0; i < modelCoordinates.getNGridpoints(); i++)
            regularIndex[i] = regularCoordinates.coordinate2index(modelCoordinates.index2coordinate(i));
    }
    const bool visco = equationType.compare(0, 5, "visco") == 0;
    relaxationFrequency.clear();
    if (visco) {
        const IndexType L = config.get<IndexType>("numRelaxationMechanisms");
        SCAI_ASSERT_ERROR(L >= 1 && L <= 4, "numRelaxationMechanisms more than 4 is not available here!")
        const char *keys[4] = {"relaxationFrequency", "relaxationFrequency2", "relaxationFrequency3", "relaxationFrequency4"};
        for (IndexType l = 0; l < L; l++)
            relaxationFrequency.push_back(config.get<ValueType>(keys[l]));
    }
    const size_t N = (size_t)modelCoordinates.getNGridpoints();
    raw = std::make_shared<std::map<std::string, std::vector<ValueType>>>();
    for (auto const &p : parsOf(equationType)) {
        std::vector<ValueType> v(N);
        if (modelRead != 0) {
            if (regularIndex.empty())
                IO::readVector(v, config.get<std::string>("ModelFilename") + "." + p.suffix, config.get<IndexType>("FileFormat"));
            else {
                std::vector<ValueType> regular((size_t)config.get<IndexType>("NX") * config.get<IndexType>("NY") * config.get<IndexType>("NZ"));
                IO::readVector(regular, config.get<std::string>("ModelFilename") + "." + p.suffix, config.get<IndexType>("FileFormat"));
                for (size_t i = 0; i < N; i++)
                    v[i] = regular[(size_t)regularIndex[i]];
            }
        } else
            std::fill(v.begin(), v.end(), config.get<ValueType>(p.key));
        // EM inputs are relative: scale to SI (TMEM.cpp:262-292, ViscoTMEM.cpp:274-293)
        ValueType scale = 1;
        if (std::string(p.name) == "magneticPermeability") scale = MagneticPermeabilityVacuum;
        if (std::string(p.name) == "dielectricPermittivity") scale = DielectricPermittivityVacuum;
        if (std::string(p.name) == "tauElectricConductivity") scale = (ValueType)(1.0 / (2.0 * M_PI * config.get<ValueType>("CenterFrequencyCPML")));
        if (scale != 1)
            for (auto &x : v)
                x *= scale;
        (*raw)[p.name] = std::move(v);
    }
    centerFrequencyCPML = !seismic && visco ? config.get<ValueType>("CenterFrequencyCPML") : ValueType(0);
    if (!seismic && visco && modelRead != 0) {
        // ViscoTMEM.cpp:320-350, ViscoEMEM.cpp:290-306: the files hold the real EFFECTIVE permittivity and conductivity at the reference
        // frequency; the solver works on the static ones (ModelparameterEM.cpp:473-554 with calculateType = 2)
        ValueType aAverage, bAverage;
        relaxationAverages(aAverage, bAverage);
        auto &eps = (*raw)["dielectricPermittivity"], &sig = (*raw)["electricConductivity"];
        auto const &tauEps = (*raw)["tauDielectricPermittivity"], &tauSig = (*raw)["tauElectricConductivity"];
        for (size_t i = 0; i < N; i++) {
            ValueType b = bAverage * tauEps[i];
            ValueType a = ValueType(1) - aAverage * tauEps[i];
            a -= b * tauSig[i];
            ValueType epsStatic = (eps[i] - sig[i] * tauSig[i]) / a;
            if (epsStatic < DielectricPermittivityVacuum) // searchAndReplace(.., eps0, eps0, 1)
                epsStatic = DielectricPermittivityVacuum;
            ValueType sigStatic = sig[i] - b * epsStatic;
            if (sigStatic < 0)
                sigStatic = 0;
            eps[i] = epsStatic;
            sig[i] = sigStatic;
        }
    }
    dirtyFlag = true;
}

// a = mean_l (w tau_l)^2 / (1 + (w tau_l)^2), b = mean_l w^2 tau_l / (1 + (w tau_l)^2) at w = 2 pi CenterFrequencyCPML (ModelparameterEM.cpp:417-423, 448-454)
template <typename ValueType> void Modelparameter::Modelparameter<ValueType>::relaxationAverages(ValueType &aAverage, ValueType &bAverage) const
{
    aAverage = bAverage = 0;
    const ValueType w_ref = 2.0 * M_PI * centerFrequencyCPML;
    for (ValueType f : relaxationFrequency) {
        const ValueType relaxationTime = 1.0 / (2.0 * M_PI * f);
        aAverage += (w_ref * w_ref * relaxationTime * relaxationTime / (1 + w_ref * w_ref * relaxationTime * relaxationTime));
        bAverage += (w_ref * w_ref * relaxationTime / (1 + w_ref * w_ref * relaxationTime * relaxationTime));
    }
    aAverage /= relaxationFrequency.size();
    bAverage /= relaxationFrequency.size();
}

template <typename ValueType> void Modelparameter::Modelparameter<ValueType>::init(std::string const &name, std::vector<ValueType> const &values)
{
    if (raw.use_count() > 1)
        raw = std::make_shared<std::map<std::string, std::vector<ValueType>>>(*raw);
    (*raw)[name] = values;
    dirtyFlag = true;
}

template <typename ValueType> void Modelparameter::Modelparameter<ValueType>::write(std::string filename, IndexType fileFormat) const
{
    if (seismic) {
        for (auto const &p : parsOf(equationType))
            IO::writeVector(at(p.name), filename + "." + p.suffix, fileFormat);
        return;
    }
    // EM models are written as they are read: relative permeability / permittivity, tauSigmar relative to the reference relaxation time and,
    // for the visco types, the real effective permittivity and conductivity (TMEM.cpp:312-330, ViscoTMEM.cpp:382-402, ModelparameterEM.cpp:410-467)
    const bool visco = equationType.compare(0, 5, "visco") == 0;
    std::vector<ValueType> mu(at("magneticPermeability")), eps(at("dielectricPermittivity")), sig(at("electricConductivity"));
    for (auto &x : mu)
        x /= MagneticPermeabilityVacuum;
    if (visco) {
        ValueType aAverage, bAverage;
        relaxationAverages(aAverage, bAverage);
        auto const &tauEps = at("tauDielectricPermittivity"), &tauSig = at("tauElectricConductivity");
        for (size_t i = 0; i < eps.size(); i++) {
            const ValueType epsStatic = eps[i], sigStatic = sig[i];
            sig[i] = std::max(epsStatic * (bAverage * tauEps[i]) + sigStatic, ValueType(0));
            eps[i] = std::max(epsStatic * (ValueType(1) - aAverage * tauEps[i]) + sigStatic * tauSig[i], DielectricPermittivityVacuum);
        }
    }
    for (auto &x : eps)
        x /= DielectricPermittivityVacuum;
    IO::writeVector(mu, filename + ".mur", fileFormat);
    IO::writeVector(sig, filename + ".sigma", fileFormat);
    IO::writeVector(eps, filename + ".epsilonr", fileFormat);
    if (visco) {
        std::vector<ValueType> tauSig(at("tauElectricConductivity"));
        const ValueType relaxationTime_ref = 1.0 / (2.0 * M_PI * centerFrequencyCPML);
        for (auto &x : tauSig)
            x /= relaxationTime_ref;
        IO::writeVector(tauSig, filename + ".tauSigmar", fileFormat);
        IO::writeVector(at("tauDielectricPermittivity"), filename + ".tauEpsilon", fileFormat);
    }
}

template <typename ValueType> std::vector<ValueType> const &Modelparameter::Modelparameter<ValueType>::at(std::string const &name) const
{
    auto it = raw->find(name);
    if (it == raw->end())
        COMMON_THROWEXCEPTION("There is no " << name << " parameter in an " << equationType << " modelling")
    return it->second;
}

template <typename ValueType> std::vector<ValueType> Modelparameter::Modelparameter<ValueType>::getParameter(std::string const &name) const
{
    SCAI_ASSERT_ERROR(h, "The model is not bound to a forward solver yet (initForwardSolver)")
    return h->getMaterial(name);
}

template <typename ValueType> ValueType Modelparameter::Modelparameter<ValueType>::getMaxVelocity() const
{
    if (seismic) {
        auto const &v = (equationType == "sh" || equationType == "viscosh") ? getVelocityS() : getVelocityP();
        return *std::max_element(v.begin(), v.end());
    }
    auto const &e = getDielectricPermittivity(), &m = getMagneticPermeability();
    ValueType vmax = 0;
    for (size_t i = 0; i < e.size(); i++)
        vmax = std::max(vmax, (ValueType)(1.0 / std::sqrt((double)e[i] * m[i])));
    return vmax;
}

template <typename ValueType> ValueType Modelparameter::Modelparameter<ValueType>::getMinVelocity() const
{
    if (seismic) {
        auto const &v = equationType == "acoustic" ? getVelocityP() : getVelocityS();
        ValueType vmin = 3e8f;
        for (ValueType x : v)
            if (x > 0)
                vmin = std::min(vmin, x);
        return vmin;
    }
    auto const &e = getDielectricPermittivity(), &m = getMagneticPermeability();
    ValueType vmin = 3e8f;
    for (size_t i = 0; i < e.size(); i++)
        vmin = std::min(vmin, (ValueType)(1.0 / std::sqrt((double)e[i] * m[i])));
    return vmin;
}

template <typename ValueType> typename Modelparameter::Modelparameter<ValueType>::ModelparameterPtr Modelparameter::Factory<ValueType>::Create(std::string type)
{
    std::transform(type.begin(), type.end(), type.begin(), ::tolower);
    return std::make_shared<Modelparameter<ValueType>>(type); // throws "Unkown type" for anything else
}

template <typename ValueType>
void Modelparameter::Modelparameter<ValueType>::getModelPerShot(Modelparameter<ValueType> &modelPerShot, Acquisition::Coordinates<ValueType> const &mc,
                                                                Acquisition::Coordinates<ValueType> const &mcBig, Acquisition::coordinate3D const &cut) const
{
    SCAI_ASSERT_ERROR(modelPerShot.equationType == equationType, "model per shot of another equation type")
    SCAI_ASSERT_ERROR(cut.x >= 0 && cut.x + mc.getNX() <= mcBig.getNX() && cut.y + mc.getNY() <= mcBig.getNY() && cut.z + mc.getNZ() <= mcBig.getNZ(),
                      "the model per shot (cut at x = " << cut.x << ") does not lie inside the big model")
    const IndexType nx = mc.getNX(), ny = mc.getNY(), nz = mc.getNZ();
    auto out = std::make_shared<std::map<std::string, std::vector<ValueType>>>();
    for (auto const &kv : *raw) {
        std::vector<ValueType> v((size_t)nx * ny * nz);
        for (IndexType y = 0; y < ny; y++)
            for (IndexType z = 0; z < nz; z++) {
                const ValueType *src = &kv.second[(size_t)mcBig.coordinate2index(cut.x, y + cut.y, z + cut.z)];
                std::copy(src, src + nx, &v[(size_t)mc.coordinate2index(0, y, z)]);
            }
        (*out)[kv.first] = std::move(v);
    }
    modelPerShot.raw = out;
    modelPerShot.relaxationFrequency = relaxationFrequency;
    modelPerShot.centerFrequencyCPML = centerFrequencyCPML;
    modelPerShot.dirtyFlag = true;
}

template <typename ValueType> std::vector<ValueType> Modelparameter::Modelparameter<ValueType>::getCompensation(ValueType DT, IndexType tStep) const
{
    if (seismic)
        COMMON_THROWEXCEPTION("There is no compensation in an Seismic modelling")
    std::vector<ValueType> c = getElectricConductivity();
    std::vector<ValueType> const &eps = getDielectricPermittivity();
    const ValueType f = tStep * DT;
    for (size_t i = 0; i < c.size(); i++) {
        ValueType v = c[i] / eps[i];
        v *= f;
        c[i] = std::exp(v);
    }
    return c;
}

template class Modelparameter::Modelparameter<float>;
template class Modelparameter::Factory<float>;

#include "Acquisition.hpp"
#include "IO.hpp"
#include <algorithm>
#include <fstream>
#include <iomanip>

using namespace KITGPI;

const char *const Acquisition::SeismogramTypeString[4] = {"p", "vx", "vy", "vz"};
const char *const Acquisition::SeismogramTypeStringEM[4] = {"ez", "ex", "ey", "hz"};

// ---------------------------------------------------------------------------------------------------------------------
// acquisition files
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
    // tokens of the data lines of an acquisition file: `#` lines and blank lines are skipped (AcquisitionSettings.hpp:68-72)
    std::vector<std::vector<std::string>> tokenLines(std::string const &fileName, const char *what)
    {
        std::ifstream in(fileName, std::ios::binary);
        if (!in.is_open())
            COMMON_THROWEXCEPTION("Could not open " << what << " acquisition file " << fileName)
        std::vector<std::vector<std::string>> rows;
        std::string line;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] == '#' || std::all_of(line.begin(), line.end(), [](unsigned char c) { return std::isspace(c); }))
                continue;
            std::istringstream ss(line);
            std::vector<std::string> tok;
            std::string t;
            while (ss >> t)
                tok.push_back(t);
            rows.push_back(tok);
        }
        return rows;
    }
}

template <typename ValueType> void Acquisition::readAllSettings(std::vector<sourceSettings<ValueType>> &allSettings, std::string fileName)
{
    allSettings.clear();
    IndexType row = 0;
    for (auto const &tok : tokenLines(fileName, "source")) {
        if (tok.size() != 10)
            COMMON_THROWEXCEPTION("Wrong number of parameters in line of source acquisition file (" << fileName << ")")
        sourceSettings<ValueType> s;
        try {
            s.sourceNo = std::stoi(tok[0]);
            s.sourceCoords.x = std::stoi(tok[1]);
            s.sourceCoords.y = std::stoi(tok[2]);
            s.sourceCoords.z = std::stoi(tok[3]);
            s.sourceType = std::stoi(tok[4]);
            s.waveletType = std::stoi(tok[5]);
            s.waveletShape = std::stoi(tok[6]);
            s.fc = std::stof(tok[7]);
            s.amp = std::stof(tok[8]);
            s.tShift = std::stof(tok[9]);
        } catch (const std::exception &e) {
            COMMON_THROWEXCEPTION("Invalid argument while reading file " << fileName << " Message: " << e.what())
        }
        s.row = row++;
        allSettings.push_back(s);
    }
}

void Acquisition::readAllSettings(std::vector<receiverSettings> &allSettings, std::string const &fileName)
{
    allSettings.clear();
    for (auto const &tok : tokenLines(fileName, "receiver")) {
        if (tok.size() != 4)
            COMMON_THROWEXCEPTION("Wrong number of parameters in line of receiver acquisition file (" << fileName << ")")
        receiverSettings r;
        try {
            r.receiverCoords.x = std::stoi(tok[0]);
            r.receiverCoords.y = std::stoi(tok[1]);
            r.receiverCoords.z = std::stoi(tok[2]);
            r.receiverType = std::stoi(tok[3]);
        } catch (const std::exception &e) {
            COMMON_THROWEXCEPTION("Invalid argument while reading file " << fileName << " Message: " << e.what())
        }
        allSettings.push_back(r);
    }
}

template <typename ValueType> void Acquisition::calcuniqueShotNo(std::vector<IndexType> &uniqueShotNo, std::vector<sourceSettings<ValueType>> const &settings)
{
    uniqueShotNo.clear();
    for (auto const &s : settings) {
        const IndexType no = std::abs(s.sourceNo);
        if (std::find(uniqueShotNo.begin(), uniqueShotNo.end(), no) == uniqueShotNo.end())
            uniqueShotNo.push_back(no);
    }
}

template <typename ValueType>
void Acquisition::createSettingsForShot(std::vector<sourceSettings<ValueType>> &settings, std::vector<sourceSettings<ValueType>> const &allSettings, IndexType shotNumber)
{
    settings.clear();
    for (auto const &s : allSettings)
        if (std::abs(s.sourceNo) == shotNumber)
            settings.push_back(s);
}

// ---------------------------------------------------------------------------------------------------------------------
// wavelets: float arithmetic statement by statement as the LAMA vector expressions of Acquisition/SourceSignal/*.cpp
// ---------------------------------------------------------------------------------------------------------------------
void Acquisition::SourceSignal::calc(IndexType shape, std::vector<ValueType> &signal, IndexType NT, ValueType DT, ValueType FC, ValueType AMP, ValueType Tshift)
{
    typedef ValueType T;
    signal.assign(NT, T(0));
    auto window = [&](IndexType &i1, IndexType &i2) { // SinW / SinThree / IntgSinThree: one period after Tshift
        i1 = (IndexType)std::floor(Tshift / DT);
        i2 = i1 + (IndexType)std::floor(1.0 / FC / DT);
        if (i2 > NT)
            i2 = NT - 1;
    };
    switch (shape) {
    case 1: { // Ricker.cpp:41-52
        const T help = (T)(1.5 / FC + Tshift), w = (T)(M_PI * FC);
        for (IndexType k = 0; k < NT; k++) {
            T tau = (T(0) + (T)k * DT) - help;
            tau *= w;
            const T h2 = tau * tau;
            const T e = std::exp(T(-1.0) * h2);
            signal[k] = (AMP * (T(1.0) - T(2.0) * h2)) * e;
        }
        break;
    }
    case 2: { // SinW.cpp
        IndexType i1, i2;
        window(i1, i2);
        IndexType count = 0;
        for (IndexType i = i1; i <= i2 && i < NT; i++, count++) {
            if (i < 0)
                continue;
            double temp = 2.0 * count * DT * M_PI * FC;
            const T a = (T)std::sin(temp), b = (T)std::sin(2.0 * temp);
            signal[i] = AMP * (a - b / T(2.0));
        }
        break;
    }
    case 3: { // SinThree.cpp
        IndexType i1, i2;
        window(i1, i2);
        IndexType count = 0;
        for (IndexType i = i1; i <= i2 && i < NT; i++, count++) {
            if (i < 0)
                continue;
            double temp = std::sin(count * DT * M_PI * FC);
            signal[i] = AMP * (T)std::pow(temp, 3);
        }
        break;
    }
    case 4: { // FGaussian.cpp
        const T help0 = (T)(1.2 / FC + Tshift), w = (T)(M_PI * FC);
        for (IndexType k = 0; k < NT; k++) {
            T tau = ((T)k * DT - help0) * w;
            const T help = T(-2.0) * tau;
            tau = std::exp(T(-1.0) * tau * tau);
            signal[k] = AMP * help * tau;
        }
        break;
    }
    case 5: { // Spike.cpp
        const IndexType idx = (IndexType)std::floor(Tshift / DT);
        SCAI_ASSERT_ERROR(idx >= 0 && idx < NT, "Spike wavelet: tShift outside the time axis")
        signal[idx] = AMP * T(1.0);
        break;
    }
    case 6: { // IntgSinThree.cpp
        IndexType i1, i2;
        window(i1, i2);
        std::vector<T> help(NT, T(0)), zero(NT, T(0));
        IndexType count = 0;
        for (IndexType i = i1; i <= i2 && i < NT; i++, count++) {
            if (i < 0)
                continue;
            double temp = std::cos(count * DT * M_PI * FC);
            help[i] = (T)temp;
            zero[i] = (T)std::pow(temp, 3);
        }
        for (IndexType k = 0; k < NT; k++) {
            T z = T(0.25) * zero[k] - T(0.75) * help[k];
            z = T(0.5) + z;
            z = z / FC;
            z = z / (T)M_PI;
            z = z / T(0.75);
            signal[k] = AMP * z;
        }
        break;
    }
    case 7: { // Ricker_GprMax.cpp
        const T help0 = (T)(1 / FC + Tshift), zeta = (T)(2.0 * M_PI * M_PI * FC * FC);
        T h = T(1.0) / zeta;
        h = h / 4;
        h = std::exp(h);
        const T z2 = T(-2.0) * (zeta * h);
        for (IndexType k = 0; k < NT; k++) {
            const T tau = (T)k * DT - help0;
            T help = std::exp(T(-1.0) * (zeta * (tau * tau)));
            signal[k] = AMP * z2 * (help * tau);
        }
        break;
    }
    case 8: { // Berlage.cpp
        const T help0 = (T)(1.0 / FC + Tshift), alpha = 2 * FC;
        T maxNorm = 0;
        for (IndexType k = 0; k < NT; k++) {
            const T tau = (T)k * DT - help0;
            const T c = std::cos((T)(2 * M_PI * FC) * tau);
            T sgn = tau > 0 ? T(1) : (tau < 0 ? T(-1) : T(0));
            sgn += 1;
            const T heaviside = sgn > 0 ? T(1) : T(0);
            T v = c * std::pow(tau, T(2));
            v *= std::exp(tau * -alpha);
            v *= heaviside;
            signal[k] = v;
            maxNorm = std::max(maxNorm, std::abs(v));
        }
        for (auto &v : signal)
            v *= AMP / maxNorm;
        break;
    }
    case 9: { // Sin.cpp
        const T w = (T)(2.0 * M_PI * FC);
        for (IndexType k = 0; k < NT; k++)
            signal[k] = AMP * std::sin(((T)k * DT - Tshift) * w);
        break;
    }
    default: COMMON_THROWEXCEPTION("Unknown wavelet shape ")
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Common
// ---------------------------------------------------------------------------------------------------------------------
void Common::resampleRows(std::vector<ValueType> &data, IndexType numRows, IndexType numCols, ValueType resamplingCoeff, IndexType &numColsNew)
{
    numColsNew = IndexType(std::floor(ValueType(numCols - 1) / resamplingCoeff)) + 1;
    if (resamplingCoeff == ValueType(1)) {
        numColsNew = numCols;
        return; // identity matrix
    }
    std::vector<ValueType> out((size_t)numRows * numColsNew, ValueType(0));
    for (IndexType j = 0; j < numColsNew; j++) {
        const ValueType rel = j * resamplingCoeff;
        const IndexType left = (IndexType)std::floor(rel);
        const ValueType fr = std::fmod(rel, ValueType(1.0));
        for (IndexType r = 0; r < numRows; r++) {
            ValueType v = 0;
            if (left < numCols)
                v += (1 - fr) * data[(size_t)r * numCols + left];
            if (left + 1 < numCols)
                v += fr * data[(size_t)r * numCols + left + 1];
            out[(size_t)r * numColsNew + j] = v;
        }
    }
    data.swap(out);
}

bool Common::checkEquationType(std::string type)
{
    std::transform(type.begin(), type.end(), type.begin(), ::tolower);
    for (const char *s : {"acoustic", "elastic", "viscoelastic", "sh", "viscosh"})
        if (type == s)
            return true;
    for (const char *s : {"emem", "tmem", "viscoemem", "viscotmem"})
        if (type == s)
            return false;
    COMMON_THROWEXCEPTION("Unkown equation type " << type)
}

// ---------------------------------------------------------------------------------------------------------------------
// Seismogram / SeismogramHandler
// ---------------------------------------------------------------------------------------------------------------------
template <typename ValueType> void Acquisition::Seismogram<ValueType>::allocate(IndexType numTraces, IndexType NT)
{
    numSamples = NT;
    data.assign((size_t)numTraces * NT, ValueType(0));
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::normalizeTrace(IndexType normalizeTraces)
{
    if (data.empty() || normalizeTraces <= 0)
        return;
    for (IndexType i = 0; i < getNumTraces(); i++) {
        ValueType *row = &data[(size_t)i * numSamples];
        ValueType norm = 0;
        if (normalizeTraces == 1) {
            for (IndexType k = 0; k < numSamples; k++)
                norm = std::max(norm, std::abs(row[k]));
        } else {
            double s = 0;
            for (IndexType k = 0; k < numSamples; k++)
                s += (double)row[k] * row[k];
            norm = (ValueType)std::sqrt(s);
        }
        if (norm == 0)
            norm = 1;
        for (IndexType k = 0; k < numSamples; k++)
            row[k] /= norm;
    }
    if (normalizeTraces == 3 && useAGC) { // Seismogram.cpp:235-237
        SCAI_ASSERT_ERROR(inverseAGC.size() == data.size(), "calcInverseAGC first")
        for (size_t k = 0; k < data.size(); k++)
            data[k] *= inverseAGC[k];
        useAGC = false;
    } else if (normalizeTraces == 4) { // Seismogram.cpp:238-258
        std::vector<ValueType> envelope(data);
        Common::calcEnvelope(envelope, getNumTraces(), numSamples);
        const ValueType waterLevel = 1e-3;
        for (IndexType i = 0; i < getNumTraces(); i++) {
            ValueType *row = &data[(size_t)i * numSamples], *env = &envelope[(size_t)i * numSamples];
            ValueType envMax = 0;
            for (IndexType k = 0; k < numSamples; k++)
                envMax = std::max(envMax, std::abs(env[k]));
            const ValueType level = envMax != 0 ? waterLevel * envMax : waterLevel * waterLevel;
            for (IndexType k = 0; k < numSamples; k++)
                row[k] /= env[k] + level;
        }
    }
}

namespace
{
    // the window walk both AGC functions share (Seismogram.cpp:283-313, 341-381): `term(t)` is what sample t adds to the running sum
    template <typename ValueType, typename Term, typename Result>
    void agcWalk(IndexType NT, IndexType NAGC, ValueType waterLevel, Term term, Result result)
    {
        ValueType sumTemp = 0, NWIN = (ValueType)NAGC;
        for (IndexType t = NT - NAGC; t < NT; t++) {
            sumTemp += term(t);
            sumTemp += waterLevel;
        }
        for (IndexType t = NT - 1; t >= 0; t--) {
            if (t >= NT - NAGC) { // ramping on
                sumTemp += term(t - NAGC);
                sumTemp += waterLevel;
                NWIN += 1;
            } else if (t >= NAGC && t < NT - NAGC) { // full window
                sumTemp += term(t - NAGC);
                sumTemp -= term(t + NAGC);
            } else if (t < NAGC) { // ramping off
                sumTemp -= term(t + NAGC);
                sumTemp -= waterLevel;
                NWIN -= 1;
            }
            result(t, sumTemp / NWIN);
        }
    }
}

template <typename ValueType> std::vector<ValueType> Acquisition::Seismogram<ValueType>::getAGCSum()
{
    std::vector<ValueType> AGCSum;
    if (data.empty())
        return AGCSum;
    std::vector<ValueType> const original(data);
    normalizeTrace(2);
    std::vector<ValueType> const dataNorm(data);
    data = original;
    AGCSum.assign(data.size(), ValueType(0));
    const IndexType NT = numSamples;
    IndexType NAGC = (IndexType)std::round(1.0 / (frequencyAGC * DT));
    NAGC = std::min(NAGC, NT / 2);
    for (IndexType i = 0; i < getNumTraces(); i++) {
        ValueType const *row = &dataNorm[(size_t)i * NT];
        ValueType *out = &AGCSum[(size_t)i * NT];
        agcWalk<ValueType>(NT, NAGC, ValueType(0), [&](IndexType t) { return row[t]; }, [&](IndexType t, ValueType mean) { out[t] = mean; });
    }
    return AGCSum;
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::calcInverseAGC()
{
    if (data.empty())
        return;
    useAGC = true;
    std::vector<ValueType> const original(data);
    normalizeTrace(2);
    std::vector<ValueType> const dataNorm(data);
    data = original;
    inverseAGC.assign(data.size(), ValueType(0));
    const IndexType NT = numSamples;
    IndexType NAGC = (IndexType)std::round(1.0 / (frequencyAGC * DT));
    NAGC = std::min(NAGC, NT / 2);
    double all = 0;
    for (ValueType v : dataNorm)
        all += (double)v * v;
    for (IndexType i = 0; i < getNumTraces(); i++) {
        ValueType const *row = &dataNorm[(size_t)i * NT];
        ValueType *out = &inverseAGC[(size_t)i * NT];
        double s = 0;
        for (IndexType t = 0; t < NT; t++)
            s += (double)row[t] * row[t];
        ValueType waterLevel = (ValueType)std::sqrt(s);
        waterLevel *= waterLevel;
        waterLevel /= NT;
        if (waterLevel != 0)
            waterLevel *= ValueType(1e-3);
        else
            waterLevel = ValueType(1e-3) * (ValueType)std::sqrt(all) / NT / getNumTraces();
        agcWalk<ValueType>(NT, NAGC, waterLevel, [&](IndexType t) { return row[t] * row[t]; },
                           [&](IndexType t, ValueType meanSquare) { out[t] = meanSquare > 0 ? 1 / std::sqrt(meanSquare) : ValueType(0); });
    }
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::allocateCOP(IndexType numshots, IndexType NT)
{
    numshotsCOP = numshots;
    dataCOP.assign((size_t)numshots * NT, ValueType(0));
    inverseAGCCOP.assign((size_t)numshots * NT, ValueType(0));
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::assignCOP()
{
    SCAI_ASSERT_ERROR(numSamples > 0 && dataCOP.size() == (size_t)numshotsCOP * numSamples, "no common-offset profile of " << numSamples << " samples per trace")
    data = dataCOP;
    inverseAGC = inverseAGCCOP;
    coordinates1D.resize(numshotsCOP, coordinates1D.empty() ? 0 : coordinates1D[0]); // one trace per shot (SU headers: the geometry of the last shot)
    std::fill(dataCOP.begin(), dataCOP.end(), ValueType(0));
    std::fill(inverseAGCCOP.begin(), inverseAGCCOP.end(), ValueType(0));
}

template <typename ValueType> bool Acquisition::Seismogram<ValueType>::isFinite() const
{
    for (ValueType v : data)
        if (!std::isfinite(v))
            return false;
    return true;
}

template <typename ValueType>
void Acquisition::Seismogram<ValueType>::write(IndexType seismogramFormat, std::string const &filename, Coordinates<ValueType> const *modelCoordinates)
{
    if (data.empty())
        return;
    if (getNumTraces() == 1 && numshotsCOP > 1) { // Seismogram.cpp:84-96: the trace of this shot joins the common-offset profile
        SCAI_ASSERT_ERROR(shotInd >= 0 && shotInd < numshotsCOP && dataCOP.size() == (size_t)numshotsCOP * numSamples, "common-offset profile: shot index " << shotInd)
        std::vector<ValueType> const &from = seismogramFormat != 5 ? data : inverseAGC;
        std::vector<ValueType> &to = seismogramFormat != 5 ? dataCOP : inverseAGCCOP;
        SCAI_ASSERT_ERROR(from.size() == (size_t)numSamples, "no inverse AGC function to write")
        std::copy(from.begin(), from.end(), to.begin() + (size_t)shotInd * numSamples);
        return;
    }
    SCAI_ASSERT_ERROR(seismogramFormat == 1 || seismogramFormat == 2 || seismogramFormat == 4 || seismogramFormat == 5,
                      "SeismogramFormat " << seismogramFormat << " (3 = frv) is not available in the B200 host layer")
    std::string name = filename + "." + (isSeismic ? SeismogramTypeString[type] : SeismogramTypeStringEM[type]);
    IndexType seismoFormat = seismogramFormat;
    std::vector<ValueType> out(seismogramFormat != 5 ? data : inverseAGC);
    if (seismogramFormat == 5) { // Seismogram.cpp:112-116
        SCAI_ASSERT_ERROR(out.size() == data.size(), "no inverse AGC function to write")
        seismoFormat = 1;
        name += ".inverseAGC";
    }
    IndexType ns = numSamples;
    if (outputDT > 0 && DT > 0)
        Common::resampleRows(out, getNumTraces(), numSamples, outputDT / DT, ns); // Seismogram.cpp:610-618 setSeismoDT
    auto const put = [&](std::string const &file) {
        if (seismoFormat == 4) { // Seismogram.cpp:119-121
            SCAI_ASSERT_ERROR(modelCoordinates, "SeismogramFormat 4 (SU) needs the model coordinates for the trace headers")
            SUIO::writeSU(file, out, getNumTraces(), ns, coordinates1D, outputDT > 0 ? outputDT : DT, sourceCoordinate1D, *modelCoordinates);
        } else
            IO::writeMatrix(out, getNumTraces(), ns, file, seismoFormat);
    };
    put(name);
    if (outputInstantaneous != 0 && seismogramFormat != 5) { // Seismogram.cpp:127-145: a second file with the instantaneous property
        if (outputInstantaneous == 1) {
            Common::calcEnvelope(out, getNumTraces(), ns);
            name += ".envelope";
        } else if (outputInstantaneous == 2) {
            Common::calcInstantaneousPhase(out, getNumTraces(), ns, 2);
            name += ".instantaneousPhase";
        }
        put(name);
    }
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::read(IndexType seismogramFormat, std::string const &filename, bool readOriginal)
{
    const bool readSingleTrace = getNumTraces() == 1 && numshotsCOP > 1;
    std::string base = filename;
    if (readSingleTrace) {
        const size_t pos = filename.find(".shot");
        if (pos != std::string::npos)
            base = filename.substr(0, pos);
    }
    std::string name = base + "." + (isSeismic ? SeismogramTypeString[type] : SeismogramTypeStringEM[type]);
    IndexType seismoFormat = seismogramFormat;
    if (seismogramFormat == 5) {
        seismoFormat = 1;
        name += ".inverseAGC";
        useAGC = true;
    }
    std::vector<ValueType> in;
    IndexType r = 0, c = 0;
    if (seismoFormat == 4)
        SUIO::readDataSU(name, in, r, c); // Seismogram.cpp:180-193
    else
        IO::readMatrix(in, r, c, name, seismoFormat);
    if (readSingleTrace && r > 1) { // one row of the profile
        const IndexType row = readOriginal ? shotIndIncr : shotInd;
        SCAI_ASSERT_ERROR(row >= 0 && row < r, name << " holds " << r << " traces, trace " << row << " was asked for")
        in = std::vector<ValueType>(in.begin() + (size_t)row * c, in.begin() + (size_t)(row + 1) * c);
        r = 1;
    }
    if (!coordinates1D.empty())
        SCAI_ASSERT_ERROR(r == getNumTraces(), name << " holds " << r << " traces, the acquisition has " << getNumTraces())
    if (seismogramFormat == 5)
        inverseAGC = in;
    else {
        data = in;
        numSamples = c;
    }
}

template <typename ValueType> IndexType Acquisition::SeismogramHandler<ValueType>::getNumTracesTotal() const
{
    IndexType n = 0;
    for (auto const &s : seismo)
        n += s.getNumTraces();
    return n;
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setDT(ValueType dt)
{
    for (auto &s : seismo)
        s.setDT(dt);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setSeismoDT(ValueType dt)
{
    for (auto &s : seismo)
        s.setSeismoDT(dt);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::resetData()
{
    for (auto &s : seismo)
        s.resetData();
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::normalize(IndexType normalizeTraces)
{
    for (auto &s : seismo)
        s.normalizeTrace(normalizeTraces);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setFrequencyAGC(ValueType f)
{
    for (auto &s : seismo)
        s.setFrequencyAGC(f);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::calcInverseAGC()
{
    for (auto &s : seismo)
        s.calcInverseAGC();
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setInstantaneousTrace(IndexType instantaneousTraces)
{
    for (auto &s : seismo)
        s.setInstantaneousTrace(instantaneousTraces);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::allocateCOP(IndexType numshots, IndexType NT)
{
    for (auto &s : seismo)
        s.allocateCOP(numshots, NT);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setShotInd(IndexType shotIndTrue, IndexType shotIndIncr)
{
    for (auto &s : seismo)
        s.setShotInd(shotIndTrue, shotIndIncr);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::sumShotDomain(SeismogramHandler<ValueType> &other)
{
    for (IndexType t = 0; t < NUM_ELEMENTS_SEISMOGRAMTYPE; t++) {
        auto &a = seismo[t].getDataCOP(), &b = other.seismo[t].getDataCOP();
        auto &c = seismo[t].getInverseAGCCOP(), &d = other.seismo[t].getInverseAGCCOP();
        SCAI_ASSERT_ERROR(a.size() == b.size() && c.size() == d.size(), "common-offset profiles of the shot domains differ in size")
        for (size_t k = 0; k < a.size(); k++)
            a[k] += b[k];
        for (size_t k = 0; k < c.size(); k++)
            c[k] += d[k];
    }
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::assignCOP()
{
    for (auto &s : seismo)
        if (s.getNumTraces() > 0) {
            s.assignCOP();
            break;
        }
}
template <typename ValueType> bool Acquisition::SeismogramHandler<ValueType>::isFinite() const
{
    for (auto const &s : seismo)
        if (!s.isFinite())
            return false;
    return true;
}
template <typename ValueType>
void Acquisition::SeismogramHandler<ValueType>::write(IndexType seismogramFormat, std::string const &filename, Coordinates<ValueType> const *modelCoordinates)
{
    for (auto &s : seismo)
        s.write(seismogramFormat, filename, modelCoordinates);
}
template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::setSourceCoordinate(IndexType sourceCoord)
{
    for (auto &s : seismo)
        s.setSourceCoordinate(sourceCoord);
}

// ---------------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------------
template <typename ValueType>
template <typename Settings>
void Acquisition::AcquisitionGeometry<ValueType>::setAcquisition(std::vector<Settings> const &allSettings, Coordinates<ValueType> const &modelCoordinates, IndexType NT)
{
    SCAI_ASSERT_ERROR(!allSettings.empty(), "The acquisition is empty")
    coordinates1D.clear();
    types.clear();
    traceOfEntry.clear();
    IndexType count[NUM_ELEMENTS_SEISMOGRAMTYPE] = {0, 0, 0, 0};
    for (auto const &s : allSettings) {
        const IndexType type = s.getType();
        SCAI_ASSERT_ERROR(type >= 1 && type <= NUM_ELEMENTS_SEISMOGRAMTYPE, "Unkown Seismogram Type " << type)
        coordinates1D.push_back(modelCoordinates.coordinate2index(s.getCoords())); // throws outside the grid
        types.push_back(type);
        traceOfEntry.push_back(count[type - 1]++);
    }
    for (IndexType t = 0; t < NUM_ELEMENTS_SEISMOGRAMTYPE; t++) {
        auto &sg = seismograms.getSeismogram(t);
        sg.setTraceType(t, seismograms.getIsSeismic());
        sg.getCoordinates1D().clear();
        for (size_t k = 0; k < types.size(); k++)
            if (types[k] == t + 1)
                sg.getCoordinates1D().push_back(coordinates1D[k]);
        sg.allocate(count[t], NT);
    }
    version++;
}

// source geometry from the trace headers of <SourceFilename>.<component>.su (suHandler.cpp:14-27, 65-86): grid coordinates
// x = sx 10^scalco / DH + 0.5, y = sdepth 10^scalel / DH + 0.5, z = sy 10^scalco / DH + 0.5 (truncated), type = component of the
// file name, wavelet type 3 (signal from file).  The reference value-initialises the other fields (sourceSettingsVec.resize):
// shot number 0 and signal row 0 for every trace, i.e. all SU sources fire together as one shot with the first trace of
// SourceSignalFilename; components without a file are skipped here (the reference requires all four files).
template <typename ValueType> static void readSourceSettingsFromSU(std::vector<Acquisition::sourceSettings<ValueType>> &all, std::string const &filename, ValueType DH)
{
    all.clear();
    for (IndexType comp = 0; comp < Acquisition::NUM_ELEMENTS_SEISMOGRAMTYPE; comp++) {
        const std::string name = filename + "." + Acquisition::SeismogramTypeString[comp];
        const IndexType ntr = SUIO::numTracesSU(name);
        for (IndexType tr = 0; tr < ntr; tr++) {
            const double sco = std::pow(10.0, SUIO::readHeaderWordSU(name, tr, "scalco")), sel = std::pow(10.0, SUIO::readHeaderWordSU(name, tr, "scalel"));
            Acquisition::sourceSettings<ValueType> s{};
            s.sourceCoords.x = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "sx") * sco) / DH + 0.5);
            s.sourceCoords.y = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "sdepth") * sel) / DH + 0.5);
            s.sourceCoords.z = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "sy") * sco) / DH + 0.5);
            s.sourceType = comp + 1;
            s.waveletType = 3;
            all.push_back(s);
        }
    }
    SCAI_ASSERT_ERROR(!all.empty(), "No file with name: " << filename << ".'comp'.su could be read")
}

template <typename ValueType> void Acquisition::Seismogram<ValueType>::filterTraces(Filter::Filter<ValueType> const &freqFilter)
{
    if (getNumSamples() != 0 && getNumTraces() != 0)
        freqFilter.apply(data, getNumTraces(), getNumSamples());
}

template <typename ValueType> void Acquisition::SeismogramHandler<ValueType>::filter(Filter::Filter<ValueType> const &freqFilter)
{
    for (auto &s : seismo)
        s.filterTraces(freqFilter);
}

template <typename ValueType> void Acquisition::Sources<ValueType>::getAcquisitionSettings(Configuration::Configuration const &config, ValueType shotIncr)
{
    // Sources.cpp:502-514: with useStreamConfig the sources are those of the big model
    std::string filename = config.get<std::string>("SourceFilename");
    if (config.getAndCatch("useStreamConfig", false)) {
        Configuration::Configuration configBig(config.get<std::string>("streamConfigFilename"));
        filename = configBig.get<std::string>("SourceFilename");
    }
    std::vector<sourceSettings<ValueType>> allSettings;
    if (config.getAndCatch("initSourcesFromSU", false)) // Sources.cpp:511-512
        readSourceSettingsFromSU<ValueType>(allSettings, filename, config.get<ValueType>("DH"));
    else
        readAllSettings(allSettings, filename + ".txt");
    // Sources.cpp:516-559: one shot every shotIncr metres along the axis the shot line runs on
    const IndexType numshots = (IndexType)allSettings.size();
    const ValueType DH = config.get<ValueType>("DH");
    shotIndsIncr.clear();
    allSourceSettings.clear();
    if (numshots > 1 && shotIncr > DH) {
        std::vector<IndexType> sourceLength;
        const IndexType dx = std::abs(allSettings[0].sourceCoords.x - allSettings[1].sourceCoords.x), dy = std::abs(allSettings[0].sourceCoords.y - allSettings[1].sourceCoords.y);
        SCAI_ASSERT_ERROR(dx != dy, "shotIncr: the first two sources do not tell the direction of the shot line")
        for (IndexType k = 0; k < numshots; k++)
            sourceLength.push_back(dx > dy ? allSettings[k].sourceCoords.x : allSettings[k].sourceCoords.y);
        IndexType numshotsIncr = 0, shotIncrInd = sourceLength[0], shotIncrStep;
        allSourceSettings.push_back(allSettings[0]);
        shotIndsIncr.push_back(0);
        for (IndexType shotInd = 0; shotInd < numshots - 1; shotInd++) {
            if (shotIncrInd >= sourceLength[shotInd] && shotIncrInd <= sourceLength[shotInd + 1]) {
                if (shotIndsIncr[numshotsIncr] != shotInd && std::abs(shotIncrInd - sourceLength[shotInd]) < std::abs(shotIncrInd - sourceLength[shotInd + 1])) {
                    allSourceSettings.push_back(allSettings[shotInd]);
                    shotIndsIncr.push_back(shotInd);
                    numshotsIncr++;
                } else if (shotIndsIncr[numshotsIncr] != shotInd + 1 && std::abs(shotIncrInd - sourceLength[shotInd]) >= std::abs(shotIncrInd - sourceLength[shotInd + 1])) {
                    allSourceSettings.push_back(allSettings[shotInd + 1]);
                    shotIndsIncr.push_back(shotInd + 1);
                    numshotsIncr++;
                }
                shotIncrStep = (IndexType)std::round(shotIncr * (numshotsIncr + 1) / DH);
                shotIncrInd = sourceLength[0] + shotIncrStep;
                shotInd--; // the same interval may hold the next position too
            } else if (shotInd == numshots - 2 && shotIndsIncr[numshotsIncr] != shotInd + 1 &&
                       std::abs(sourceLength[shotInd + 1] - sourceLength[shotInd]) > std::abs(shotIncrInd - sourceLength[shotInd + 1])) {
                allSourceSettings.push_back(allSettings[shotInd + 1]);
                shotIndsIncr.push_back(shotInd + 1);
            }
        }
    } else {
        allSourceSettings = allSettings;
        for (IndexType k = 0; k < numshots; k++)
            shotIndsIncr.push_back(k);
    }
}

template <typename ValueType> void Acquisition::Sources<ValueType>::calcSourceSettingsEncode(Configuration::Configuration const &config, IndexType &seedtime, ValueType fc1, ValueType fc2)
{
    const IndexType numshotsIncr = (IndexType)shotIndsIncr.size();
    const IndexType useSourceEncode = config.getAndCatch("useSourceEncode", 0), useRandomSource = config.getAndCatch("useRandomSource", 0);
    SCAI_ASSERT_ERROR(useSourceEncode * useRandomSource == 0, "useSourceEncode and useRandomSource are not compatible!")
    SCAI_ASSERT_ERROR(useSourceEncode * config.getAndCatch("useSourceSignalTaper", 0) == 0, "useSourceEncode and useSourceSignalTaper are not compatible!")
    if (config.getAndCatch("useStreamConfig", 0) != 0)
        SCAI_ASSERT_ERROR(useSourceEncode == 0 || useSourceEncode == 3, "useSourceEncode must be 0 or 3 when useStreamConfig != 0!")
    // gradientDomain != 0 (Sources.cpp:584-612): the frequencies the shots are looked at, multiples of the FFT bin width between fc1 and fc2
    // (default 2 CenterFrequencyCPML).  Without encoding every shot gets the whole (widened) list; with it, every second bin from an even one
    // ("ensure steady-state wavefields"), one per shot of a supershot.
    const IndexType gradientDomain = config.getAndCatch("gradientDomain", 0);
    std::vector<ValueType> fc12;
    sourceFC.clear();
    if (gradientDomain != 0) {
        if (fc2 == 0)
            fc2 = config.get<ValueType>("CenterFrequencyCPML") * 2;
        const IndexType NT = static_cast<IndexType>((config.get<ValueType>("T") / config.get<ValueType>("DT")) + 0.5);
        const IndexType nFFT = Common::calcNextPowTwo<ValueType>(NT - 1);
        const ValueType df = 1 / (nFFT * config.get<ValueType>("DT"));
        IndexType fc1Ind = (IndexType)std::ceil(fc1 / df), fc2Ind = (IndexType)std::ceil(fc2 / df);
        if (useSourceEncode == 0) {
            fc1Ind = (IndexType)std::ceil((ValueType)fc1Ind / 2);
            fc2Ind *= 2;
            if (fc1Ind == 0)
                fc1Ind = 1;
            for (IndexType k = 0; k < fc2Ind - fc1Ind + 1; k++)
                fc12.push_back(fc1Ind * df + k * df);
            sourceFC.assign(numshotsIncr, fc12);
        } else {
            if (fc1Ind == 0)
                fc1Ind = 2;
            if (fc1Ind % 2 == 1)
                fc1Ind += 1;
            const IndexType nfc12 = (IndexType)std::ceil(ValueType(fc2Ind - fc1Ind + 1) / 2);
            for (IndexType k = 0; k < nfc12; k++)
                fc12.push_back(fc1Ind * df + k * 2 * df);
        }
    }
    sourceSettingsEncode.clear();
    if (useSourceEncode == 0)
        return;
    sourceSettingsEncode = allSourceSettings;
    const IndexType numShotDomains = config.get<IndexType>("NumShotDomains"); // the number of supershots
    const IndexType numShotPerSuperShot = (IndexType)std::ceil(ValueType(numshotsIncr) / numShotDomains);
    if (gradientDomain != 0)
        SCAI_ASSERT_ERROR((IndexType)fc12.size() >= numShotPerSuperShot, "The number of frequency is less than numShotPerSuperShot!")
    std::srand(seedtime);
    seedtime++;
    std::vector<IndexType> shotHistory(numShotDomains, 0);
    const IndexType base = (IndexType)(numShotDomains * 1e4) + 1;
    if (useSourceEncode == 1) { // randomly, with random polarity
        for (IndexType shotInd = 0; shotInd < numshotsIncr; shotInd++) {
            if (shotInd < numShotDomains) { // every supershot gets one shot first
                sourceSettingsEncode[shotInd].sourceNo = base + shotInd;
                shotHistory[shotInd]++;
            } else {
                IndexType sourceInd = std::rand() % numShotDomains;
                while (shotHistory[sourceInd] >= numShotPerSuperShot)
                    sourceInd = std::rand() % numShotDomains;
                shotHistory[sourceInd]++;
                sourceSettingsEncode[shotInd].sourceNo = base + sourceInd;
            }
        }
        for (IndexType shotInd = 0; shotInd < numshotsIncr; shotInd++) {
            const IndexType signAmp = (std::rand() % 2) > 0 ? 1 : -1;
            sourceSettingsEncode[shotInd].amp *= signAmp;
        }
    } else if (useSourceEncode == 2) { // sequentially to cover the global area
        for (IndexType shotInd = 0; shotInd < numshotsIncr; shotInd++)
            sourceSettingsEncode[shotInd].sourceNo = base + shotInd % numShotDomains;
    } else if (useSourceEncode == 3) { // sequentially to cover a local area
        for (IndexType shotInd = 0; shotInd < numshotsIncr; shotInd++)
            sourceSettingsEncode[shotInd].sourceNo = base + shotInd / numShotPerSuperShot;
    } else
        COMMON_THROWEXCEPTION("useSourceEncode must be 0, 1, 2 or 3")
    if (gradientDomain != 0) { // Sources.cpp:648-674: every shot of a supershot fires a sine of its own, randomly drawn frequency
        std::vector<IndexType> uniqueShotNosEncode;
        calcuniqueShotNo(uniqueShotNosEncode, sourceSettingsEncode);
        std::vector<ValueType> uniqueFC(numShotPerSuperShot, ValueType(0));
        const IndexType nfc12 = (IndexType)fc12.size();
        for (IndexType shotIndEncode = 0; shotIndEncode < numShotDomains && shotIndEncode < (IndexType)uniqueShotNosEncode.size(); shotIndEncode++) {
            std::vector<IndexType> fc12History(nfc12, 0);
            IndexType jf = 0;
            for (IndexType shotInd = 0; shotInd < numshotsIncr; shotInd++) {
                if (std::abs(sourceSettingsEncode[shotInd].sourceNo) != uniqueShotNosEncode[shotIndEncode])
                    continue;
                sourceSettingsEncode[shotInd].waveletType = 1;  // synthetic signal
                sourceSettingsEncode[shotInd].waveletShape = 9; // sin(t)
                IndexType fcInd = std::rand() % nfc12;
                while (fc12History[fcInd] > 0)
                    fcInd = std::rand() % nfc12;
                fc12History[fcInd]++;
                sourceSettingsEncode[shotInd].fc = fc12[fcInd];
                uniqueFC[jf++] = fc12[fcInd];
            }
            sourceFC.push_back(uniqueFC);
        }
    }
}

// <SourceFilename>.sourceFC.txt (Sources.cpp:801-846)
template <typename ValueType> void Acquisition::Sources<ValueType>::writeSourceFC(Configuration::Configuration const &config, IndexType stage, IndexType iteration) const
{
    const IndexType gradientDomain = config.getAndCatch("gradientDomain", 0), useSourceEncode = config.getAndCatch("useSourceEncode", 0);
    if (gradientDomain == 0 || sourceFC.empty())
        return;
    std::string filename = config.get<std::string>("SourceFilename");
    if (config.getAndCatch("useStreamConfig", false))
        filename = Configuration::Configuration(config.get<std::string>("streamConfigFilename")).get<std::string>("SourceFilename");
    if (stage != 0)
        filename += ".stage_" + std::to_string(stage) + ".It_" + std::to_string(iteration);
    std::ofstream out(filename + ".sourceFC.txt");
    out << "# Shot frequency used in the frequency domain gradient (gradientDomain = " << gradientDomain << ", useSourceEncode = " << useSourceEncode
        << ", numShotDomains = " << config.get<IndexType>("NumShotDomains") << ", NF = " << sourceFC[0].size() << ")\n# Shot number | frequency (Hz)\n";
    if (useSourceEncode == 0) {
        out << std::setw(13) << allSourceSettings[0].sourceNo;
        for (ValueType f : sourceFC[0])
            out << std::setw(14) << f;
        out << "\n";
        return;
    }
    std::vector<IndexType> uniqueShotNosEncode;
    calcuniqueShotNo(uniqueShotNosEncode, sourceSettingsEncode);
    SCAI_ASSERT_ERROR(sourceSettingsEncode.size() == shotIndsIncr.size(), "sourceSettingsEncode.size() != shotIndsIncr.size()")
    for (auto no : uniqueShotNosEncode) {
        out << std::setw(13) << no;
        for (auto const &e : sourceSettingsEncode)
            if (std::abs(e.sourceNo) == no)
                out << std::setw(14) << e.fc;
        out << "\n";
    }
}

void Acquisition::getRandomShotInds(std::vector<IndexType> &uniqueShotInds, std::vector<IndexType> &shotHistory, IndexType numshots, IndexType maxcount, IndexType useRandomSource,
                                    IndexType &seedtime)
{
    const IndexType numShotDomains = (IndexType)uniqueShotInds.size();
    if (useRandomSource == 1) {
        std::vector<IndexType> randomShotIndHistory(numShotDomains, 0);
        std::srand(seedtime);
        seedtime++;
        for (IndexType shotDomainInd = 0; shotDomainInd < numShotDomains; shotDomainInd++) {
            bool repeat = false;
            const IndexType randomShotInd = std::rand() % numshots;
            randomShotIndHistory[shotDomainInd] = randomShotInd;
            for (IndexType i = 0; i < shotDomainInd; i++)
                if (randomShotIndHistory[i] == randomShotInd) {
                    repeat = true;
                    break;
                }
            if (shotHistory[randomShotInd] >= maxcount || repeat)
                shotDomainInd--;
            else {
                uniqueShotInds[shotDomainInd] = randomShotInd;
                shotHistory[randomShotInd]++;
            }
        }
    } else if (useRandomSource == 2 || useRandomSource == 3) {
        IndexType sum = 0;
        for (auto c : shotHistory)
            sum += c;
        sum /= numShotDomains; // passes done so far
        const IndexType step = (IndexType)std::ceil(ValueType(numshots) / numShotDomains);
        for (IndexType shotDomainInd = 0; shotDomainInd < numShotDomains; shotDomainInd++) {
            IndexType ind = useRandomSource == 2 ? sum + shotDomainInd * step : sum * numShotDomains + shotDomainInd;
            ind %= numshots;
            uniqueShotInds[shotDomainInd] = ind;
            shotHistory[ind]++;
        }
    }
}

template <typename ValueType>
void Acquisition::Sources<ValueType>::calcUniqueShotInds(Configuration::Configuration const &config, std::vector<IndexType> &shotHistory, IndexType maxcount, IndexType &seedtime)
{
    const IndexType numshotsIncr = (IndexType)shotIndsIncr.size();
    const IndexType useSourceEncode = config.getAndCatch("useSourceEncode", 0), useRandomSource = config.getAndCatch("useRandomSource", 0);
    const IndexType numShotDomains = config.get<IndexType>("NumShotDomains");
    SCAI_ASSERT_ERROR(useSourceEncode * useRandomSource == 0, "useSourceEncode and useRandomSource are not compatible!")
    uniqueShotInds.clear();
    if (useRandomSource != 0) {
        uniqueShotInds.assign(numShotDomains, 0);
        getRandomShotInds(uniqueShotInds, shotHistory, numshotsIncr, maxcount, useRandomSource, seedtime);
    } else if (useSourceEncode != 0) {
        for (IndexType k = 0; k < numShotDomains; k++)
            uniqueShotInds.push_back(k);
    } else {
        for (IndexType k = 0; k < numshotsIncr; k++)
            uniqueShotInds.push_back(k);
    }
}

template <typename ValueType> void Acquisition::Sources<ValueType>::writeShotIndsIncr(Configuration::Configuration const &config, std::vector<IndexType> const &uniqueShotNos) const
{
    const ValueType shotIncr = config.getAndCatch("shotIncr", ValueType(0));
    if (!(shotIncr > config.get<ValueType>("DH")))
        return;
    SCAI_ASSERT_ERROR(shotIndsIncr.size() == uniqueShotNos.size(), "shotIndsIncr.size() != uniqueShotNos.size()")
    std::string filename = config.get<std::string>("SourceFilename");
    if (config.getAndCatch("useStreamConfig", false))
        filename = Configuration::Configuration(config.get<std::string>("streamConfigFilename")).get<std::string>("SourceFilename");
    std::ofstream out(filename + ".shotIncr.txt");
    out << "# Shot indices (shotIncr = " << shotIncr << " m, numshots = " << shotIndsIncr.size() << ")\n# Shot index | shot number\n";
    for (size_t k = 0; k < shotIndsIncr.size(); k++)
        out << std::setw(12) << shotIndsIncr[k] + 1 << std::setw(12) << uniqueShotNos[k] << "\n";
}

template <typename ValueType> void Acquisition::Sources<ValueType>::writeSourceEncode(Configuration::Configuration const &config) const
{
    const IndexType useSourceEncode = config.getAndCatch("useSourceEncode", 0);
    if (useSourceEncode == 0)
        return;
    std::vector<IndexType> uniqueShotNosEncode;
    calcuniqueShotNo(uniqueShotNosEncode, sourceSettingsEncode);
    SCAI_ASSERT_ERROR(sourceSettingsEncode.size() == shotIndsIncr.size(), "sourceSettingsEncode.size() != shotIndsIncr.size()")
    std::string filename = config.get<std::string>("SourceFilename");
    if (config.getAndCatch("useStreamConfig", false))
        filename = Configuration::Configuration(config.get<std::string>("streamConfigFilename")).get<std::string>("SourceFilename");
    std::ofstream out(filename + ".encode.txt");
    out << "# Shot indices used in source encode (useSourceEncode = " << useSourceEncode << ", numShotDomains = " << config.get<IndexType>("NumShotDomains")
        << ", numshots = " << shotIndsIncr.size() << ")\n# Shot number | shot index (selected)\n";
    for (auto no : uniqueShotNosEncode) {
        out << std::setw(13) << no;
        for (size_t k = 0; k < sourceSettingsEncode.size(); k++)
            if (std::abs(sourceSettingsEncode[k].sourceNo) == no)
                out << std::setw(5) << k + 1;
        out << "\n";
    }
}

void Acquisition::getuniqueShotInd(IndexType &shotInd, std::vector<IndexType> const &uniqueShotNos, IndexType shotNumber)
{
    for (size_t i = 0; i < uniqueShotNos.size(); i++)
        if (uniqueShotNos[i] == shotNumber) {
            shotInd = (IndexType)i;
            break;
        }
}
template <typename ValueType> void Acquisition::getuniqueShotInd(IndexType &shotInd, std::vector<sourceSettings<ValueType>> const &enc, IndexType shotNumber)
{
    for (size_t i = 0; i < enc.size(); i++)
        if (std::abs(enc[i].sourceNo) == shotNumber) {
            shotInd = (IndexType)i;
            break;
        }
}

template <typename ValueType>
void Acquisition::getCutCoord(Configuration::Configuration const &config, std::vector<coordinate3D> &cutCoordinates, std::vector<sourceSettings<ValueType>> const &big,
                              Coordinates<ValueType> const &modelCoordinates, Coordinates<ValueType> const &modelCoordinatesBig)
{
    cutCoordinates.clear();
    std::vector<IndexType> uniqueShotNos;
    calcuniqueShotNo(uniqueShotNos, big);
    SCAI_ASSERT_ERROR(big.size() == uniqueShotNos.size(), "sourceSettingsBig.size() != uniqueShotNos.size()")
    const IndexType numshotsIncr = (IndexType)big.size();
    SCAI_ASSERT_ERROR(modelCoordinates.getDH() == modelCoordinatesBig.getDH(), "DH != DHBig")
    const IndexType useSourceEncode = config.getAndCatch("useSourceEncode", 0);
    const IndexType numShotDomains = config.get<IndexType>("NumShotDomains");
    const IndexType numShotPerSuperShot = (IndexType)std::ceil(ValueType(numshotsIncr) / numShotDomains);
    IndexType minX = big[0].sourceCoords.x;
    for (auto const &s : big)
        minX = std::min(minX, s.sourceCoords.x);
    minX -= (IndexType)std::round((modelCoordinates.getX0() - modelCoordinatesBig.getX0()) / modelCoordinates.getDH());
    IndexType sourceCoordX = 0;
    for (IndexType i = 0; i < numshotsIncr; i++) {
        if (big[i].sourceNo >= 0 && (useSourceEncode == 0 || (useSourceEncode == 3 && i % numShotPerSuperShot == 0)))
            sourceCoordX = big[i].sourceCoords.x; // a negative sourceNo keeps the cut of the shot before it
        SCAI_ASSERT_ERROR(sourceCoordX != 0, "sourceCoordX cannot be 0 when sourceSettingsBig[i].sourceNo < 0")
        coordinate3D c;
        c.x = sourceCoordX - minX > 0 ? sourceCoordX - minX : 0;
        c.y = 0;
        c.z = 0;
        cutCoordinates.push_back(c);
    }
}

namespace
{
    template <typename ValueType> void moveIntoSubModel(Acquisition::coordinate3D &c, Acquisition::coordinate3D const &cut, Acquisition::Coordinates<ValueType> const &mc, IndexType W, const char *what, size_t i)
    {
        c.x -= cut.x;
        SCAI_ASSERT_ERROR(c.x >= W && c.x < mc.getNX() - W, "settings[" << i << "]." << what << ".x = " << c.x)
        c.y -= cut.y;
        SCAI_ASSERT_ERROR(c.y >= W && c.y < mc.getNY() - W, "settings[" << i << "]." << what << ".y = " << c.y)
        c.z -= cut.z;
        if (mc.getNZ() > W * 2)
            SCAI_ASSERT_ERROR(c.z >= W && c.z < mc.getNZ() - W, "settings[" << i << "]." << what << ".z = " << c.z)
    }
}

template <typename ValueType>
void Acquisition::getSettingsPerShot(std::vector<sourceSettings<ValueType>> &settings, std::vector<sourceSettings<ValueType>> const &allSettings, std::vector<coordinate3D> const &cutCoordinates,
                                     Coordinates<ValueType> const &modelCoordinates, IndexType BoundaryWidth)
{
    settings = allSettings;
    for (size_t i = 0; i < settings.size(); i++)
        moveIntoSubModel(settings[i].sourceCoords, cutCoordinates[i], modelCoordinates, BoundaryWidth, "sourceCoords", i);
}

template <typename ValueType>
void Acquisition::getSettingsPerShot(std::vector<receiverSettings> &settings, std::vector<receiverSettings> const &allSettings, coordinate3D const &cutCoordinate,
                                     Coordinates<ValueType> const &modelCoordinates, IndexType BoundaryWidth)
{
    settings = allSettings;
    for (size_t i = 0; i < settings.size(); i++)
        moveIntoSubModel(settings[i].receiverCoords, cutCoordinate, modelCoordinates, BoundaryWidth, "receiverCoords", i);
}

void Acquisition::writeCutCoordToFile(Configuration::Configuration const &config, std::string const &sourceFilename, std::vector<coordinate3D> const &cutCoordinates,
                                      std::vector<IndexType> const &uniqueShotNos, IndexType NXPerShot)
{
    if (!config.getAndCatch("useStreamConfig", false))
        return;
    std::ofstream out(sourceFilename + ".cut.txt");
    out << "# Coordinate for cutting model per shot (the first row is the size of modelPerShot)\n# ShotNumber | index_x | index_y | index_z\n";
    out << std::setw(12) << (long long)NXPerShot * config.get<IndexType>("NY") * config.get<IndexType>("NZ") << std::setw(10) << NXPerShot << std::setw(10) << config.get<IndexType>("NY")
        << std::setw(10) << config.get<IndexType>("NZ") << "\n";
    for (size_t i = 0; i < uniqueShotNos.size(); i++)
        out << std::setw(12) << uniqueShotNos[i] << std::setw(10) << cutCoordinates[i].x << std::setw(10) << cutCoordinates[i].y << std::setw(10) << cutCoordinates[i].z << "\n";
}

template <typename ValueType>
void Acquisition::Sources<ValueType>::init(std::vector<sourceSettings<ValueType>> const &shotSettings, Configuration::Configuration const &config,
                                           Coordinates<ValueType> const &modelCoordinates)
{
    this->seismograms.setIsSeismic(Common::checkEquationType(config.get<std::string>("equationType")));
    const ValueType DT = config.get<ValueType>("DT");
    const IndexType NT = static_cast<IndexType>((config.get<ValueType>("T") / DT) + 0.5);
    this->setAcquisition(shotSettings, modelCoordinates, NT);
    this->seismograms.setDT(DT);
    this->seismograms.setSeismoDT(DT);
    // signals (Sources.cpp:99-145): 1 = synthetic, 2 = one row of <SourceSignalFilename> for all, 3 = one row per source
    std::vector<ValueType> fileSignals;
    IndexType fileRows = 0, fileCols = 0;
    std::vector<ValueType> sig;
    bool flag2 = false, flag3 = false;
    for (size_t k = 0; k < shotSettings.size(); k++) {
        auto const &s = shotSettings[k];
        auto &sg = this->seismograms.getSeismogram(s.sourceType - 1);
        ValueType *row = &sg.getData()[(size_t)this->traceOfEntry[k] * NT];
        if (s.waveletType == 1) {
            SourceSignal::calc(s.waveletShape, sig, NT, DT, s.fc, s.amp, s.tShift);
            std::copy(sig.begin(), sig.end(), row);
        } else if (s.waveletType == 2 || s.waveletType == 3) {
            (s.waveletType == 2 ? flag2 : flag3) = true;
            SCAI_ASSERT_ERROR(!(flag2 && flag3), "Combination of wavelet type 2 and 3 not supported")
            if (fileSignals.empty()) {
                if (config.getAndCatch("initSourcesFromSU", false)) { // Sources.cpp:235-245: traces of the SU file SourceSignalFilename
                    std::string name = config.get<std::string>("SourceSignalFilename");
                    if (name.size() > 3 && name.compare(name.size() - 3, 3, ".su") == 0)
                        name.erase(name.size() - 3);
                    SUIO::readDataSU(name, fileSignals, fileRows, fileCols);
                } else
                    IO::readMatrix(fileSignals, fileRows, fileCols, config.get<std::string>("SourceSignalFilename"), config.get<IndexType>("SeismogramFormat"));
            }
            const IndexType r = s.waveletType == 2 ? 0 : s.row;
            SCAI_ASSERT_ERROR(r < fileRows && fileCols == NT, "source signal file must hold one row of " << NT << " samples per source")
            std::copy(fileSignals.begin() + (size_t)r * fileCols, fileSignals.begin() + (size_t)(r + 1) * fileCols, row);
        } else
            COMMON_THROWEXCEPTION("Unknown wavelet type ")
    }
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::init(std::vector<receiverSettings> const &allSettings, Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates)
{
    this->seismograms.setIsSeismic(Common::checkEquationType(config.get<std::string>("equationType")));
    const IndexType NT = static_cast<IndexType>((config.get<ValueType>("T") / config.get<ValueType>("DT")) + 0.5);
    this->setAcquisition(allSettings, modelCoordinates, NT);
    this->seismograms.setDT(config.get<ValueType>("DT"));
    this->seismograms.setSeismoDT(config.get<ValueType>("seismoDT"));
    this->seismograms.setInstantaneousTrace(config.getAndCatch("instantaneousTraces", 0)); // Receivers.cpp:31
}

// receiver geometry from the trace headers of <filename>.<component>.su (suHandler.cpp:33-52,88-107): grid coordinates
// x = gx 10^scalco / DH, y = gelev 10^scalel / DH, z = gy 10^scalco / DH (truncated), type = component of the file name
template <typename ValueType> static void readReceiverSettingsFromSU(std::vector<Acquisition::receiverSettings> &all, std::string const &filename, ValueType DH)
{
    all.clear();
    IndexType missing = 0;
    for (IndexType comp = 0; comp < Acquisition::NUM_ELEMENTS_SEISMOGRAMTYPE; comp++) {
        const std::string name = filename + "." + Acquisition::SeismogramTypeString[comp];
        const IndexType ntr = SUIO::numTracesSU(name);
        if (ntr == 0) {
            missing++;
            continue;
        }
        for (IndexType tr = 0; tr < ntr; tr++) {
            const double sco = std::pow(10.0, SUIO::readHeaderWordSU(name, tr, "scalco")), sel = std::pow(10.0, SUIO::readHeaderWordSU(name, tr, "scalel"));
            Acquisition::receiverSettings r;
            r.receiverCoords.x = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "gx") * sco) / DH);
            r.receiverCoords.y = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "gelev") * sel) / DH);
            r.receiverCoords.z = static_cast<IndexType>((ValueType)(SUIO::readHeaderWordSU(name, tr, "gy") * sco) / DH);
            r.receiverType = comp + 1;
            all.push_back(r);
        }
    }
    SCAI_ASSERT_ERROR(missing < Acquisition::NUM_ELEMENTS_SEISMOGRAMTYPE, "No file with name: " << filename << ".'comp'.su could be read")
}

template <typename ValueType> void Acquisition::Receivers<ValueType>::init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates)
{
    if (config.getAndCatch("initReceiverFromSU", false)) { // Receivers.cpp: acquisition from the SU trace headers
        std::vector<receiverSettings> all;
        readReceiverSettingsFromSU<ValueType>(all, config.get<std::string>("ReceiverFilename"), modelCoordinates.getDH());
        init(all, config, modelCoordinates);
        return;
    }
    std::vector<receiverSettings> all;
    readAllSettings(all, config.get<std::string>("ReceiverFilename") + ".txt");
    init(all, config, modelCoordinates);
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates, IndexType shotNumber)
{
    // Receivers.cpp:229-246: <ReceiverFilename>.shot_<n> as a text file or, with initReceiverFromSU, as SU files per component
    std::vector<receiverSettings> all;
    const std::string name = config.get<std::string>("ReceiverFilename") + ".shot_" + std::to_string(shotNumber);
    if (config.getAndCatch("initReceiverFromSU", false))
        readReceiverSettingsFromSU<ValueType>(all, name, config.get<ValueType>("DH"));
    else
        readAllSettings(all, name + ".txt");
    init(all, config, modelCoordinates);
}

// Receivers.cpp:219-227 + 262-276: the receivers of the whole survey (txt or SU headers) and the mark matrix `<ReceiverFilename>.mark.mtx`
template <typename ValueType>
void Acquisition::Receivers<ValueType>::readMarkMatrix(Configuration::Configuration const &config, std::vector<receiverSettings> &all, std::vector<ValueType> &mark, IndexType numshots) const
{
    all.clear();
    if (config.getAndCatch("initReceiverFromSU", false))
        readReceiverSettingsFromSU<ValueType>(all, config.get<std::string>("ReceiverFilename"), config.get<ValueType>("DH"));
    else
        readAllSettings(all, config.get<std::string>("ReceiverFilename") + ".txt");
    std::string markName = config.get<std::string>("ReceiverFilename") + ".mark";
    if (config.getAndCatch("useStreamConfig", false)) {
        Configuration::Configuration configBig(config.get<std::string>("streamConfigFilename"));
        markName = configBig.get<std::string>("ReceiverFilename") + ".mark";
    }
    IndexType rows = 0, cols = 0;
    IO::readMatrix(mark, rows, cols, markName, 1);
    SCAI_ASSERT_ERROR(rows == numshots && cols == (IndexType)all.size() + 1,
                      "the receiver mark matrix must have numshots = " << numshots << " rows and numrecs + 1 = " << all.size() + 1 << " columns")
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates, IndexType shotNumber, IndexType numshots,
                                             std::vector<IndexType> const &shotIndsIncr, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode)
{
    std::vector<receiverSettings> active;
    getAcquisitionSettings(config, active, shotNumber, numshots, shotIndsIncr, sourceSettingsEncode);
    init(active, config, modelCoordinates);
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::getAcquisitionSettings(Configuration::Configuration const &config, std::vector<receiverSettings> &active, IndexType shotNumber, IndexType numshots,
                                                               std::vector<IndexType> const &shotIndsIncr, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode)
{
    std::vector<receiverSettings> all;
    std::vector<ValueType> mark;
    readMarkMatrix(config, all, mark, numshots);
    const IndexType numrecs = (IndexType)all.size(), cols = numrecs + 1;
    receiverMarkVector.assign(cols, ValueType(0));
    if (sourceSettingsEncode.empty()) { // a plain shot: its own row
        bool found = false;
        for (IndexType row : shotIndsIncr) {
            SCAI_ASSERT_ERROR(row >= 0 && row < numshots, "shot index outside the mark matrix")
            if ((IndexType)mark[(size_t)row * cols] == shotNumber) {
                std::copy(mark.begin() + (size_t)row * cols, mark.begin() + (size_t)(row + 1) * cols, receiverMarkVector.begin());
                found = true;
                break;
            }
        }
        SCAI_ASSERT_ERROR(found, "receiverMarkVector[0] != shotNumber")
    } else { // a supershot: the union of the rows of its shots
        SCAI_ASSERT_ERROR(sourceSettingsEncode.size() == shotIndsIncr.size(), "sourceSettingsEncode.size() != shotIndsIncr.size()")
        for (size_t k = 0; k < shotIndsIncr.size(); k++)
            if (std::abs(sourceSettingsEncode[k].sourceNo) == shotNumber)
                for (IndexType c = 0; c < cols; c++)
                    receiverMarkVector[c] += mark[(size_t)shotIndsIncr[k] * cols + c];
    }
    for (auto &v : receiverMarkVector) // UnaryOp::SIGN
        v = v > 0 ? ValueType(1) : (v < 0 ? ValueType(-1) : ValueType(0));
    receiverMarkVector[0] = (ValueType)shotNumber;
    active.clear();
    for (IndexType r = 0; r < numrecs; r++)
        if (receiverMarkVector[r + 1] != 0)
            active.push_back(all[r]);
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::writeReceiverMark(Configuration::Configuration const &config, IndexType shotNumber, IndexType stage, IndexType iteration) const
{
    if (config.getAndCatch("useSourceEncode", 0) == 0)
        return;
    std::string name = config.get<std::string>("ReceiverFilename");
    if (config.getAndCatch("useStreamConfig", false)) {
        Configuration::Configuration configBig(config.get<std::string>("streamConfigFilename"));
        name = configBig.get<std::string>("ReceiverFilename");
    }
    if (stage != 0)
        name += ".stage_" + std::to_string(stage) + ".It_" + std::to_string(iteration);
    IO::writeVector(receiverMarkVector, name + ".shot_" + std::to_string(shotNumber) + ".mark", 1);
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::decode(Configuration::Configuration const &config, std::string const &filename, IndexType shotNumber,
                                               std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode, IndexType encodeType)
{
    encode(config, filename, shotNumber, sourceSettingsEncode, encodeType + 2); // Receivers.cpp:569-573
}

template <typename ValueType>
void Acquisition::Receivers<ValueType>::encode(Configuration::Configuration const &config, std::string const &filename, IndexType shotNumber,
                                               std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode, IndexType encodeType)
{
    if (config.getAndCatch("useSourceEncode", 0) == 0)
        return;
    SCAI_ASSERT_ERROR(config.get<IndexType>("useReceiversPerShot") == 2, "useReceiversPerShot != 2")
    const bool toSuper = encodeType == 0 || encodeType == 1;
    // the shots of the source file, and the rows the shotIncr selection keeps
    Sources<ValueType> allSources;
    allSources.getAcquisitionSettings(config, ValueType(0));
    std::vector<IndexType> uniqueShotNos;
    calcuniqueShotNo(uniqueShotNos, allSources.getSourceSettings());
    const IndexType numshots = (IndexType)uniqueShotNos.size();
    allSources.getAcquisitionSettings(config, config.getAndCatch("shotIncr", ValueType(0)));
    std::vector<IndexType> const shotIndsIncr = allSources.getShotIndsIncr();
    SCAI_ASSERT_ERROR(sourceSettingsEncode.size() == shotIndsIncr.size(), "sourceSettingsEncode.size() != shotIndsIncr.size()")

    std::vector<receiverSettings> all;
    std::vector<ValueType> mark;
    readMarkMatrix(config, all, mark, numshots);
    const IndexType numrecs = (IndexType)all.size(), cols = numrecs + 1;
    SCAI_ASSERT_ERROR((IndexType)receiverMarkVector.size() == cols, "the receivers of the supershot were not set up from the mark matrix")
    std::vector<IndexType> receiverTypes; // types among the receivers of the supershot, in order of appearance
    for (IndexType r = 0; r < numrecs; r++)
        if (receiverMarkVector[r + 1] != 0 && std::find(receiverTypes.begin(), receiverTypes.end(), all[r].receiverType) == receiverTypes.end())
            receiverTypes.push_back(all[r].receiverType);

    const IndexType seismoFormat = config.get<IndexType>("SeismogramFormat");
    const IndexType gradientDomain = config.getAndCatch("gradientDomain", 0);
    Filter::Filter<ValueType> freqFilter;
    if (gradientDomain != 0)
        freqFilter.init(config.get<ValueType>("DT"), static_cast<IndexType>((config.get<ValueType>("T") / config.get<ValueType>("DT")) + 0.5));
    auto const &names = this->getSeismogramHandler().getIsSeismic() ? SeismogramTypeString : SeismogramTypeStringEM;
    auto const markOf = [&](IndexType shotRow, IndexType col) { return mark[(size_t)shotRow * cols + col]; };

    for (IndexType type : receiverTypes) {
        Seismogram<ValueType> &seismo = this->getSeismogramHandler().getSeismogram(type - 1);
        std::vector<ValueType> &data = seismo.getData();
        std::vector<std::vector<ValueType>> &dataDecode = seismo.getDataDecode();
        const IndexType NT = seismo.getNumSamples();
        IndexType countDecode = 0;
        if (toSuper)
            std::fill(data.begin(), data.end(), ValueType(0));
        else {
            dataDecode.clear();
            for (size_t k = 0; k < shotIndsIncr.size(); k++)
                if (std::abs(sourceSettingsEncode[k].sourceNo) == shotNumber)
                    dataDecode.emplace_back();
        }
        // polarity and frequency of the shot; rows = 0: ONE trace through the vector form of the filter (common-offset branch of the reference),
        // else the matrix form (which scales the decoded matrix to maximum amplitude 1, Filter.cpp:331-341)
        auto const finish = [&](std::vector<ValueType> &traces, IndexType rows, sourceSettings<ValueType> const &enc) {
            if (enc.amp < 0)
                for (auto &v : traces)
                    v = -v;
            if (gradientDomain != 0) {
                freqFilter.calc("ideal", "bp", 1, enc.fc);
                if (rows == 0)
                    freqFilter.apply(traces);
                else
                    freqFilter.apply(traces, rows, NT);
            }
        };
        if (numshots == numrecs) { // common-offset data: receiver k belongs to shot k, one trace per shot
            std::vector<ValueType> dataSingle((size_t)NT, ValueType(0));
            IndexType countEncode = 0;
            for (size_t k = 0; k < shotIndsIncr.size(); k++) {
                const IndexType row = shotIndsIncr[k];
                if (receiverMarkVector[row + 1] == 0 || all[row].receiverType != type)
                    continue;
                if (toSuper && std::abs(sourceSettingsEncode[k].sourceNo) != shotNumber)
                    continue;
                SCAI_ASSERT_ERROR((size_t)(countEncode + 1) * NT <= data.size(), "more marked shots than traces of the supershot")
                if (toSuper) {
                    if (encodeType == 1) { // row `row` of the gather <filename>.<type>
                        std::vector<ValueType> gather;
                        IndexType r = 0, c = 0;
                        IO::readMatrix(gather, r, c, filename + "." + names[type - 1], seismoFormat);
                        SCAI_ASSERT_ERROR(row < r && c == NT, "common-offset gather " << filename << "." << names[type - 1] << " does not fit")
                        std::copy(gather.begin() + (size_t)row * NT, gather.begin() + (size_t)(row + 1) * NT, dataSingle.begin());
                    } else {
                        SCAI_ASSERT_ERROR(countDecode < (IndexType)dataDecode.size() && dataDecode[countDecode].size() == (size_t)NT, "no decoded data to encode")
                        dataSingle = dataDecode[countDecode];
                    }
                    if (markOf(row, row + 1) != 0) {
                        std::vector<ValueType> trace(dataSingle);
                        finish(trace, 0, sourceSettingsEncode[k]);
                        for (IndexType t = 0; t < NT; t++)
                            data[(size_t)countEncode * NT + t] += trace[t];
                    }
                } else {
                    SCAI_ASSERT_ERROR(countDecode < (IndexType)dataDecode.size(), "more marked shots than shots of the supershot")
                    if (markOf(row, row + 1) != 0) {
                        std::vector<ValueType> trace(data.begin() + (size_t)countEncode * NT, data.begin() + (size_t)(countEncode + 1) * NT);
                        finish(trace, 0, sourceSettingsEncode[k]);
                        dataSingle = trace;
                    }
                    dataDecode[countDecode] = dataSingle;
                }
                countEncode++;
                countDecode++;
            }
            continue;
        }
        for (size_t k = 0; k < shotIndsIncr.size(); k++) { // multi-offset data
            if (std::abs(sourceSettingsEncode[k].sourceNo) != shotNumber)
                continue;
            const IndexType row = shotIndsIncr[k];
            const IndexType sourceNo = (IndexType)markOf(row, 0);
            IndexType numrecSingle = 0; // every receiver the shot marks, whatever its type: the matrix of a type keeps that many rows
            for (IndexType r = 0; r < numrecs; r++)
                numrecSingle += (IndexType)markOf(row, r + 1);
            const std::string single = filename + ".shot_" + std::to_string(sourceNo) + "." + names[type - 1];
            std::vector<ValueType> dataSingle((size_t)numrecSingle * NT, ValueType(0));
            IndexType count = 0, countEncode = 0;
            if (toSuper) {
                if (encodeType == 1) {
                    IndexType r = 0, c = 0;
                    IO::readMatrix(dataSingle, r, c, single, seismoFormat);
                    SCAI_ASSERT_ERROR(r == numrecSingle && c == NT, single << " must hold " << numrecSingle << " traces of " << NT << " samples")
                } else {
                    SCAI_ASSERT_ERROR(countDecode < (IndexType)dataDecode.size() && dataDecode[countDecode].size() == dataSingle.size(), "no decoded data to encode")
                    dataSingle = dataDecode[countDecode];
                }
                finish(dataSingle, numrecSingle, sourceSettingsEncode[k]);
            }
            for (IndexType r = 0; r < numrecs; r++) {
                if (receiverMarkVector[r + 1] == 0 || all[r].receiverType != type)
                    continue;
                if (markOf(row, r + 1) != 0) {
                    SCAI_ASSERT_ERROR((size_t)(countEncode + 1) * NT <= data.size() && count < numrecSingle, "mark matrix and seismogram do not fit")
                    if (toSuper)
                        for (IndexType t = 0; t < NT; t++)
                            data[(size_t)countEncode * NT + t] += dataSingle[(size_t)count * NT + t];
                    else
                        std::copy(data.begin() + (size_t)countEncode * NT, data.begin() + (size_t)(countEncode + 1) * NT, dataSingle.begin() + (size_t)count * NT);
                    count++;
                }
                countEncode++;
            }
            if (!toSuper) {
                finish(dataSingle, numrecSingle, sourceSettingsEncode[k]);
                dataDecode[countDecode] = dataSingle;
                if (encodeType == 3)
                    IO::writeMatrix(dataSingle, numrecSingle, NT, single, seismoFormat);
            }
            countDecode++;
        }
    }
}

template void Acquisition::readAllSettings<float>(std::vector<sourceSettings<float>> &, std::string);
template void Acquisition::calcuniqueShotNo<float>(std::vector<IndexType> &, std::vector<sourceSettings<float>> const &);
template void Acquisition::getuniqueShotInd<float>(IndexType &, std::vector<sourceSettings<float>> const &, IndexType);
template void Acquisition::getCutCoord<float>(Configuration::Configuration const &, std::vector<coordinate3D> &, std::vector<sourceSettings<float>> const &, Coordinates<float> const &,
                                              Coordinates<float> const &);
template void Acquisition::getSettingsPerShot<float>(std::vector<sourceSettings<float>> &, std::vector<sourceSettings<float>> const &, std::vector<coordinate3D> const &, Coordinates<float> const &,
                                                     IndexType);
template void Acquisition::getSettingsPerShot<float>(std::vector<receiverSettings> &, std::vector<receiverSettings> const &, coordinate3D const &, Coordinates<float> const &, IndexType);
template void Acquisition::createSettingsForShot<float>(std::vector<sourceSettings<float>> &, std::vector<sourceSettings<float>> const &, IndexType);
template class Acquisition::Seismogram<float>;
template class Acquisition::SeismogramHandler<float>;
template class Acquisition::AcquisitionGeometry<float>;
template class Acquisition::Sources<float>;
template class Acquisition::Receivers<float>;

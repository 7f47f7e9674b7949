// Acquisition.hpp — sources, receivers and seismograms of the reference (mirror of src/Acquisition/*).
//   * source file  `<SourceFilename>.txt`: 10 columns  sourceNo X Y Z sourceType waveletType waveletShape fc amp tShift
//     (AcquisitionSettings.hpp:55-109); receiver file `<ReceiverFilename>.txt`: 4 columns X Y Z receiverType;
//     per-shot receivers `<ReceiverFilename>.shot_<n>.txt` (Receivers.cpp:209-239)
//   * types 1..4 = P,VX,VY,VZ (seismic) or EZ,EX,EY,HZ (EM)                     (Acquisition.hpp:17-55)
//   * the nine analytic wavelets of Acquisition/SourceSignal/*.cpp              (Sources.cpp:165-214)
//   * seismograms are nTraces x NT matrices per type, resampled to seismoDT on output and written as
//     `<SeismogramFilename>.shot_<n>.<p|vx|vy|vz|ez|ex|ey|hz>.<mtx|lmf>`        (Seismogram.cpp:82-147, Simulation.cpp:531)
#pragma once
#include "Common.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Filter.hpp"
#include <array>

namespace KITGPI
{
    namespace Acquisition
    {
        enum SeismogramType { P, VX, VY, VZ };
        enum SeismogramTypeEM { EZ, EX, EY, HZ };
        constexpr IndexType NUM_ELEMENTS_SEISMOGRAMTYPE = 4;
        extern const char *const SeismogramTypeString[4];
        extern const char *const SeismogramTypeStringEM[4];

        template <typename ValueType> struct sourceSettings {
            IndexType sourceNo;
            coordinate3D sourceCoords;
            IndexType sourceType, waveletType, waveletShape;
            ValueType fc, amp, tShift;
            IndexType row;
            coordinate3D getCoords() const { return sourceCoords; }
            IndexType getType() const { return sourceType; }
        };
        struct receiverSettings {
            coordinate3D receiverCoords;
            IndexType receiverType;
            coordinate3D getCoords() const { return receiverCoords; }
            IndexType getType() const { return receiverType; }
        };

        template <typename ValueType> void readAllSettings(std::vector<sourceSettings<ValueType>> &allSettings, std::string fileName);
        void readAllSettings(std::vector<receiverSettings> &allSettings, std::string const &fileName);
        //! unique |sourceNo| values in order of first appearance (AcquisitionSettings.hpp calcuniqueShotNo)
        template <typename ValueType> void calcuniqueShotNo(std::vector<IndexType> &uniqueShotNo, std::vector<sourceSettings<ValueType>> const &sourceSettings);
        template <typename ValueType>
        void createSettingsForShot(std::vector<sourceSettings<ValueType>> &settings, std::vector<sourceSettings<ValueType>> const &allSettings, IndexType shotNumber);

        //! index of shotNumber in the list of unique shot numbers / of the first encoded source with that number (AcquisitionSettings.hpp:208-232)
        void getuniqueShotInd(IndexType &shotInd, std::vector<IndexType> const &uniqueShotNos, IndexType shotNumber);
        template <typename ValueType> void getuniqueShotInd(IndexType &shotInd, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode, IndexType shotNumber);
        //! per-shot model cut-outs (useStreamConfig, AcquisitionSettings.hpp:136-187, 364-430): where the sub-model of a shot starts in
        //! the big model, and the acquisition settings of the big model moved into the sub-model
        template <typename ValueType>
        void getCutCoord(Configuration::Configuration const &config, std::vector<coordinate3D> &cutCoordinates, std::vector<sourceSettings<ValueType>> const &sourceSettingsBig,
                         Coordinates<ValueType> const &modelCoordinates, Coordinates<ValueType> const &modelCoordinatesBig);
        template <typename ValueType>
        void getSettingsPerShot(std::vector<sourceSettings<ValueType>> &settings, std::vector<sourceSettings<ValueType>> const &allSettings, std::vector<coordinate3D> const &cutCoordinates,
                                Coordinates<ValueType> const &modelCoordinates, IndexType BoundaryWidth);
        template <typename ValueType>
        void getSettingsPerShot(std::vector<receiverSettings> &settings, std::vector<receiverSettings> const &allSettings, coordinate3D const &cutCoordinate,
                                Coordinates<ValueType> const &modelCoordinates, IndexType BoundaryWidth);
        void writeCutCoordToFile(Configuration::Configuration const &config, std::string const &sourceFilename, std::vector<coordinate3D> const &cutCoordinates,
                                 std::vector<IndexType> const &uniqueShotNos, IndexType NXPerShot);
        //! shot sequence of one pass over the shot domains (useRandomSource 1 random without repetition, 2 / 3 sequential covering the
        //! global / local area; AcquisitionSettings.hpp:437-483)
        void getRandomShotInds(std::vector<IndexType> &uniqueShotInds, std::vector<IndexType> &shotHistory, IndexType numshots, IndexType maxcount, IndexType useRandomSource,
                               IndexType &seedtime);

        namespace SourceSignal
        {
            //! waveletShape 1 Ricker, 2 SinW, 3 SinThree, 4 FGaussian, 5 Spike, 6 IntgSinThree, 7 Ricker_GprMax, 8 Berlage, 9 Sin
            void calc(IndexType waveletShape, std::vector<ValueType> &signal, IndexType NT, ValueType DT, ValueType FC, ValueType AMP, ValueType Tshift);
        }

        //! traces of one type
        template <typename ValueType> class Seismogram
        {
          public:
            void allocate(IndexType numTraces, IndexType NT);
            void resetData() { std::fill(data.begin(), data.end(), ValueType(0)); }
            IndexType getNumTraces() const { return (IndexType)coordinates1D.size(); }
            IndexType getNumSamples() const { return numSamples; }
            std::vector<ValueType> &getData() { return data; }             // row-major nTraces x NT
            std::vector<ValueType> const &getData() const { return data; }
            std::vector<IndexType> &getCoordinates1D() { return coordinates1D; }
            std::vector<IndexType> const &getCoordinates1D() const { return coordinates1D; }
            void setDT(ValueType dt) { DT = dt; }
            void setSeismoDT(ValueType dt) { outputDT = dt; }
            void setTraceType(IndexType t, bool seismic) { type = t; isSeismic = seismic; }
            IndexType getTraceType() const { return type; }
            //! 1 maximum, 2 l2 norm, 3 l2 norm and automatic gain control (after calcInverseAGC), 4 l2 norm and envelope with a water
            //! level (Seismogram.cpp:215-260)
            void normalizeTrace(IndexType normalizeTraces);
            //! automatic gain control (Seismogram.cpp:262-385): running mean of the l2-normalised trace / inverse running rms over a window of
            //! 1 / (frequencyAGC DT) samples each side, evaluated from the end of the trace backwards
            void setFrequencyAGC(ValueType f) { frequencyAGC = f; }
            std::vector<ValueType> getAGCSum();
            void calcInverseAGC();
            std::vector<ValueType> &getInverseAGC() { return inverseAGC; }
            void setInstantaneousTrace(IndexType instantaneousTraces) { outputInstantaneous = instantaneousTraces; } // Seismogram.cpp write: 1 envelope, 2 phase
            //! common-offset profile: a survey of single-trace shots is gathered into ONE matrix, row = shot index, written once at the
            //! end instead of a file per shot (Seismogram.cpp:84-96, SeismogramHandler.cpp:489-516, Simulation.cpp:356-361, 535-560)
            std::vector<ValueType> &getDataCOP() { return dataCOP; }
            std::vector<ValueType> &getInverseAGCCOP() { return inverseAGCCOP; }
            void allocateCOP(IndexType numshots, IndexType NT);
            void setShotInd(IndexType shotIndTrue, IndexType shotIndIncr_) { shotInd = shotIndTrue; shotIndIncr = shotIndIncr_; }
            IndexType getNumShotsCOP() const { return numshotsCOP; }
            void assignCOP(); // data = dataCOP (numshots traces), dataCOP = 0
            void filterTraces(Filter::Filter<ValueType> const &freqFilter); // Seismogram.cpp:509-520
            bool isFinite() const;
            //! SeismogramFormat 1 = mtx, 2 = lmf, 4 = SU (needs the model coordinates for the trace headers, Seismogram.cpp:82-147)
            //! 5 = the inverse AGC function as `<filename>.<type>.inverseAGC.mtx`.  A single trace goes into the common-offset profile when one is allocated.
            void write(IndexType seismogramFormat, std::string const &filename, Coordinates<ValueType> const *modelCoordinates = nullptr);
            //! Seismogram.cpp:158-207.  5 = the inverse AGC function `<filename>.<type>.inverseAGC.mtx` (used by the next normalizeTrace(3)).  When a common-offset
            //! profile is allocated and this seismogram holds ONE trace, `filename` (`...shot_<n>`) is cut at ".shot" and the trace is row shotInd of the
            //! profile file (row shotIndIncr with readOriginal: the numbering of the original, unselected shots)
            void read(IndexType seismogramFormat, std::string const &filename, bool readOriginal = false);
            void setSourceCoordinate(IndexType sourceCoord) { sourceCoordinate1D = sourceCoord; } // Seismogram.cpp:561
            IndexType getSourceCoordinate() const { return sourceCoordinate1D; }
            //! the traces of the single shots a supershot was decoded into (Seismogram.hpp getDataDecode): row-major matrices, traces x NT
            std::vector<std::vector<ValueType>> &getDataDecode() { return dataDecode; }
            std::vector<std::vector<ValueType>> const &getDataDecode() const { return dataDecode; }

          private:
            std::vector<std::vector<ValueType>> dataDecode;
            std::vector<ValueType> inverseAGC, dataCOP, inverseAGCCOP;
            ValueType frequencyAGC = 0;
            bool useAGC = false;
            IndexType outputInstantaneous = 0, shotInd = 0, shotIndIncr = 0, numshotsCOP = 0;
            std::vector<ValueType> data;
            std::vector<IndexType> coordinates1D;
            IndexType sourceCoordinate1D = 0;
            IndexType numSamples = 0, type = 0;
            bool isSeismic = true;
            ValueType DT = 0, outputDT = 0;
        };

        template <typename ValueType> class SeismogramHandler
        {
          public:
            Seismogram<ValueType> &getSeismogram(IndexType type) { return seismo[type]; }
            Seismogram<ValueType> const &getSeismogram(IndexType type) const { return seismo[type]; }
            IndexType getNumTracesTotal() const;
            IndexType getNumTracesGlobal(IndexType type) const { return seismo[type].getNumTraces(); }
            void setIsSeismic(bool s) { isSeismic = s; }
            bool getIsSeismic() const { return isSeismic; }
            void setDT(ValueType dt);
            void setSeismoDT(ValueType dt);
            void resetData();
            void normalize(IndexType normalizeTraces);
            void setFrequencyAGC(ValueType f);
            void calcInverseAGC();
            void setInstantaneousTrace(IndexType instantaneousTraces);
            void allocateCOP(IndexType numshots, IndexType NT);
            void setShotInd(IndexType shotIndTrue, IndexType shotIndIncr);
            //! adds the common-offset profiles another shot domain has gathered (SeismogramHandler.cpp:479-487: the reduction over commInterShot)
            void sumShotDomain(SeismogramHandler<ValueType> &other);
            void assignCOP(); // first type that has traces (SeismogramHandler.cpp:505-516)
            void filter(Filter::Filter<ValueType> const &freqFilter); // SeismogramHandler.cpp:76-86
            bool isFinite() const;
            void write(IndexType seismogramFormat, std::string const &filename, Coordinates<ValueType> const *modelCoordinates = nullptr);
            void setSourceCoordinate(IndexType sourceCoord); // SeismogramHandler.cpp:340

          private:
            std::array<Seismogram<ValueType>, NUM_ELEMENTS_SEISMOGRAMTYPE> seismo;
            bool isSeismic = true;
        };

        //! coordinates + types -> per-type 1-D coordinate lists (AcquisitionGeometry.hpp:76-120, AcquisitionGeometry.cpp:18-76)
        template <typename ValueType> class AcquisitionGeometry
        {
          public:
            SeismogramHandler<ValueType> &getSeismogramHandler() { return seismograms; }
            SeismogramHandler<ValueType> const &getSeismogramHandler() const { return seismograms; }
            IndexType getNumTracesGlobal() const { return (IndexType)coordinates1D.size(); }
            std::vector<IndexType> const &get1DCoordinates() const { return coordinates1D; }
            std::vector<IndexType> const &getSeismogramTypes() const { return types; }
            unsigned long getVersion() const { return version; } // bumped whenever the geometry or the signals change

          protected:
            template <typename Settings> void setAcquisition(std::vector<Settings> const &allSettings, Coordinates<ValueType> const &modelCoordinates, IndexType NT);
            SeismogramHandler<ValueType> seismograms;
            std::vector<IndexType> coordinates1D, types; // in file order; types 1..4
            std::vector<IndexType> traceOfEntry;        // row inside the seismogram of its type
            unsigned long version = 0;
        };

        template <typename ValueType> class Sources : public AcquisitionGeometry<ValueType>
        {
          public:
            //! reads <SourceFilename>.txt (or the SU files); shotIncr > DH keeps one shot every shotIncr metres (Sources.cpp:500-560)
            void getAcquisitionSettings(Configuration::Configuration const &config, ValueType shotIncr = 0);
            std::vector<sourceSettings<ValueType>> const &getSourceSettings() const { return allSourceSettings; }
            void setSourceSettings(std::vector<sourceSettings<ValueType>> const &settings) { allSourceSettings = settings; }
            std::vector<IndexType> const &getShotIndsIncr() const { return shotIndsIncr; }
            //! useSourceEncode 1 / 2 / 3: the shots are merged into NumShotDomains supershots (random with random polarity / interleaved
            //! / blockwise; Sources.cpp:568-646, time-domain part)
            //! gradientDomain != 0 adds the frequency bookkeeping of the inversion's frequency-domain gradient (Sources.cpp:584-612, 648-674): the
            //! frequency list per shot (getSourceFC) and, with encoding, a sine source of its own frequency for every shot of a supershot
            void calcSourceSettingsEncode(Configuration::Configuration const &config, IndexType &seedtime, ValueType fc1 = 0, ValueType fc2 = 0);
            std::vector<std::vector<ValueType>> const &getSourceFC() const { return sourceFC; }
            void writeSourceFC(Configuration::Configuration const &config, IndexType stage = 0, IndexType iteration = 0) const; // <SourceFilename>.sourceFC.txt
            std::vector<sourceSettings<ValueType>> const &getSourceSettingsEncode() const { return sourceSettingsEncode; }
            //! the shot indices one pass over the shot domains works on (Sources.cpp:687-714)
            void calcUniqueShotInds(Configuration::Configuration const &config, std::vector<IndexType> &shotHistory, IndexType maxcount, IndexType &seedtime);
            std::vector<IndexType> const &getUniqueShotInds() const { return uniqueShotInds; }
            void writeShotIndsIncr(Configuration::Configuration const &config, std::vector<IndexType> const &uniqueShotNos) const; // <SourceFilename>.shotIncr.txt
            void writeSourceEncode(Configuration::Configuration const &config) const;                                              // <SourceFilename>.encode.txt
            //! sources of one shot: geometry + signals (Sources.cpp:16-46, 99-145)
            void init(std::vector<sourceSettings<ValueType>> const &shotSettings, Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates);
            IndexType getRowOfEntry(IndexType k) const { return this->traceOfEntry[k]; }

          private:
            std::vector<sourceSettings<ValueType>> allSourceSettings; // after the shotIncr selection (sourceSettingsShotIncr of the reference)
            std::vector<sourceSettings<ValueType>> sourceSettingsEncode;
            std::vector<std::vector<ValueType>> sourceFC;
            std::vector<IndexType> shotIndsIncr, uniqueShotInds;
        };

        template <typename ValueType> class Receivers : public AcquisitionGeometry<ValueType>
        {
          public:
            void init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates);                       // <ReceiverFilename>.txt
            void init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates, IndexType shotNumber); // .shot_<n>.txt
            void init(std::vector<receiverSettings> const &allSettings, Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates);
            //! useReceiversPerShot = 2 (Receivers.cpp:250-300): ONE receiver file for the survey and a mark matrix `<ReceiverFilename>.mark.mtx`
            //! of numshots rows [shot number | one mark per receiver]: a shot records the receivers its row marks; a supershot
            //! (sourceSettingsEncode not empty) those any of its shots marks.  shotIndsIncr = the rows of the selected shots (shotIncr).
            void init(Configuration::Configuration const &config, Coordinates<ValueType> const &modelCoordinates, IndexType shotNumber, IndexType numshots,
                      std::vector<IndexType> const &shotIndsIncr, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode);
            //! the receivers of the survey file the marks select for the shot / supershot (Receivers.cpp:250-300); sets the mark vector
            void getAcquisitionSettings(Configuration::Configuration const &config, std::vector<receiverSettings> &active, IndexType shotNumber, IndexType numshots,
                                        std::vector<IndexType> const &shotIndsIncr, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode);
            std::vector<ValueType> const &getReceiverMarkVector() const { return receiverMarkVector; }
            //! Receivers.cpp:353-576, the time-domain part: `decode` splits the seismograms of supershot `shotNumber` into those of its
            //! shots (the receivers the mark matrix gives each shot, polarity of the encoding undone; with gradientDomain != 0 the
            //! ideal band pass at the shot's frequency) and writes them as `<filename>.shot_<n>.<type>` (encodeType 1; 0 keeps them in
            //! getDataDecode only); `encode` is the way back (encodeType 1 reads the single-shot files, 0 takes getDataDecode).
            //! Common-offset surveys (as many shots as receivers) keep one trace per shot and write nothing, as the reference.
            void decode(Configuration::Configuration const &config, std::string const &filename, IndexType shotNumber, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode,
                        IndexType encodeType);
            void encode(Configuration::Configuration const &config, std::string const &filename, IndexType shotNumber, std::vector<sourceSettings<ValueType>> const &sourceSettingsEncode,
                        IndexType encodeType);
            //! `<ReceiverFilename>.shot_<n>.mark.mtx`: the mark vector of the supershot (Receivers.cpp:308-328; only with useSourceEncode)
            void writeReceiverMark(Configuration::Configuration const &config, IndexType shotNumber, IndexType stage = 0, IndexType iteration = 0) const;
            IndexType getRowOfEntry(IndexType k) const { return this->traceOfEntry[k]; }

          private:
            void readMarkMatrix(Configuration::Configuration const &config, std::vector<receiverSettings> &all, std::vector<ValueType> &mark, IndexType numshots) const;
            std::vector<ValueType> receiverMarkVector; // [shot number | 0 / 1 per receiver of the file]
        };
    }
}

// Unit tests of the LAMA-free host layer (wave-simulation_b200/host), modelled on the reference's googletest cases
// (src/Tests/UnitTest/{Configuration,Coordinates,Ricker,...}UnitTest.cpp): same known answers, plain asserts.
// Built and run by tests/test_host_layer.py; needs no GPU (linked against the emulation build of the solver library, which the
// model classes reference; nothing of it runs here).
#include "Acquisition.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include "IO.hpp"
#include "Modelparameter.hpp"
#include <cstdio>
#include <fstream>

using namespace KITGPI;

static int failures = 0;
#define EXPECT(cond)                                                                                                   \
    if (!(cond)) {                                                                                                     \
        std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                                  \
        failures++;                                                                                                    \
    }
#define EXPECT_THROW(stmt)                                                                                             \
    {                                                                                                                  \
        bool thrown = false;                                                                                           \
        try {                                                                                                          \
            stmt;                                                                                                      \
        } catch (std::exception const &) {                                                                             \
            thrown = true;                                                                                             \
        }                                                                                                              \
        EXPECT(thrown);                                                                                                \
    }

int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : ".";
    // ---- Configuration (ConfigurationUnitTest.cpp)
    {
        std::ofstream f(dir + "/configuration_1.txt");
        f << "# Comment\ntestvalue1=1.2124445\ntestvalue2=100\ntestvalue3=-1.2124445\ntestvalue4=-100\ntestvalue5=test123\n# Comment # comment # comment\n"
             "TESTVALUE6=capiTAL # test 1 2 3 \ntestvalue7=1 # comment \ntestvalue8=0 #comment really long\ntestvalue9=/file/path/test.mtx\n";
        f.close();
        std::ofstream g(dir + "/configuration_2.txt");
        g << "testvalue1=99\ntestvalue10=7\n";
        g.close();
        EXPECT_THROW(Configuration::Configuration bad(dir + "/configuration_100.txt"));
        Configuration::Configuration config(dir + "/configuration_1.txt");
        EXPECT(config.get<double>("testvalue1") == 1.2124445);
        EXPECT(config.get<int>("testvalue2") == 100);
        EXPECT(config.get<double>("testvalue3") == -1.2124445);
        EXPECT(config.get<int>("testvalue4") == -100);
        EXPECT(config.get<std::string>("testvalue5") == "test123");
        EXPECT(config.get<std::string>("TESTVALUE6") == "capiTAL");
        EXPECT(config.get<std::string>("testvalue6") == "capiTAL");
        EXPECT(config.get<bool>("testvalue7"));
        EXPECT(!config.get<bool>("testvalue8"));
        EXPECT(config.get<std::string>("testvalue9") == "/file/path/test.mtx");
        EXPECT_THROW(config.get<std::string>("UnkownValue"));
        EXPECT(config.getAndCatch("UnkownValue", 42) == 42);
        config.readFromFile(dir + "/configuration_2.txt"); // first occurrence wins unless overwrite
        EXPECT(config.get<double>("testvalue1") == 1.2124445);
        EXPECT(config.get<int>("testvalue10") == 7);
        config.readFromFile(dir + "/configuration_2.txt", true);
        EXPECT(config.get<double>("testvalue1") == 99);
        // the 2D quirk of Configuration.cpp:77-82
        std::ofstream q(dir + "/configuration_2d.txt");
        q << "dimension=2D              # Dimension\nNZ=105\n";
        q.close();
        Configuration::Configuration c2(dir + "/configuration_2d.txt");
        EXPECT(c2.get<int>("NZ") == 1);
    }
    // ---- Coordinates (CoordinatesUnitTest.cpp:28-130)
    {
        Acquisition::Coordinates<ValueType> test(5, 15, 10, 1.0f);
        Acquisition::coordinate3D r = test.index2coordinate(112);
        EXPECT(r.x == 2 && r.y == 2 && r.z == 2);
        EXPECT(test.coordinate2index(4, 2, 3) == 119);
        EXPECT(test.locatedOnSurface(2) && test.locatedOnSurface(49) && !test.locatedOnSurface(80));
        Acquisition::coordinate3D d = test.edgeDistance({1, 13, 4});
        EXPECT(d.x == 1 && d.y == 1 && d.z == 4);
        EXPECT_THROW(test.coordinate2index(5, 0, 0));
        EXPECT_THROW(test.coordinate2index(0, -1, 0));
    }
    // ---- wavelets (RickerUnitTest.cpp:11-44 and siblings: one sample against an independent evaluation)
    {
        std::vector<ValueType> s;
        const IndexType NT = 200;
        const ValueType DT = 2e-3f, FC = 5.0f, AMP = 5.0f, TS = 0.0f;
        Acquisition::SourceSignal::calc(1, s, NT, DT, FC, AMP, TS);
        const int k = 150;
        double tau = M_PI * FC * (k * (double)DT - 1.5 / FC - TS);
        EXPECT(std::abs(s[k] - AMP * (1 - 2 * tau * tau) * std::exp(-tau * tau)) < 1e-5);
        Acquisition::SourceSignal::calc(4, s, NT, DT, FC, AMP, TS); // FGaussian
        tau = M_PI * FC * (k * (double)DT - 1.2 / FC - TS);
        EXPECT(std::abs(s[k] - AMP * (-2 * tau) * std::exp(-tau * tau)) < 1e-5);
        Acquisition::SourceSignal::calc(3, s, NT, DT, FC, AMP, 0.021f); // SinThree starts at floor(tShift/DT)
        EXPECT(s[5] == 0 && std::abs(s[10 + 20] - AMP * std::pow(std::sin(20 * (double)DT * M_PI * FC), 3)) < 1e-5);
        Acquisition::SourceSignal::calc(5, s, NT, DT, FC, AMP, 0.1f); // Spike
        EXPECT(s[50] == AMP && s[49] == 0 && s[51] == 0);
        Acquisition::SourceSignal::calc(9, s, NT, DT, FC, AMP, TS);
        EXPECT(std::abs(s[k] - AMP * std::sin(2 * M_PI * FC * k * DT)) < 1e-4);
        Acquisition::SourceSignal::calc(7, s, NT, DT, FC, AMP, TS); // Ricker_GprMax: zero crossing at t = 1/fc
        EXPECT(std::abs(s[100]) < 1e-6 * AMP && s[90] != 0);
        Acquisition::SourceSignal::calc(8, s, NT, DT, FC, AMP, TS); // Berlage: causal, normalised to AMP
        ValueType mx = 0;
        for (ValueType v : s)
            mx = std::max(mx, std::abs(v));
        EXPECT(std::abs(mx - AMP) < 1e-4 && s[50] == 0);
        EXPECT_THROW(Acquisition::SourceSignal::calc(17, s, NT, DT, FC, AMP, TS));
    }
    // ---- AcquisitionUnitTest.cpp:9-32, AcousticUnitTest.cpp:12-25, ReceiversUnitTest.cpp:11-25: enumeration and names of the seismogram
    //      types; model / receiver initialisation from a configuration that lacks their keys throws
    {
        EXPECT(Acquisition::P == 0 && Acquisition::VX == 1 && Acquisition::VY == 2 && Acquisition::VZ == 3);
        EXPECT(std::string(Acquisition::SeismogramTypeString[Acquisition::P]) == "p" && std::string(Acquisition::SeismogramTypeString[Acquisition::VX]) == "vx" &&
               std::string(Acquisition::SeismogramTypeString[Acquisition::VY]) == "vy" && std::string(Acquisition::SeismogramTypeString[Acquisition::VZ]) == "vz");
        EXPECT(Acquisition::NUM_ELEMENTS_SEISMOGRAMTYPE == 4);
        std::ofstream f(dir + "/configuration_bare.txt");
        f << "NX=10\nNY=10\nNZ=1\nDH=50\n";
        f.close();
        Configuration::Configuration bare(dir + "/configuration_bare.txt");
        Acquisition::Coordinates<ValueType> coords(bare.get<IndexType>("NX"), bare.get<IndexType>("NY"), bare.get<IndexType>("NZ"), bare.get<ValueType>("DH"));
        Modelparameter::Modelparameter<ValueType> acoustic("acoustic");
        EXPECT_THROW(acoustic.init(bare, coords));
        Acquisition::Receivers<ValueType> receivers;
        EXPECT_THROW(receivers.init(bare, coords));
    }
    // ---- acquisition files
    {
        std::ofstream f(dir + "/sources.txt");
        f << "# sourceNo X Y Z type wtype wshape fc amp tshift\n1 20 0 0 2 1 1 5.0 5.0 0.0\n\n 2 30 1 0 1 1 1 10 1 0.1\n-2 31 1 0 3 1 4 10 1 0\n";
        f.close();
        std::vector<Acquisition::sourceSettings<ValueType>> all, shot;
        Acquisition::readAllSettings(all, dir + "/sources.txt");
        EXPECT(all.size() == 3 && all[1].sourceCoords.x == 30 && all[2].waveletShape == 4 && all[2].row == 2);
        std::vector<IndexType> uniq;
        Acquisition::calcuniqueShotNo(uniq, all);
        EXPECT(uniq.size() == 2 && uniq[0] == 1 && uniq[1] == 2);
        Acquisition::createSettingsForShot(shot, all, 2);
        EXPECT(shot.size() == 2);
        std::ofstream b(dir + "/bad.txt");
        b << "1 2 3\n";
        b.close();
        EXPECT_THROW(Acquisition::readAllSettings(all, dir + "/bad.txt"));
        std::vector<Acquisition::receiverSettings> rec;
        EXPECT_THROW(Acquisition::readAllSettings(rec, dir + "/bad.txt"));
        EXPECT_THROW(Acquisition::readAllSettings(rec, dir + "/missing.txt"));
    }
    // ---- receivers by mark matrix, supershots decoded into their shots and encoded again (Receivers.cpp:250-300, 353-576)
    for (int commonOffset = 0; commonOffset < 2; commonOffset++) {
        const IndexType numrecs = commonOffset ? 3 : 4, NT = 5;
        {
            std::ofstream c(dir + "/enc.txt");
            c << "dimension=2D\nequationType=elastic\nNX=50\nNY=20\nNZ=1\nDH=10\nDT=0.1\nT=0.5\nseismoDT=0.1\nUseVariableGrid=0\nuseSourceEncode=2\nuseReceiversPerShot=2\n"
                 "SeismogramFormat=1\nSourceFilename=" << dir << "/encsrc\nReceiverFilename=" << dir << "/encrec\n";
            std::ofstream f(dir + "/encsrc.txt");
            f << "1 5 1 0 1 1 1 5 1 0\n2 15 1 0 1 1 1 5 1 0\n3 25 1 0 1 1 1 5 1 0\n";
            std::ofstream r(dir + "/encrec.txt");
            r << "10 2 0 3\n20 2 0 1\n30 2 0 3\n";
            if (!commonOffset)
                r << "40 2 0 3\n";
            std::ofstream m(dir + "/encrec.mark.mtx"); // coordinate format, as LAMA writes a sparse matrix
            if (commonOffset)
                m << "%%MatrixMarket matrix coordinate real general\n3 4 6\n1 1 1\n1 2 1\n2 1 2\n2 3 1\n3 1 3\n3 4 1\n";
            else // shot 1: receivers 1, 2, 4;  shot 2: receiver 3;  shot 3: receivers 2 (p), 3, 4
                m << "%%MatrixMarket matrix coordinate real general\n3 5 10\n1 1 1\n1 2 1\n1 3 1\n1 5 1\n2 1 2\n2 4 1\n3 1 3\n3 3 1\n3 4 1\n3 5 1\n";
        }
        Configuration::Configuration config(dir + "/enc.txt");
        Acquisition::Coordinates<ValueType> coords(config);
        // supershot 20001 = shots 1 and 3 (3 with negative polarity), supershot 20002 = shot 2
        std::vector<Acquisition::sourceSettings<ValueType>> enc;
        Acquisition::readAllSettings(enc, dir + "/encsrc.txt");
        enc[0].sourceNo = 20001;
        enc[1].sourceNo = 20002;
        enc[2].sourceNo = 20001;
        enc[2].amp = -1;
        const std::vector<IndexType> rows = {0, 1, 2};
        Acquisition::Receivers<ValueType> receivers;
        receivers.init(config, coords, 20001, 3, rows, enc);
        auto const &mark = receivers.getReceiverMarkVector();
        auto &vy = receivers.getSeismogramHandler().getSeismogram(Acquisition::VY).getData();
        auto &p = receivers.getSeismogramHandler().getSeismogram(Acquisition::P).getData();
        auto const fill = [&]() {
            for (size_t k = 0; k < vy.size(); k++)
                vy[k] = (ValueType)(100 * (k / NT) + k % NT + 1);
            for (size_t k = 0; k < p.size(); k++)
                p[k] = (ValueType)(-7 - (ValueType)k);
        };
        fill();
        if (commonOffset) {
            // the union of rows 1 and 3: receivers 1 and 3 (both vy); one trace per shot
            EXPECT(mark.size() == 4 && mark[0] == 20001 && mark[1] == 1 && mark[2] == 0 && mark[3] == 1 && vy.size() == (size_t)2 * NT && p.empty());
            receivers.decode(config, dir + "/encseis", 20001, enc, 0);
            auto const &dec = receivers.getSeismogramHandler().getSeismogram(Acquisition::VY).getDataDecode();
            EXPECT(dec.size() == 2 && dec[0].size() == (size_t)NT && dec[0][0] == 1 && dec[0][4] == 5 && dec[1][0] == -101 && dec[1][4] == -105);
            std::fill(vy.begin(), vy.end(), ValueType(9));
            receivers.encode(config, dir + "/encseis", 20001, enc, 0);
            EXPECT(vy[0] == 1 && vy[4] == 5 && vy[5] == 101 && vy[9] == 105);
            continue;
        }
        EXPECT(mark.size() == (size_t)numrecs + 1 && mark[1] == 1 && mark[2] == 1 && mark[3] == 1 && mark[4] == 1 && vy.size() == (size_t)3 * NT && p.size() == (size_t)NT);
        receivers.decode(config, dir + "/encseis", 20001, enc, 1); // to getDataDecode and to files
        auto const &dec = receivers.getSeismogramHandler().getSeismogram(Acquisition::VY).getDataDecode();
        auto const &decP = receivers.getSeismogramHandler().getSeismogram(Acquisition::P).getDataDecode();
        // shot 1 marks 3 receivers (2 of them vy: traces 0 and 2 of the supershot), shot 3 marks 3 (vy traces 1 and 2), sign -1; a matrix keeps
        // one row per marked receiver of ANY type (Receivers.cpp:507-512), the rows of the other types stay zero
        EXPECT(dec.size() == 2 && dec[0].size() == (size_t)3 * NT && dec[0][0] == 1 && dec[0][NT] == 201 && dec[0][2 * NT] == 0);
        EXPECT(dec[1].size() == (size_t)3 * NT && dec[1][0] == -101 && dec[1][NT + 4] == -205 && dec[1][2 * NT] == 0);
        EXPECT(decP.size() == 2 && decP[0][0] == -7 && decP[0][NT] == 0 && decP[1][0] == 7);
        std::vector<ValueType> file;
        IndexType r = 0, c = 0;
        IO::readMatrix(file, r, c, dir + "/encseis.shot_3.vy", 1);
        EXPECT(r == 3 && c == NT && file == dec[1]);
        IO::readMatrix(file, r, c, dir + "/encseis.shot_1.p", 1);
        EXPECT(r == 3 && c == NT && file == decP[0]);
        for (IndexType encodeType = 0; encodeType < 2; encodeType++) { // back: from getDataDecode, from the files
            std::fill(vy.begin(), vy.end(), ValueType(9));
            std::fill(p.begin(), p.end(), ValueType(9));
            receivers.encode(config, dir + "/encseis", 20001, enc, encodeType);
            // receiver 1 (trace 0) is shot 1's alone, receiver 3 (trace 1) shot 3's alone, receiver 4 (trace 2) and the p receiver are shared: twice
            EXPECT(vy[0] == 1 && vy[NT] == 101 && vy[2 * NT + 1] == 2 * 202 && p[0] == -14);
        }
        receivers.writeReceiverMark(config, 20001);
        std::vector<ValueType> mv(numrecs + 1);
        IO::readVector(mv, dir + "/encrec.shot_20001.mark", 1);
        EXPECT(mv == mark);
        // a plain shot takes its own row; a shot that has no row is refused
        receivers.init(config, coords, 2, 3, rows, {});
        EXPECT(receivers.getNumTracesGlobal() == 1 && receivers.getReceiverMarkVector()[3] == 1);
        EXPECT_THROW(receivers.init(config, coords, 7, 3, rows, {}));
        EXPECT_THROW(receivers.init(config, coords, 2, 4, rows, {})); // numshots does not fit the matrix
    }
    // ---- EM model files: relative values on disk, the visco types hold the real EFFECTIVE permittivity / conductivity (ViscoTMEM.cpp:320-402)
    {
        const char *base = "dimension=2D\nNX=6\nNY=5\nNZ=1\nDH=0.02\nUseVariableGrid=0\nfileFormat=1\nnumRelaxationMechanisms=2\nrelaxationFrequency=5.0e7\n"
                           "relaxationFrequency2=3.0e8\nCenterFrequencyCPML=1.0e8\nmur=1.5\nsigma=0.01\nepsilonr=6\ntauSigmar=0.3\ntauEpsilon=0.2\n";
        {
            std::ofstream c0(dir + "/em0.txt"), c1(dir + "/em1.txt");
            c0 << base << "equationType=viscotmem\nModelRead=0\nModelFilename=" << dir << "/emmodel\n";
            c1 << base << "equationType=viscotmem\nModelRead=1\nModelFilename=" << dir << "/emmodel\n";
        }
        Configuration::Configuration c0(dir + "/em0.txt"), c1(dir + "/em1.txt");
        Acquisition::Coordinates<ValueType> coords(c0);
        Modelparameter::Modelparameter<ValueType> a("viscotmem"), b("viscotmem");
        a.init(c0, coords);
        std::vector<ValueType> epsStatic = a.getDielectricPermittivity(), sigStatic = a.getElectricConductivity();
        for (size_t i = 0; i < epsStatic.size(); i++) { // a heterogeneous model
            epsStatic[i] *= 1 + 0.05f * (i % 7);
            sigStatic[i] *= 1 + 0.1f * (i % 5);
        }
        a.init("dielectricPermittivity", epsStatic);
        a.init("electricConductivity", sigStatic);
        a.write(dir + "/emmodel", 1);
        std::vector<ValueType> file(30);
        IO::readVector(file, dir + "/emmodel.mur", 1);
        EXPECT(std::abs(file[3] - 1.5f) < 1e-6);
        IO::readVector(file, dir + "/emmodel.tauSigmar", 1);
        EXPECT(std::abs(file[3] - 0.3f) < 1e-6);
        IO::readVector(file, dir + "/emmodel.epsilonr", 1);
        {   // independent evaluation of the effective permittivity of point 3 in double precision
            const double w = 2 * M_PI * 1e8, t1 = 1 / (2 * M_PI * 5e7), t2 = 1 / (2 * M_PI * 3e8);
            const double aAv = 0.5 * (w * w * t1 * t1 / (1 + w * w * t1 * t1) + w * w * t2 * t2 / (1 + w * w * t2 * t2));
            const double tauSig = 0.3 / w;
            const double want = ((double)epsStatic[3] * (1 - aAv * 0.2) + (double)sigStatic[3] * tauSig) / 8.8541878176e-12;
            EXPECT(std::abs(file[3] - want) < 1e-4 * want);
            EXPECT(file[3] > 6.0f * 0.8f && std::abs(file[3] - 6.0f * 1.15f) > 1e-2); // relative, and not the static value
        }
        b.init(c1, coords);
        double worst = 0;
        for (size_t i = 0; i < epsStatic.size(); i++) {
            worst = std::max(worst, std::abs((double)b.getDielectricPermittivity()[i] / epsStatic[i] - 1));
            worst = std::max(worst, std::abs((double)b.getElectricConductivity()[i] / sigStatic[i] - 1));
            worst = std::max(worst, std::abs((double)b.getMagneticPermeability()[i] / a.getMagneticPermeability()[i] - 1));
        }
        EXPECT(worst < 2e-5); // effective -> static undoes static -> effective (9 significant digits in the mtx file)
        // the non-dispersive type writes its parameters as relative values only
        {
            std::ofstream c2(dir + "/em2.txt");
            c2 << base << "equationType=tmem\nModelRead=0\nModelFilename=" << dir << "/emmodel2\n";
        }
        Configuration::Configuration c2(dir + "/em2.txt");
        Modelparameter::Modelparameter<ValueType> t("tmem");
        t.init(c2, coords);
        t.write(dir + "/emmodel2", 1);
        IO::readVector(file, dir + "/emmodel2.epsilonr", 1);
        EXPECT(std::abs(file[7] - 6.0f) < 1e-5);
        IO::readVector(file, dir + "/emmodel2.sigma", 1);
        EXPECT(std::abs(file[7] - 0.01f) < 1e-8);
    }
    // ---- models on a variable grid: ModelRead = 1 maps the REGULAR model file onto the grid, ModelRead = 2 reads the grid itself (Acoustic.cpp init)
    {
        {
            std::ofstream g(dir + "/vgGrid.txt");
            g << "# interface dhFactor\n0 1\n4 3\n10 1\n";
            std::ofstream c(dir + "/vg.txt");
            c << "dimension=2D\nequationType=acoustic\nNX=9\nNY=12\nNZ=1\nDH=10\nUseVariableGrid=1\nfileFormat=1\nModelRead=1\nModelFilename=" << dir << "/vgmodel\n"
              << "gridConfigurationFilename=" << dir << "/vgGrid.txt\n";
        }
        Configuration::Configuration c(dir + "/vg.txt");
        Acquisition::Coordinates<ValueType> coords(c), regular(9, 12, 1, 10);
        const IndexType n = coords.getNGridpoints();
        EXPECT(n < 9 * 12 && n > 0);
        std::vector<ValueType> vp(9 * 12), rho(9 * 12);
        for (size_t i = 0; i < vp.size(); i++) {
            vp[i] = 1500 + (ValueType)i;
            rho[i] = 2000 - (ValueType)i;
        }
        IO::writeVector(vp, dir + "/vgmodel.vp", 1);
        IO::writeVector(rho, dir + "/vgmodel.density", 1);
        Modelparameter::Modelparameter<ValueType> m("acoustic");
        m.init(c, coords);
        bool ok = (IndexType)m.getVelocityP().size() == n;
        for (IndexType i = 0; ok && i < n; i++) {
            const IndexType r = regular.coordinate2index(coords.index2coordinate(i));
            ok = m.getVelocityP()[i] == vp[r] && m.getDensity()[i] == rho[r];
        }
        EXPECT(ok);
        // written on the variable grid and read back as such
        m.write(dir + "/vgmodel2", 1);
        {
            std::ofstream c2(dir + "/vg2.txt");
            c2 << "dimension=2D\nequationType=acoustic\nNX=9\nNY=12\nNZ=1\nDH=10\nUseVariableGrid=1\nfileFormat=1\nModelRead=2\nModelFilename=" << dir << "/vgmodel2\n"
               << "gridConfigurationFilename=" << dir << "/vgGrid.txt\n";
            std::ofstream c3(dir + "/vg3.txt");
            c3 << "dimension=2D\nequationType=acoustic\nNX=9\nNY=12\nNZ=1\nDH=10\nUseVariableGrid=0\nfileFormat=1\nModelRead=2\nModelFilename=" << dir << "/vgmodel2\n";
        }
        Configuration::Configuration c2(dir + "/vg2.txt"), c3(dir + "/vg3.txt");
        Modelparameter::Modelparameter<ValueType> m2("acoustic"), m3("acoustic");
        m2.init(c2, coords);
        EXPECT(m2.getVelocityP() == m.getVelocityP() && m2.getDensity() == m.getDensity());
        EXPECT_THROW(m3.init(c3, regular)); // "Read variable model (ModelRead=2) not available if regular grid is chosen!"
    }
    // ---- Seismogram::write / read: common-offset profile rows, the inverse AGC function (Seismogram.cpp:82-207)
    {
        Acquisition::Seismogram<ValueType> one;
        one.setTraceType(Acquisition::VY, true);
        one.getCoordinates1D().assign(1, 7);
        one.allocate(1, 6);
        one.setDT(0.1f);
        one.setSeismoDT(0.1f);
        one.allocateCOP(3, 6);
        for (IndexType shot = 0; shot < 3; shot++) { // three single-trace shots land in rows 0..2 of the profile, no per-shot file
            for (IndexType k = 0; k < 6; k++)
                one.getData()[k] = (ValueType)(10 * shot + k);
            one.setShotInd(shot, 2 - shot);
            one.write(1, dir + "/cop.shot_" + std::to_string(shot));
        }
        EXPECT(!std::ifstream(dir + "/cop.shot_1.vy.mtx").good());
        Acquisition::Seismogram<ValueType> all(one);
        all.assignCOP();
        EXPECT(all.getNumTraces() == 3 && all.getData()[6 + 2] == 12);
        all.write(1, dir + "/cop");
        // a single-trace shot reads its row back: row shotInd, or row shotIndIncr of the original numbering
        one.setShotInd(1, 2);
        one.read(1, dir + "/cop.shot_1");
        EXPECT(one.getNumTraces() == 1 && one.getData()[3] == 13);
        one.read(1, dir + "/cop.shot_1", true);
        EXPECT(one.getData()[3] == 23);
        // the inverse AGC function travels through format 5 and is applied by the next normalizeTrace(3)
        Acquisition::Seismogram<ValueType> g;
        g.setTraceType(Acquisition::P, true);
        g.getCoordinates1D().assign(2, 3);
        g.allocate(2, 40);
        g.setDT(0.01f);
        g.setSeismoDT(0.01f);
        for (IndexType k = 0; k < 80; k++)
            g.getData()[k] = std::sin(0.3f * k) * (1 + k % 7);
        g.setFrequencyAGC(10);
        g.calcInverseAGC();
        const std::vector<ValueType> gain = g.getInverseAGC();
        g.write(5, dir + "/agc");
        Acquisition::Seismogram<ValueType> h(g);
        std::fill(h.getInverseAGC().begin(), h.getInverseAGC().end(), ValueType(0));
        h.read(5, dir + "/agc");
        double worst = 0;
        for (size_t k = 0; k < gain.size(); k++)
            worst = std::max(worst, std::abs((double)h.getInverseAGC()[k] / gain[k] - 1));
        EXPECT(gain.size() == 80 && worst < 1e-6);
        g.normalizeTrace(3);
        h.normalizeTrace(3);
        worst = 0;
        for (size_t k = 0; k < 80; k++)
            worst = std::max(worst, std::abs((double)h.getData()[k] - g.getData()[k]));
        EXPECT(worst < 1e-5);
    }
    // ---- file formats: mtx and lmf round trips, resampling (Common.hpp:202-233)
    {
        std::vector<ValueType> m = {1, 2, 3, 4, 5, 6}, back;
        IndexType r, c;
        for (IndexType fmt : {1, 2}) {
            IO::writeMatrix(m, 2, 3, dir + "/mat", fmt);
            IO::readMatrix(back, r, c, dir + "/mat", fmt);
            EXPECT(r == 2 && c == 3 && back == m);
            IO::writeVector(m, dir + "/vec", fmt);
            std::vector<ValueType> v(6);
            IO::readVector(v, dir + "/vec", fmt);
            EXPECT(v == m);
            std::vector<ValueType> w(5);
            EXPECT_THROW(IO::readVector(w, dir + "/vec", fmt));
        }
        EXPECT_THROW(IO::writeVector(m, dir + "/vec", 3));
        std::vector<ValueType> d = {0, 1, 2, 3, 4, 10, 11, 12, 13, 14};
        IndexType nn;
        Common::resampleRows(d, 2, 5, 2.0f, nn);
        EXPECT(nn == 3 && d[0] == 0 && d[1] == 2 && d[2] == 4 && d[3] == 10 && d[5] == 14);
        std::vector<ValueType> e = {0, 1, 2, 3, 4};
        Common::resampleRows(e, 1, 5, 1.5f, nn);
        EXPECT(nn == 3 && e[1] == 1.5f && e[2] == 3.0f);
    }
    // ---- derivative descriptor (Derivatives.cpp:2001-2042)
    {
        auto c = ForwardSolver::Derivatives::Derivatives<ValueType>::calcFDCoef(4);
        EXPECT(c.size() == 4 && c[0] == (ValueType)(1.0 / 24.0) && c[1] == (ValueType)(-9.0 / 8.0) && c[2] == (ValueType)(9.0 / 8.0));
        for (IndexType q = 2; q <= 12; q += 2) {
            auto k = ForwardSolver::Derivatives::Derivatives<ValueType>::calcFDCoef(q);
            double s = 0;
            for (IndexType j = 0; j < q; j++)
                s += k[j] * (j - q / 2 + 0.5); // exact for f(x) = x
            EXPECT(std::abs(s - 1.0) < 1e-6);
        }
        EXPECT_THROW(ForwardSolver::Derivatives::Derivatives<ValueType>::calcFDCoef(7));
        EXPECT_THROW(ForwardSolver::Derivatives::Factory<ValueType>::Create("4D"));
    }
    std::printf(failures ? "%d FAILURES\n" : "host unit tests OK\n", failures);
    return failures ? 1 : 0;
}

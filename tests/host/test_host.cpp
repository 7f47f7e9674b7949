// Unit tests of the LAMA-free host layer (wave-simulation_b200/host), modelled on the reference's googletest cases
// (src/Tests/UnitTest/{Configuration,Coordinates,Ricker,...}UnitTest.cpp): same known answers, plain asserts.
// Built and run by tests/test_host_layer.py; needs no GPU and no solver library.
#include "Acquisition.hpp"
#include "Configuration.hpp"
#include "Coordinates.hpp"
#include "Derivatives.hpp"
#include "IO.hpp"
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

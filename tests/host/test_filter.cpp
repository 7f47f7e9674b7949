// Filter / Hilbert tools of the host layer (wave-simulation_b200/host/Filter.hpp).  Reads a dense row-major matrix of traces from a raw
// float32 file, applies the requested operation and writes the result next to it; tests/test_host_layer.py compares with the
// reference's own fixture (src/Tests/UnitTest/FilterUnitTest.cpp: filterTest_signal.mtx -> filterTest_signalFiltRef.mtx, gate
// l2 < 0.005) and with numpy / scipy restatements.
// usage: test_filter <in.f32> <out.f32> <rows> <nt> <dt> <op> [family type order fc1 fc2]
#include "Acquisition.hpp"
#include "Filter.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>

using namespace KITGPI;

int main(int argc, char **argv)
{
    if (argc < 7)
        return 2;
    try {
        const IndexType rows = std::atoi(argv[3]), nt = std::atoi(argv[4]);
        const float dt = (float)std::atof(argv[5]);
        const std::string op = argv[6];
        std::vector<float> data((size_t)rows * nt);
        std::ifstream in(argv[1], std::ios::binary);
        in.read(reinterpret_cast<char *>(data.data()), (std::streamsize)(data.size() * sizeof(float)));
        if (op == "filter") {
            Filter::Filter<float> f;
            f.init(dt, nt);
            f.calc(argv[7], argv[8], std::atoi(argv[9]), (float)std::atof(argv[10]), argc > 11 ? (float)std::atof(argv[11]) : 0.0f);
            // through the seismogram classes, as WAVE-Inversion does (SeismogramHandler::filter -> Seismogram::filterTraces)
            Acquisition::SeismogramHandler<float> h;
            h.getSeismogram(2).allocate(rows, nt);
            h.getSeismogram(2).getCoordinates1D().assign(rows, 0);
            h.getSeismogram(2).getData() = data;
            h.filter(f);
            data = h.getSeismogram(2).getData();
        } else if (op == "filter1") { // trace by trace through apply(vector)
            Filter::Filter<float> f;
            f.init(dt, nt);
            f.calc(argv[7], argv[8], std::atoi(argv[9]), (float)std::atof(argv[10]), argc > 11 ? (float)std::atof(argv[11]) : 0.0f);
            for (IndexType r = 0; r < rows; r++) {
                std::vector<float> row(data.begin() + (size_t)r * nt, data.begin() + (size_t)(r + 1) * nt);
                f.apply(row);
                std::copy(row.begin(), row.end(), data.begin() + (size_t)r * nt);
            }
        } else if (op == "hilbert") {
            Hilbert::HilbertFFT<float> h;
            h.setCoefficientLength(Common::calcNextPowTwo<float>(nt - 1)); // Simulation.cpp:302-304
            h.calcHilbertCoefficient();
            h.hilbert(data, rows, nt);
        } else
            return 2;
        std::ofstream out(argv[2], std::ios::binary);
        out.write(reinterpret_cast<const char *>(data.data()), (std::streamsize)(data.size() * sizeof(float)));
    } catch (std::exception const &e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    return 0;
}

// TwoLayer — model creation tool of the reference (mirror of src/Tools/CreateModel/TwoLayer.cpp:25-89): reads the grid and
// the model file name from a configuration file and writes a two-layer model (interface at y = 40) in the configured
// file format.  Usage: TwoLayer <configuration file>.  Host-only: it does not touch the CUDA library.
#include "../Configuration.hpp"
#include "../IO.hpp"
#include <iostream>

using namespace KITGPI;

int main(int argc, char *argv[])
{
    // upper / lower layer (TwoLayer.cpp:27-28) and the depth of the second layer (:31)
    const ValueType vp1 = 3500, vs1 = 2000, rho1 = 2000, tauP1 = 0.1, tauS1 = 0.1;
    const ValueType vp2 = 4550, vs2 = 2600, rho2 = 2600, tauP2 = 0.1, tauS2 = 0.1;
    const IndexType depth = 40;
    if (argc != 2) {
        std::cout << "\n\nNo configuration file given!\n\n" << std::endl;
        return 2;
    }
    try {
        Configuration::Configuration config(argv[1]);
        const IndexType NX = config.get<IndexType>("NX"), NY = config.get<IndexType>("NY"), NZ = config.get<IndexType>("NZ");
        const size_t plane = (size_t)NX * NZ, n = plane * NY; // linear index x + z NX + y NX NZ: whole planes per depth
        auto layered = [&](ValueType top, ValueType bottom) {
            std::vector<ValueType> v(n, top);
            for (IndexType y = depth; y < NY; ++y)
                std::fill(v.begin() + (size_t)y * plane, v.begin() + (size_t)(y + 1) * plane, bottom);
            return v;
        };
        std::string type = config.get<std::string>("equationType");
        const std::string filename = config.get<std::string>("ModelFilename");
        const IndexType fileFormat = config.get<IndexType>("FileFormat");
        IO::writeVector(layered(rho1, rho2), filename + ".density", fileFormat);
        if (type.compare("sh") != 0)
            IO::writeVector(layered(vp1, vp2), filename + ".vp", fileFormat);
        if (type.compare("acoustic") != 0)
            IO::writeVector(layered(vs1, vs2), filename + ".vs", fileFormat);
        if (type.compare("viscoelastic") == 0) {
            IO::writeVector(layered(tauP1, tauP2), filename + ".tauP", fileFormat);
            IO::writeVector(layered(tauS1, tauS2), filename + ".tauS", fileFormat);
        }
    } catch (std::exception const &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}

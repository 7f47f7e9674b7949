#include "Configuration.hpp"
#include <fstream>

namespace KITGPI
{
    int verbose = 0;
}

void KITGPI::Configuration::Configuration::add2map(std::string const &KEY, std::string const &VALUE, bool overwrite)
{
    auto it = configMap.find(KEY);
    if (it == configMap.end()) {
        configMap.emplace(KEY, VALUE);
        insertionOrder.push_back(KEY);
    } else if (overwrite) {
        it->second = VALUE;
    }
}

void KITGPI::Configuration::Configuration::readFromFile(std::string const &filename, bool overwrite)
{
    std::ifstream input(filename.c_str());
    if (!input.good())
        COMMON_THROWEXCEPTION("Configuration file " << filename << " was not found " << std::endl)
    std::string line;
    bool flag2D = false;
    while (std::getline(input, line)) {
        size_t lineEnd = line.size();
        const size_t hash = line.find('#');
        if (hash == 0)
            continue; // whole-line comment
        if (hash != std::string::npos)
            lineEnd = hash; // trailing comment
        const size_t eq = line.find('=');
        if (eq == std::string::npos)
            continue;
        std::string name = lower(line.substr(0, eq));
        // the value keeps its blanks (operator>> skips them on conversion); lineEnd < eq+1 happens for `# a=b` only
        std::string val = lineEnd > eq ? line.substr(eq + 1, lineEnd - (eq + 1)) : std::string();
        if (name == "dimension" && val == "2D              ") // sic: Configuration.cpp:77
            flag2D = true;
        if (name == "nz" && flag2D)
            val = "1";
        add2map(name, val, overwrite);
    }
}

void KITGPI::Configuration::Configuration::print() const
{
    std::cout << "\t"
              << "Configuration: \n";
    for (auto const &k : insertionOrder)
        std::cout << "\t" << k << " = " << configMap.at(k) << std::endl;
    std::cout << std::endl;
}

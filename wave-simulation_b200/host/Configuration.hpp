// Configuration.hpp — `KEY=VALUE` parameter files (mirror of src/Configuration/Configuration.hpp:27-54).
// Same behaviour: case-insensitive keys, `#` comments (whole line or trailing), first occurrence wins unless
// overwrite, typed get<T>() that throws "Parameter X: Not found in Configuration file!", getAndCatch<T>() with default,
// and the 2D quirk of Configuration.cpp:77-82 (NZ forced to 1 when `dimension=2D` is followed by exactly 14 blanks).
#pragma once
#include "Common.hpp"
#include <algorithm>
#include <list>
#include <unordered_map>

namespace KITGPI
{
    namespace Configuration
    {
        class Configuration
        {
          public:
            Configuration() {}
            explicit Configuration(std::string const &filename) { readFromFile(filename); }

            void readFromFile(std::string const &filename, bool overwrite = false);
            void print() const;

            template <typename ReturnType> ReturnType get(std::string const &parameterName) const
            {
                ReturnType temp;
                auto it = configMap.find(lower(parameterName));
                if (it == configMap.end())
                    COMMON_THROWEXCEPTION("Parameter " << parameterName << ": Not found in Configuration file! " << std::endl)
                std::istringstream input(it->second);
                input >> temp;
                return temp;
            }
            template <typename ReturnType> ReturnType getAndCatch(std::string const &parameterName, ReturnType parameterValue) const
            {
                auto it = configMap.find(lower(parameterName));
                if (it == configMap.end())
                    return parameterValue;
                ReturnType temp;
                std::istringstream input(it->second);
                input >> temp;
                return temp;
            }
            template <typename InputType> void add2config(std::string const &KEY, InputType const &VALUE, bool overwrite = false)
            {
                std::ostringstream sstream;
                sstream << VALUE;
                add2map(lower(KEY), sstream.str(), overwrite);
            }
            bool has(std::string const &parameterName) const { return configMap.count(lower(parameterName)) != 0; }

          private:
            static std::string lower(std::string s)
            {
                std::transform(s.begin(), s.end(), s.begin(), ::tolower);
                return s;
            }
            void add2map(std::string const &KEY, std::string const &VALUE, bool overwrite = false);
            std::unordered_map<std::string, std::string> configMap;
            std::list<std::string> insertionOrder;
        };
    }
}

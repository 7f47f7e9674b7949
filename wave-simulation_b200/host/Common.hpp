// Common.hpp — helpers shared by the LAMA-free host layer (mirror of src/Common/Common.hpp, HostPrint.hpp).
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace KITGPI
{
    typedef float ValueType;   // src/Configuration/ValueType.hpp:5
    typedef int32_t IndexType; // scai::IndexType

    //! The reference aborts with a message through COMMON_THROWEXCEPTION; here a std::runtime_error carries it.
    struct Exception : public std::runtime_error {
        explicit Exception(const std::string &m) : std::runtime_error(m) {}
    };
    extern int verbose; // global switch of Simulation.cpp:34
}

#define COMMON_THROWEXCEPTION(msg)                                                                                     \
    {                                                                                                                  \
        std::ostringstream oss__;                                                                                     \
        oss__ << msg;                                                                                                  \
        throw KITGPI::Exception(oss__.str());                                                                          \
    }
#define SCAI_ASSERT_ERROR(cond, msg)                                                                                   \
    if (!(cond))                                                                                                       \
    COMMON_THROWEXCEPTION(msg)

// HOST_PRINT(msg) / HOST_PRINT(msg, verboseMsg): rank-0 style printing (Common/HostPrint.hpp)
#define HOST_PRINT1(msg) std::cout << msg << std::flush;
#define HOST_PRINT2(msg, vmsg)                                                                                         \
    {                                                                                                                  \
        std::cout << msg;                                                                                              \
        if (KITGPI::verbose)                                                                                           \
            std::cout << vmsg;                                                                                         \
        std::cout << std::flush;                                                                                       \
    }
#define HOST_PRINT_SEL(_1, _2, NAME, ...) NAME
#define HOST_PRINT(...) HOST_PRINT_SEL(__VA_ARGS__, HOST_PRINT2, HOST_PRINT1)(__VA_ARGS__)

namespace KITGPI
{
    namespace Common
    {
        //! number of time steps of a continuous time: Common.hpp:240-243
        inline IndexType time2index(ValueType time, ValueType DT) { return static_cast<IndexType>(time / DT + 0.5); }

        //! Linear-interpolation resampling of the columns of a row-major matrix (calcResampleMat, Common.hpp:202-233)
        void resampleRows(std::vector<ValueType> &data, IndexType numRows, IndexType numCols, ValueType resamplingCoeff, IndexType &numColsNew);

        //! dimension / equationType sanity checks (Common.hpp:249 checkEquationType)
        bool checkEquationType(std::string type); // true = seismic, false = EM, throws if unknown
    }
}

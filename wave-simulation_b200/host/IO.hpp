// IO.hpp — vector / matrix files of the reference (mirror of src/IO/IO.hpp:21-195).
//   fileFormat 1 = .mtx  MatrixMarket "array" (ASCII, column-major values, as lama writeToFile FORMATTED)
//   fileFormat 2 = .lmf  LAMA binary: int32 {0x4711E01, 0, 2}, int32 ndims, int32 sizes[ndims], float32 LE row-major
//                        (documented by par/model/readVectorfromLMF.m:6-25 and par/seismograms/readSeismogram.m:33-54)
//   fileFormat 3 = .frv  LAMA binary + separate header: not available here (throws)
#pragma once
#include "Common.hpp"

namespace KITGPI
{
    namespace IO
    {
        std::string suffix(IndexType fileFormat);
        void writeVector(std::vector<ValueType> const &vector, std::string filename, IndexType fileFormat);
        //! the size of `vector` must be set before the call (the reference asserts the file holds as many values)
        void readVector(std::vector<ValueType> &vector, std::string filename, IndexType fileFormat);
        //! row-major matrix numRows x numCols
        void writeMatrix(std::vector<ValueType> const &matrix, IndexType numRows, IndexType numCols, std::string filename, IndexType fileFormat);
        void readMatrix(std::vector<ValueType> &matrix, IndexType &numRows, IndexType &numCols, std::string filename, IndexType fileFormat);
    }
}

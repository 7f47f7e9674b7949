// IO.hpp — vector / matrix files of the reference (mirror of src/IO/IO.hpp:21-195).
//   fileFormat 1 = .mtx  MatrixMarket "array" (ASCII, column-major values, as lama writeToFile FORMATTED)
//   fileFormat 2 = .lmf  LAMA binary: int32 {0x4711E01, 0, 2}, int32 ndims, int32 sizes[ndims], float32 LE row-major
//                        (documented by par/model/readVectorfromLMF.m:6-25 and par/seismograms/readSeismogram.m:33-54)
//   fileFormat 3 = .frv  LAMA binary + separate header: not available here (throws)
//   SeismogramFormat 4 = .su  Seismic Unix: per trace a 240-byte SEG-Y trace header + ns native-endian floats
//                        (mirror of src/IO/SUIO.hpp:156-260; header layout of src/Acquisition/segy.hpp = CWP/SU segy.h)
#pragma once
#include "Common.hpp"
#include "Coordinates.hpp"

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
    namespace SUIO
    {
        //! traces = rows of the row-major matrix `data` (ntr x ns); header words as SUIO.hpp:196-246 (coordinates in mm, scalco = -3)
        void writeSU(std::string const &filename, std::vector<ValueType> const &data, IndexType ntr, IndexType ns, std::vector<IndexType> const &coordinates1D, ValueType DT,
                     IndexType sourceCoordinate1D, Acquisition::Coordinates<ValueType> const &modelCoordinates);
        //! trace data without the headers (SUIO.hpp:262-300); ns / ntr are taken from the first header and the file size
        void readDataSU(std::string const &filename, std::vector<ValueType> &data, IndexType &ntr, IndexType &ns);
        //! one header word of trace `trace` by its SU keyword: tracl offset gelev sdepth sx sy gx gy ns dt scalco ntr ... (for tests / suHandler)
        double readHeaderWordSU(std::string const &filename, IndexType trace, std::string const &key);
        //! number of traces of <filename>.su (0 if the file does not exist)
        IndexType numTracesSU(std::string const &filename);
    }
}

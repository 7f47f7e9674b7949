#include "IO.hpp"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>

using namespace KITGPI;

namespace
{
    const int32_t kLmfId = 0x4711E01;

    void writeLmf(const std::string &filename, const float *data, const std::vector<int32_t> &sizes)
    {
        std::ofstream out(filename, std::ios::binary);
        if (!out.good())
            COMMON_THROWEXCEPTION("Could not open " << filename << " for writing")
        const int32_t header[3] = {kLmfId, 0, 2};
        const int32_t ndims = (int32_t)sizes.size();
        out.write(reinterpret_cast<const char *>(header), sizeof(header));
        out.write(reinterpret_cast<const char *>(&ndims), sizeof(ndims));
        out.write(reinterpret_cast<const char *>(sizes.data()), sizeof(int32_t) * sizes.size());
        size_t n = 1;
        for (int32_t s : sizes)
            n *= (size_t)s;
        out.write(reinterpret_cast<const char *>(data), sizeof(float) * n);
    }

    std::vector<int32_t> readLmf(const std::string &filename, std::vector<float> &data, int ndimsExpected)
    {
        std::ifstream in(filename, std::ios::binary);
        if (!in.good())
            COMMON_THROWEXCEPTION("Could not open " << filename)
        int32_t header[3], ndims = 0;
        in.read(reinterpret_cast<char *>(header), sizeof(header));
        in.read(reinterpret_cast<char *>(&ndims), sizeof(ndims));
        SCAI_ASSERT_ERROR(in.good() && header[0] == kLmfId, "HEADER of " << filename << " is not 4711E01 (dense array)")
        SCAI_ASSERT_ERROR(header[1] == 0 && header[2] == 2, filename << ": index type must be int and value type float")
        SCAI_ASSERT_ERROR(ndims == ndimsExpected, filename << ": NDIMS=" << ndims << " must be " << ndimsExpected)
        std::vector<int32_t> sizes(ndims);
        in.read(reinterpret_cast<char *>(sizes.data()), sizeof(int32_t) * ndims);
        size_t n = 1;
        for (int32_t s : sizes)
            n *= (size_t)s;
        data.resize(n);
        in.read(reinterpret_cast<char *>(data.data()), sizeof(float) * n);
        SCAI_ASSERT_ERROR(in.gcount() == (std::streamsize)(sizeof(float) * n), filename << " is truncated")
        return sizes;
    }

    // MatrixMarket array: header line, "rows cols" line, values column by column
    void writeMtx(const std::string &filename, const float *rowMajor, IndexType numRows, IndexType numCols)
    {
        FILE *f = std::fopen(filename.c_str(), "w");
        if (!f)
            COMMON_THROWEXCEPTION("Could not open " << filename << " for writing")
        std::fprintf(f, "%%%%MatrixMarket matrix array real general\n%d %d\n", numRows, numCols);
        for (IndexType c = 0; c < numCols; c++)
            for (IndexType r = 0; r < numRows; r++)
                std::fprintf(f, "%.9g\n", (double)rowMajor[(size_t)r * numCols + c]);
        std::fclose(f);
    }

    void readMtx(const std::string &filename, std::vector<float> &rowMajor, IndexType &numRows, IndexType &numCols)
    {
        std::ifstream in(filename);
        if (!in.good())
            COMMON_THROWEXCEPTION("Could not open " << filename)
        std::string line;
        bool vectorHeader = false;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] != '%')
                break;
            if (line.find("vector") != std::string::npos)
                vectorHeader = true;
        }
        std::istringstream sz(line);
        numRows = numCols = 0;
        sz >> numRows;
        if (!(sz >> numCols) || vectorHeader)
            numCols = 1;
        SCAI_ASSERT_ERROR(numRows > 0 && numCols > 0, filename << ": bad MatrixMarket size line")
        rowMajor.assign((size_t)numRows * numCols, 0.0f);
        for (IndexType c = 0; c < numCols; c++)
            for (IndexType r = 0; r < numRows; r++) {
                double v;
                if (!(in >> v))
                    COMMON_THROWEXCEPTION(filename << " holds fewer than " << numRows * numCols << " values")
                rowMajor[(size_t)r * numCols + c] = (float)v;
            }
    }
}

std::string IO::suffix(IndexType fileFormat)
{
    switch (fileFormat) {
    case 1: return ".mtx";
    case 2: return ".lmf";
    case 3: COMMON_THROWEXCEPTION("fileFormat 3 (.frv, LAMA binary with separate header) is not available in the B200 host layer")
    default: COMMON_THROWEXCEPTION("Unexpected fileFormat option!")
    }
}

void IO::writeVector(std::vector<ValueType> const &vector, std::string filename, IndexType fileFormat)
{
    filename += suffix(fileFormat);
    HOST_PRINT("", "writing " << filename << "\n")
    if (fileFormat == 1)
        writeMtx(filename, vector.data(), (IndexType)vector.size(), 1);
    else
        writeLmf(filename, vector.data(), {(int32_t)vector.size()});
}

void IO::readVector(std::vector<ValueType> &vector, std::string filename, IndexType fileFormat)
{
    filename += suffix(fileFormat);
    HOST_PRINT("", "reading " << filename << "\n")
    const size_t expected = vector.size();
    if (fileFormat == 1) {
        IndexType r, c;
        readMtx(filename, vector, r, c);
    } else
        readLmf(filename, vector, 1);
    SCAI_ASSERT_ERROR(vector.size() == expected, "Read " << vector.size() << " elements from file: " << filename << ", expected " << expected << " elements!")
}

void IO::writeMatrix(std::vector<ValueType> const &matrix, IndexType numRows, IndexType numCols, std::string filename, IndexType fileFormat)
{
    filename += suffix(fileFormat);
    SCAI_ASSERT_ERROR(matrix.size() == (size_t)numRows * numCols, "matrix size mismatch")
    if (fileFormat == 1)
        writeMtx(filename, matrix.data(), numRows, numCols);
    else
        writeLmf(filename, matrix.data(), {numRows, numCols});
    HOST_PRINT("", "writing " << filename << "\n")
}

void IO::readMatrix(std::vector<ValueType> &matrix, IndexType &numRows, IndexType &numCols, std::string filename, IndexType fileFormat)
{
    filename += suffix(fileFormat);
    HOST_PRINT("", "reading " << filename << "\n")
    if (fileFormat == 1)
        readMtx(filename, matrix, numRows, numCols);
    else {
        std::vector<int32_t> s = readLmf(filename, matrix, 2);
        numRows = s[0];
        numCols = s[1];
    }
}

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
        bool vectorHeader = false, coordinate = false;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] != '%')
                break;
            if (line.find("vector") != std::string::npos)
                vectorHeader = true;
            if (line.find("coordinate") != std::string::npos)
                coordinate = true;
        }
        std::istringstream sz(line);
        numRows = numCols = 0;
        sz >> numRows;
        if (!(sz >> numCols) || vectorHeader)
            numCols = 1;
        SCAI_ASSERT_ERROR(numRows > 0 && numCols > 0, filename << ": bad MatrixMarket size line")
        rowMajor.assign((size_t)numRows * numCols, 0.0f);
        if (coordinate) { // sparse: "rows cols nnz", then one "i j value" per entry (1-based)
            long long nnz = 0;
            sz >> nnz;
            for (long long e = 0; e < nnz; e++) {
                long long i, j;
                double v;
                if (!(in >> i >> j >> v))
                    COMMON_THROWEXCEPTION(filename << " holds fewer than " << nnz << " entries")
                SCAI_ASSERT_ERROR(i >= 1 && i <= numRows && j >= 1 && j <= numCols, filename << ": entry outside the matrix")
                rowMajor[(size_t)(i - 1) * numCols + (j - 1)] = (float)v;
            }
            return;
        }
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

// ---------------------------------------------------------------------------------------------------------------------
// Seismic Unix traces (src/IO/SUIO.hpp).  The 240-byte trace header is written field by field at the byte offsets of the
// CWP/SU `segy` struct (src/Acquisition/segy.hpp): i = int32, h = int16, u = uint16, f = float32.
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
    struct SuKey {
        const char *name;
        int offset;
        char type;
    };
    const SuKey kSuKeys[] = {{"tracl", 0, 'i'},   {"tracr", 4, 'i'},   {"fldr", 8, 'i'},     {"tracf", 12, 'i'},  {"ep", 16, 'i'},      {"cdp", 20, 'i'},
                             {"cdpt", 24, 'i'},   {"trid", 28, 'h'},   {"nvs", 30, 'h'},     {"nhs", 32, 'h'},    {"duse", 34, 'h'},    {"offset", 36, 'i'},
                             {"gelev", 40, 'i'},  {"selev", 44, 'i'},  {"sdepth", 48, 'i'},  {"gdel", 52, 'i'},   {"sdel", 56, 'i'},    {"swdep", 60, 'i'},
                             {"gwdep", 64, 'i'},  {"scalel", 68, 'h'}, {"scalco", 70, 'h'},  {"sx", 72, 'i'},     {"sy", 76, 'i'},      {"gx", 80, 'i'},
                             {"gy", 84, 'i'},     {"counit", 88, 'h'}, {"ns", 114, 'u'},     {"dt", 116, 'u'},    {"d1", 180, 'f'},     {"ntr", 204, 'i'}};
    const SuKey &suKey(const char *name) // a reference into the static table above
    {
        for (auto const &k : kSuKeys)
            if (std::strcmp(name, k.name) == 0)
                return k;
        COMMON_THROWEXCEPTION("unknown SU header word " << name)
    }
    void suPut(unsigned char *hdr, const char *name, double v)
    {
        const SuKey &k = suKey(name);
        switch (k.type) {
        case 'i': { int32_t w = (int32_t)v; std::memcpy(hdr + k.offset, &w, 4); break; }
        case 'h': { int16_t w = (int16_t)v; std::memcpy(hdr + k.offset, &w, 2); break; }
        case 'u': { uint16_t w = (uint16_t)v; std::memcpy(hdr + k.offset, &w, 2); break; }
        default: { float w = (float)v; std::memcpy(hdr + k.offset, &w, 4); break; }
        }
    }
    double suGet(const unsigned char *hdr, const SuKey &k)
    {
        switch (k.type) {
        case 'i': { int32_t w; std::memcpy(&w, hdr + k.offset, 4); return w; }
        case 'h': { int16_t w; std::memcpy(&w, hdr + k.offset, 2); return w; }
        case 'u': { uint16_t w; std::memcpy(&w, hdr + k.offset, 2); return w; }
        default: { float w; std::memcpy(&w, hdr + k.offset, 4); return w; }
        }
    }
}

void KITGPI::SUIO::writeSU(std::string const &filename, std::vector<ValueType> const &data, IndexType ntr, IndexType ns, std::vector<IndexType> const &coordinates1D,
                           ValueType DT, IndexType sourceCoordinate1D, Acquisition::Coordinates<ValueType> const &modelCoordinates)
{
    SCAI_ASSERT_ERROR((IndexType)coordinates1D.size() == ntr && (IndexType)data.size() == ntr * ns, "writeSU: " << ntr << " traces of " << ns << " samples expected")
    SCAI_ASSERT_ERROR(ns <= 65535, "writeSU: the SU header holds at most 65535 samples per trace")
    const std::string name = filename + ".su";
    std::ofstream out(name, std::ios::binary);
    SCAI_ASSERT_ERROR(out.good(), "Could not open " << name)
    const ValueType DH = modelCoordinates.getDH();
    const Acquisition::coordinate3D src = modelCoordinates.index2coordinate(sourceCoordinate1D);
    const ValueType XS = src.x * DH, YS = src.y * DH, ZS = src.z * DH;
    const ValueType xshift = 800.0, yshift = 800.0; // SUIO.hpp:203
    const ValueType dtms = (ValueType)(DT * 1000000);
    for (IndexType tr = 0; tr < ntr; tr++) {
        unsigned char hdr[240];
        std::memset(hdr, 0, sizeof(hdr));
        const Acquisition::coordinate3D rec = modelCoordinates.index2coordinate(coordinates1D[tr]);
        const ValueType xr = rec.x * DH, yr = rec.y * DH, zr = rec.z * DH;
        const ValueType x = xr - XS, y = yr - YS, z = zr - ZS; // source position as reference point
        suPut(hdr, "counit", 1);
        suPut(hdr, "ntr", ntr);
        suPut(hdr, "tracl", tr + 1);
        suPut(hdr, "tracr", 1);
        suPut(hdr, "ep", 1);
        suPut(hdr, "cdp", ntr);
        suPut(hdr, "trid", 1);
        suPut(hdr, "offset", std::round(std::sqrt((XS - xr) * (XS - xr) + (YS - yr) * (YS - yr) + (ZS - zr) * (ZS - zr)) * 1000.0));
        suPut(hdr, "gelev", std::round(yr * 1000.0));
        suPut(hdr, "sdepth", std::round(YS * 1000.0));
        suPut(hdr, "gdel", std::round(std::atan2(-y, z) * 180 * 1000.0 / 3.1415926));
        suPut(hdr, "gwdep", std::round(std::sqrt(z * z + y * y) * 1000.0));
        suPut(hdr, "swdep", std::round(((360.0 / (2.0 * 3.1415926)) * std::atan2(x - xshift, y - yshift)) * 1000.0));
        suPut(hdr, "scalel", -3);
        suPut(hdr, "scalco", -3);
        suPut(hdr, "sx", std::round(XS * 1000.0));
        suPut(hdr, "sy", std::round(ZS * 1000.0));
        suPut(hdr, "gx", std::round(xr * 1000.0));
        suPut(hdr, "gy", std::round(zr * 1000.0));
        suPut(hdr, "ns", ns);
        suPut(hdr, "dt", std::round(dtms));
        suPut(hdr, "d1", (float)(uint16_t)std::round(dtms) * 1.0e-6);
        out.write(reinterpret_cast<const char *>(hdr), 240);
        out.write(reinterpret_cast<const char *>(data.data() + (size_t)tr * ns), sizeof(float) * ns);
    }
    SCAI_ASSERT_ERROR(out.good(), "Could not write " << name)
}

void KITGPI::SUIO::readDataSU(std::string const &filename, std::vector<ValueType> &data, IndexType &ntr, IndexType &ns)
{
    const std::string name = filename + ".su";
    std::ifstream in(name, std::ios::binary | std::ios::ate);
    SCAI_ASSERT_ERROR(in.good(), "Could not open " << name)
    const std::streamoff size = in.tellg();
    in.seekg(0);
    unsigned char hdr[240];
    in.read(reinterpret_cast<char *>(hdr), 240);
    SCAI_ASSERT_ERROR(in.good(), name << " holds no SU trace")
    ns = (IndexType)suGet(hdr, suKey("ns"));
    const std::streamoff rec = 240 + (std::streamoff)sizeof(float) * ns;
    SCAI_ASSERT_ERROR(ns > 0 && size % rec == 0, name << " is not a sequence of SU traces with " << ns << " samples")
    ntr = (IndexType)(size / rec);
    data.resize((size_t)ntr * ns);
    for (IndexType tr = 0; tr < ntr; tr++) {
        in.seekg((std::streamoff)tr * rec + 240);
        in.read(reinterpret_cast<char *>(data.data() + (size_t)tr * ns), sizeof(float) * ns);
    }
    SCAI_ASSERT_ERROR(in.good(), "Could not read " << name)
}

double KITGPI::SUIO::readHeaderWordSU(std::string const &filename, IndexType trace, std::string const &key)
{
    const std::string name = filename + ".su";
    std::ifstream in(name, std::ios::binary);
    SCAI_ASSERT_ERROR(in.good(), "Could not open " << name)
    unsigned char hdr[240];
    in.read(reinterpret_cast<char *>(hdr), 240);
    SCAI_ASSERT_ERROR(in.good(), name << " holds no SU trace")
    const IndexType ns = (IndexType)suGet(hdr, suKey("ns"));
    in.seekg((std::streamoff)trace * (240 + (std::streamoff)sizeof(float) * ns));
    in.read(reinterpret_cast<char *>(hdr), 240);
    SCAI_ASSERT_ERROR(in.good(), name << " has no trace " << trace)
    return suGet(hdr, suKey(key.c_str()));
}

IndexType KITGPI::SUIO::numTracesSU(std::string const &filename)
{
    const std::string name = filename + ".su";
    std::ifstream in(name, std::ios::binary | std::ios::ate);
    if (!in.good())
        return 0;
    const std::streamoff size = in.tellg();
    if (size < 240)
        return 0;
    in.seekg(0);
    unsigned char hdr[240];
    in.read(reinterpret_cast<char *>(hdr), 240);
    const IndexType ns = (IndexType)suGet(hdr, suKey("ns"));
    const std::streamoff rec = 240 + (std::streamoff)sizeof(float) * ns;
    SCAI_ASSERT_ERROR(size % rec == 0, name << " is not a sequence of SU traces with " << ns << " samples")
    return (IndexType)(size / rec);
}

#include "Filter.hpp"
#include <algorithm>
#include <cmath>

using namespace KITGPI;

template <typename ValueType> IndexType Common::calcNextPowTwo(IndexType nt)
{
    ValueType temp = std::log(ValueType(nt));
    temp /= std::log(ValueType(2.0));
    temp = std::ceil(temp);
    temp = std::pow(ValueType(2.0), temp);
    return (IndexType)temp;
}

void Common::fft(std::vector<std::complex<double>> &a, bool inverse)
{
    const size_t n = a.size();
    SCAI_ASSERT_ERROR(n > 0 && (n & (n - 1)) == 0, "fft: the length must be a power of two")
    for (size_t i = 1, j = 0; i < n; i++) { // bit reversal
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1)
            j ^= bit;
        j ^= bit;
        if (i < j)
            std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2.0 * M_PI / (double)len * (inverse ? 1.0 : -1.0);
        const std::complex<double> wl(std::cos(ang), std::sin(ang));
        for (size_t i = 0; i < n; i += len) {
            std::complex<double> w(1.0, 0.0);
            for (size_t k = 0; k < len / 2; k++) {
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
                w *= wl;
            }
        }
    }
}

template <typename ValueType> void Filter::Filter<ValueType>::init(ValueType dt, IndexType nt)
{
    SCAI_ASSERT_ERROR(dt != 0.0, "Can't initialize filter with dt = 0.0")
    SCAI_ASSERT_ERROR(nt != 0, "Can't initialize filter with nt = 0")
    zeroPadding = Common::calcNextPowTwo<ValueType>(nt - 1) - nt;
    const IndexType filterLength = zeroPadding + nt;
    filter.assign(filterLength, ComplexValueType(1.0, 0.0));
    L.clear();
    Linv.clear();
    df = 1.0 / (filterLength * (double)dt);
    fNyquist = 1.0 / (2.0 * (double)dt);
    NT = nt;
}

template <typename ValueType> void Filter::Filter<ValueType>::frequencyVector(std::vector<double> &f) const
{
    const long nFreq = (long)(fNyquist / df); // Filter.cpp:25-31: 0 .. nFreq df, then -(nFreq - 1) df .. -df
    f.clear();
    for (long k = 0; k <= nFreq; k++)
        f.push_back(k * df);
    for (long k = 0; k < nFreq - 1; k++)
        f.push_back(-(nFreq - 1) * df + k * df);
}

// coefficients of the Butterworth polynomial, poly[i] = coefficient of s^i (Filter.cpp:66-111: the product of the factors
// s^2 - 2 cos((2k + n - 1) pi / 2n) s + 1, k = 1..n/2, and s + 1 for odd n; multiplied out directly here instead of by FFT)
template <typename ValueType> std::vector<double> Filter::Filter<ValueType>::butterPoly(IndexType order)
{
    std::vector<double> poly(1, 1.0);
    auto mul = [&](std::vector<double> const &f) {
        std::vector<double> r(poly.size() + f.size() - 1, 0.0);
        for (size_t i = 0; i < poly.size(); i++)
            for (size_t j = 0; j < f.size(); j++)
                r[i + j] += poly[i] * f[j];
        poly = r;
    };
    for (IndexType k = 1; k <= order / 2; k++)
        mul({1.0, -2.0 * std::cos((2.0 * k + order - 1.0) / (2.0 * order) * M_PI), 1.0});
    if (order % 2 != 0)
        mul({1.0, 1.0});
    return poly;
}

template <typename ValueType> void Filter::Filter<ValueType>::butterworth(std::vector<ComplexValueType> &h, bool highPass, IndexType order, ValueType fc) const
{
    std::vector<double> f;
    frequencyVector(f);
    const std::vector<double> poly = butterPoly(order);
    h.assign(f.size(), ComplexValueType(0.0, 0.0));
    for (size_t k = 0; k < f.size(); k++) {
        const ComplexValueType s(0.0, highPass ? -(double)fc / f[k] : f[k] / (double)fc); // Filter.cpp:41-63
        ComplexValueType acc(0.0, 0.0), p(1.0, 0.0);
        for (IndexType i = 0; i <= order; i++) {
            acc += poly[i] * p;
            p *= s;
        }
        h[k] = 1.0 / acc;
    }
    h[0] = highPass ? ComplexValueType(0.0, 0.0) : ComplexValueType(1.0, 0.0); // :180, :200
}

template <typename ValueType> void Filter::Filter<ValueType>::calc(std::string family, std::string type, IndexType order, ValueType fc1, ValueType fc2)
{
    std::transform(family.begin(), family.end(), family.begin(), ::tolower);
    std::transform(type.begin(), type.end(), type.begin(), ::tolower);
    SCAI_ASSERT_ERROR(!filter.empty(), "Filter::init must be called first")
    if (family == "butterworth") {
        SCAI_ASSERT_ERROR(fc1 > 0, "Lower corner frequency of filter has to be greater than zero.")
        if (type == "lp")
            butterworth(filter, false, order, fc1);
        else if (type == "hp")
            butterworth(filter, true, order, fc1);
        else if (type == "bp") {
            SCAI_ASSERT_ERROR(fc2 != 0.0, "Upper corner frequency of band-pass filter can't be zero")
            std::vector<ComplexValueType> hp;
            butterworth(filter, false, order, fc2);
            butterworth(hp, true, order, fc1);
            for (size_t k = 0; k < filter.size(); k++)
                filter[k] *= hp[k];
        } else
            COMMON_THROWEXCEPTION("Invalid filter type.")
    } else if (family == "ideal") {
        SCAI_ASSERT_ERROR(fc1 > 0, "Lower corner frequency of filter has to be greater than zero.")
        if (type != "bp")
            COMMON_THROWEXCEPTION("Invalid filter type.")
        const IndexType fc1Ind = (IndexType)std::ceil((double)fc1 / df);
        if (order == 0) { // the two FFT bins of that frequency pass (Filter.cpp:248-256; the function was 1 everywhere before)
            const IndexType len = (IndexType)(2 * fNyquist / df);
            filter[fc1Ind] = ComplexValueType(0.0, 0.0);
            filter[len - fc1Ind] = ComplexValueType(0.0, 0.0);
            for (auto &v : filter)
                v = ComplexValueType(1.0, 0.0) - v;
        } else { // one-frequency DFT over the NT samples (:257-275)
            const double dt = 1.0 / (2.0 * fNyquist);
            L.resize(NT);
            Linv.resize(NT);
            for (IndexType k = 0; k < NT; k++) {
                const double ph = 2.0 * M_PI * fc1Ind * df * (k * dt);
                L[k] = std::exp(ComplexValueType(0.0, -ph));
                Linv[k] = std::exp(ComplexValueType(0.0, ph));
            }
        }
    } else
        COMMON_THROWEXCEPTION("Invalid transfer function family.")
}

template <typename ValueType> void Filter::Filter<ValueType>::apply(std::vector<ValueType> &signal) const
{
    const IndexType len = (IndexType)(2 * fNyquist / df);
    SCAI_ASSERT_ERROR((IndexType)signal.size() + zeroPadding == len, "\nFilter is designed for different input length\n\n")
    if (L.empty()) {
        std::vector<ComplexValueType> f(len, ComplexValueType(0.0, 0.0));
        for (size_t k = 0; k < signal.size(); k++)
            f[k] = (double)signal[k];
        Common::fft(f, false);
        for (IndexType k = 0; k < len; k++)
            f[k] *= filter[k] / (double)len;
        Common::fft(f, true);
        for (size_t k = 0; k < signal.size(); k++)
            signal[k] = (ValueType)f[k].real();
    } else { // Filter.cpp:297-305
        ComplexValueType t(0.0, 0.0);
        for (IndexType k = 0; k < NT; k++)
            t += L[k] * (double)signal[k];
        for (IndexType k = 0; k < NT; k++)
            signal[k] = (ValueType)((Linv[k] * t) * (1.0 / (double)NT)).real();
    }
}

template <typename ValueType> void Filter::Filter<ValueType>::apply(std::vector<ValueType> &signal, IndexType numRows, IndexType nt) const
{
    SCAI_ASSERT_ERROR(signal.size() == (size_t)numRows * nt, "matrix size")
    if (L.empty()) {
        std::vector<ValueType> row(nt);
        for (IndexType r = 0; r < numRows; r++) {
            std::copy(signal.begin() + (size_t)r * nt, signal.begin() + (size_t)(r + 1) * nt, row.begin());
            apply(row);
            std::copy(row.begin(), row.end(), signal.begin() + (size_t)r * nt);
        }
        return;
    }
    // Filter.cpp:331-341: one-frequency DFT of every row, normalised by 2 / len, then the whole matrix scaled to maximum amplitude 1
    const IndexType len = (IndexType)(2 * fNyquist / df);
    SCAI_ASSERT_ERROR(nt + zeroPadding == len, "\nFilter is designed for different input length\n\n")
    ValueType maxNorm = 0;
    for (IndexType r = 0; r < numRows; r++) {
        ComplexValueType t(0.0, 0.0);
        for (IndexType k = 0; k < NT; k++)
            t += L[k] * (double)signal[(size_t)r * nt + k];
        t *= 2.0 / (double)len;
        for (IndexType k = 0; k < NT; k++) {
            const ValueType v = (ValueType)(t * Linv[k]).real();
            signal[(size_t)r * nt + k] = v;
            maxNorm = std::max(maxNorm, std::abs(v));
        }
    }
    for (auto &v : signal)
        v *= ValueType(1) / maxNorm;
}

template <typename ValueType> void Hilbert::HilbertFFT<ValueType>::calcHilbertCoefficient()
{
    kernel.assign(kernelSize, 0.0);
    if (2 * (kernelSize / 2) == kernelSize) {
        kernel[0] = 1.0;
        kernel[kernelSize / 2] = 1.0;
        for (IndexType i = 1; i < kernelSize / 2; i++)
            kernel[i] = 2.0;
    } else {
        kernel[0] = 1.0;
        for (IndexType i = 1; i < (kernelSize + 1) / 2; i++)
            kernel[i] = 2.0;
    }
}

template <typename ValueType> void Hilbert::HilbertFFT<ValueType>::hilbert(std::vector<ValueType> &data) const
{
    SCAI_ASSERT_ERROR((IndexType)kernel.size() == kernelSize && kernelSize >= (IndexType)data.size(), "HilbertFFT: calcHilbertCoefficient with a length >= the trace length first")
    std::vector<std::complex<double>> f(kernelSize, std::complex<double>(0.0, 0.0));
    for (size_t k = 0; k < data.size(); k++)
        f[k] = (double)data[k];
    Common::fft(f, false);
    for (IndexType k = 0; k < kernelSize; k++)
        f[k] *= kernel[k] / (double)kernelSize;
    Common::fft(f, true);
    for (size_t k = 0; k < data.size(); k++)
        data[k] = (ValueType)f[k].imag();
}

template <typename ValueType> void Hilbert::HilbertFFT<ValueType>::hilbert(std::vector<ValueType> &data, IndexType numRows, IndexType nt) const
{
    SCAI_ASSERT_ERROR(data.size() == (size_t)numRows * nt, "matrix size")
    std::vector<ValueType> row(nt);
    for (IndexType r = 0; r < numRows; r++) {
        std::copy(data.begin() + (size_t)r * nt, data.begin() + (size_t)(r + 1) * nt, row.begin());
        hilbert(row);
        std::copy(row.begin(), row.end(), data.begin() + (size_t)r * nt);
    }
}

template <typename ValueType> void Common::calcEnvelope(std::vector<ValueType> &data, IndexType numRows, IndexType nt)
{
    SCAI_ASSERT_ERROR(data.size() == (size_t)numRows * nt, "matrix size")
    if (data.empty())
        return;
    Hilbert::HilbertFFT<ValueType> hilbertHandler;
    IndexType kernelSize = calcNextPowTwo<ValueType>(nt - 1);
    while (kernelSize < nt) // nt = 2^k + 1
        kernelSize *= 2;
    hilbertHandler.setCoefficientLength(kernelSize);
    hilbertHandler.calcHilbertCoefficient();
    std::vector<ValueType> imag(data);
    hilbertHandler.hilbert(imag, numRows, nt);
    for (size_t k = 0; k < data.size(); k++)
        data[k] = std::sqrt(data[k] * data[k] + imag[k] * imag[k]);
}

template <typename ValueType> void Common::calcInstantaneousPhase(std::vector<ValueType> &data, IndexType numRows, IndexType nt, IndexType phaseType)
{
    SCAI_ASSERT_ERROR(data.size() == (size_t)numRows * nt, "matrix size")
    for (IndexType r = 0; r < numRows; r++) {
        ValueType *re = &data[(size_t)r * nt];
        std::vector<ValueType> im(re, re + nt);
        for (auto &v : im)
            v = -v;
        if (phaseType == 1) {
            for (IndexType t = 0; t < nt; t++)
                re[t] = std::atan(im[t] / re[t]);
        } else if (phaseType == 2) {
            for (IndexType t = 0; t < nt; t++)
                re[t] = std::atan2(im[t], re[t]);
        } else if (phaseType == 3 && nt > 0) {
            ValueType phase = std::atan2(im[0], re[0]);
            re[0] = phase;
            im[0] = phase;
            for (IndexType t = 1; t < nt; t++) {
                phase = std::atan2(im[t], re[t]);
                im[t] = phase;
                ValueType d = phase - im[t - 1];
                d = d > M_PI ? d - 2 * M_PI : (d < -M_PI ? d + 2 * M_PI : d);
                re[t] = re[t - 1] + d;
            }
        }
    }
}

template void Common::calcEnvelope<float>(std::vector<float> &, IndexType, IndexType);
template void Common::calcEnvelope<double>(std::vector<double> &, IndexType, IndexType);
template void Common::calcInstantaneousPhase<float>(std::vector<float> &, IndexType, IndexType, IndexType);
template void Common::calcInstantaneousPhase<double>(std::vector<double> &, IndexType, IndexType, IndexType);
template IndexType Common::calcNextPowTwo<float>(IndexType);
template IndexType Common::calcNextPowTwo<double>(IndexType);
template class Filter::Filter<float>;
template class Filter::Filter<double>;
template class Hilbert::HilbertFFT<float>;
template class Hilbert::HilbertFFT<double>;

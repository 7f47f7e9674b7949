// Filter.hpp — frequency filters and the Hilbert transform of traces (mirror of src/Filter/Filter.hpp:20-61 and
// src/Common/Hilbert.hpp, HilbertFFT.cpp): the tools WAVE-Inversion applies to seismograms and source signals through the classes
// of libSimulation (SeismogramHandler::filter, Seismogram::filterTraces; Simulation.cpp:404-408 for the Hilbert-transformed source).
// Signals are zero-padded to the next power of two of (nt - 1) (Common::calcNextPowTwo), transformed with a radix-2 FFT, multiplied
// by the transfer function and transformed back.  Host code: traces are nTraces x NT matrices of a few MB.
#pragma once
#include "Common.hpp"
#include <complex>

namespace KITGPI
{
    namespace Common
    {
        //! 2^ceil(log2(nt)) evaluated in ValueType like Common.hpp:187-194
        template <typename ValueType> IndexType calcNextPowTwo(IndexType nt);
        //! in-place radix-2 transform of a power-of-two length; inverse = conjugate kernel WITHOUT the 1/n factor (as lama::fft / ifft)
        void fft(std::vector<std::complex<double>> &a, bool inverse);
        //! envelope sqrt(x^2 + H(x)^2) of the rows of a row-major numRows x nt matrix (Common.hpp:272-290)
        template <typename ValueType> void calcEnvelope(std::vector<ValueType> &data, IndexType numRows, IndexType nt);
        //! "instantaneous phase" of the rows as Common.hpp:297-340 computes it: the imaginary part is -x (the Hilbert transform is
        //! commented out there); phaseType 1 atan(im / re), 2 atan2(im, re), 3 unwrapped
        template <typename ValueType> void calcInstantaneousPhase(std::vector<ValueType> &data, IndexType numRows, IndexType nt, IndexType phaseType);
    }

    namespace Filter
    {
        template <typename ValueType> class Filter
        {
          public:
            typedef std::complex<double> ComplexValueType;
            //! transfer function = 1: nothing is filtered yet (Filter.cpp:8-19)
            void init(ValueType dt, IndexType nt);
            //! transFcnFmly "butterworth" with filterType "lp" | "hp" | "bp", or "ideal" with "bp" (one frequency bin; order 0: FFT bin,
            //! else a one-frequency DFT); Filter.cpp:121-263
            void calc(std::string transFcnFmly, std::string filterType, IndexType order, ValueType fc1, ValueType fc2 = 0.0);
            void apply(std::vector<ValueType> &signal) const;                                 // one trace of nt samples
            void apply(std::vector<ValueType> &signal, IndexType numRows, IndexType nt) const; // row-major traces
            std::vector<ComplexValueType> const &getTransferFunction() const { return filter; }

          private:
            void frequencyVector(std::vector<double> &f) const;
            static std::vector<double> butterPoly(IndexType order);
            void butterworth(std::vector<ComplexValueType> &h, bool highPass, IndexType order, ValueType fc) const;
            IndexType zeroPadding = 0, NT = 0;
            std::vector<ComplexValueType> filter;
            std::vector<ComplexValueType> L, Linv; // one-frequency DFT (ideal band-pass with order != 0)
            double df = 0, fNyquist = 0;
        };
    }

    namespace Hilbert
    {
        //! Hilbert transform of traces through the analytic signal (HilbertFFT.cpp:10-74)
        template <typename ValueType> class HilbertFFT
        {
          public:
            void setCoefficientLength(IndexType size) { kernelSize = size; }
            void calcHilbertCoefficient();
            void hilbert(std::vector<ValueType> &data) const;                                  // one trace
            void hilbert(std::vector<ValueType> &data, IndexType numRows, IndexType nt) const; // row-major traces

          private:
            IndexType kernelSize = 0;
            std::vector<double> kernel;
        };
    }
}

// ws_kernels_general.cu — instantiation + launch dispatch of the general kernels (see ws_kernels_general.cuh)
#include "../../include/wavesim.h"
#include "ws_kernels_general.cuh"
#include "ws_launch.hpp"

namespace {

template <int EQ, int DIM, bool EXACT>
void launchT(const WsParams &P, int pass, dim3 grid, dim3 block, cudaStream_t st)
{
    auto k0 = wsgen::kGeneral<EQ, DIM, EXACT, 0>;
    auto k1 = wsgen::kGeneral<EQ, DIM, EXACT, 1>;
    if (pass == 0)
        WS_LAUNCH(k0, grid, block, 0, st, P);
    else
        WS_LAUNCH(k1, grid, block, 0, st, P);
}

template <int EQ, int DIM>
void launchE(const WsParams &P, bool exact, int pass, dim3 grid, dim3 block, cudaStream_t st)
{
    if (exact)
        launchT<EQ, DIM, true>(P, pass, grid, block, st);
    else
        launchT<EQ, DIM, false>(P, pass, grid, block, st);
}

template <int EQ>
void launchD(const WsParams &P, bool exact, int pass, dim3 grid, dim3 block, cudaStream_t st)
{
    if (P.dim == 3)
        launchE<EQ, 3>(P, exact, pass, grid, block, st);
    else
        launchE<EQ, 2>(P, exact, pass, grid, block, st);
}
template <int EQ>
void launch2(const WsParams &P, bool exact, int pass, dim3 grid, dim3 block, cudaStream_t st)
{
    launchE<EQ, 2>(P, exact, pass, grid, block, st);
}

} // namespace

void wsGeneralGrid(const WsParams &P, dim3 &grid, dim3 &block)
{
    wsPointGrid(P.nx, P.nz, P.yhi - P.ylo, grid, block);
}

void wsLaunchGeneral(const WsParams &P, bool exact, int pass, cudaStream_t st)
{
    if (P.yhi <= P.ylo)
        return;
    dim3 grid, block;
    wsGeneralGrid(P, grid, block);
    switch (P.eq) {
    case WS_EQ_ACOUSTIC: launchD<WS_EQ_ACOUSTIC>(P, exact, pass, grid, block, st); break;
    case WS_EQ_ELASTIC: launchD<WS_EQ_ELASTIC>(P, exact, pass, grid, block, st); break;
    case WS_EQ_VISCOELASTIC: launchD<WS_EQ_VISCOELASTIC>(P, exact, pass, grid, block, st); break;
    case WS_EQ_SH: launch2<WS_EQ_SH>(P, exact, pass, grid, block, st); break;
    case WS_EQ_VISCOSH: launch2<WS_EQ_VISCOSH>(P, exact, pass, grid, block, st); break;
    case WS_EQ_TMEM: launch2<WS_EQ_TMEM>(P, exact, pass, grid, block, st); break;
    case WS_EQ_VISCOTMEM: launch2<WS_EQ_VISCOTMEM>(P, exact, pass, grid, block, st); break;
    case WS_EQ_EMEM: launchD<WS_EQ_EMEM>(P, exact, pass, grid, block, st); break;
    case WS_EQ_VISCOEMEM: launchD<WS_EQ_VISCOEMEM>(P, exact, pass, grid, block, st); break;
    default: break;
    }
}

void wsLaunchAbsFirstHalf(const WsParams &P, bool exact, int f0, int f1, int f2, cudaStream_t st)
{
    if (P.yhi <= P.ylo)
        return;
    dim3 grid, block;
    wsGeneralGrid(P, grid, block);
    auto kt = wsgen::kAbsFirstHalf<true>;
    auto kf = wsgen::kAbsFirstHalf<false>;
    if (exact)
        WS_LAUNCH(kt, grid, block, 0, st, P, f0, f1, f2);
    else
        WS_LAUNCH(kf, grid, block, 0, st, P, f0, f1, f2);
}

void wsLaunchDivCurl(const WsParams &P, float *out, int which, cudaStream_t st)
{
    if (P.yhi <= P.ylo)
        return;
    dim3 grid, block;
    wsGeneralGrid(P, grid, block);
    const bool em = P.eq == WS_EQ_TMEM || P.eq == WS_EQ_VISCOTMEM || P.eq == WS_EQ_EMEM || P.eq == WS_EQ_VISCOEMEM;
    const int divCoef = P.eq == WS_EQ_VISCOEMEM ? 1 : 0; // Wavefields3Dviscoemem.cpp:99 passes the conductivity
    auto s2 = wsgen::kDivCurl<2, false>;
    auto s3 = wsgen::kDivCurl<3, false>;
    auto e2 = wsgen::kDivCurl<2, true>;
    auto e3 = wsgen::kDivCurl<3, true>;
    if (em) {
        if (P.dim == 3)
            WS_LAUNCH(e3, grid, block, 0, st, P, out, which, divCoef);
        else
            WS_LAUNCH(e2, grid, block, 0, st, P, out, which, divCoef);
    } else {
        if (P.dim == 3)
            WS_LAUNCH(s3, grid, block, 0, st, P, out, which, divCoef);
        else
            WS_LAUNCH(s2, grid, block, 0, st, P, out, which, divCoef);
    }
}

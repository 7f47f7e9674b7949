// ws_kernels_tma.cu — instantiation, TMA program / tensor-map construction and launch dispatch of the warp-specialised TMA
// marching kernels (ws_kernels_tma.cuh).  Built once per FD order (-DWS_TMA_Q=<q>: the kernels of that order) and once
// without (program builder + dispatcher), so that the orders compile in parallel.
#include "../../include/wavesim.h"
#include "ws_kernels_tma.cuh"
#include "ws_launch.hpp"

#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int kTmaMaxSmem = 227 * 1024 - 1024;

template <int EQ, int DIM, int Q, int PASS, int NL> void launchK(const WsParams &P, const wstma::TmaProg &prog, cudaStream_t st)
{
    using G = wstma::Geo<DIM, Q, NL>;
    auto k = wstma::kTma<EQ, DIM, Q, PASS, NL>;
    if (wsOptInSmem(reinterpret_cast<const void *>(k), kTmaMaxSmem) != cudaSuccess)
        return; // the error stays pending: ws_step / ws_run report it through cudaGetLastError
    const int ny = P.yhi - P.ylo;
    const dim3 grid((P.nx + G::NS * G::TX - 1) / (G::NS * G::TX), (P.nz + G::TZ - 1) / G::TZ, (ny + P.tmaChunk - 1) / P.tmaChunk);
    const size_t smem = (size_t)prog.nst * prog.stageFloats * sizeof(float);
    k<<<grid, G::NTHR, smem, st>>>(P, prog);
}

template <int EQ, int DIM, int Q> void launchT(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st)
{
    if constexpr (DIM == 2) { // a 2-D strip stages few bytes per thread: fewer points per thread give more warps per staged byte
        if (nl == 2) {
            if (pass == 0)
                launchK<EQ, DIM, Q, 0, 2>(P, prog, st);
            else
                launchK<EQ, DIM, Q, 1, 2>(P, prog, st);
            return;
        }
        if (nl == 1) {
            if (pass == 0)
                launchK<EQ, DIM, Q, 0, 1>(P, prog, st);
            else
                launchK<EQ, DIM, Q, 1, 1>(P, prog, st);
            return;
        }
    }
    constexpr bool heavy = DIM == 3 && (EQ == WS_EQ_ELASTIC || EQ == WS_EQ_VISCOELASTIC || EQ == WS_EQ_VISCOEMEM);
    if constexpr (heavy) {
        if (nl == 1) {
            if (pass == 0)
                launchK<EQ, DIM, Q, 0, 1>(P, prog, st);
            else
                launchK<EQ, DIM, Q, 1, 1>(P, prog, st);
            return;
        }
    }
    if (pass == 0)
        launchK<EQ, DIM, Q, 0, 4>(P, prog, st);
    else
        launchK<EQ, DIM, Q, 1, 4>(P, prog, st);
}

template <int EQ, int Q> void launchD(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st)
{
    if (P.dim == 3)
        launchT<EQ, 3, Q>(P, pass, prog, nl, st);
    else
        launchT<EQ, 2, Q>(P, pass, prog, nl, st);
}

template <int Q> void launchQ(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st)
{
    switch (P.eq) {
    case WS_EQ_ACOUSTIC: launchD<WS_EQ_ACOUSTIC, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_ELASTIC: launchD<WS_EQ_ELASTIC, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_VISCOELASTIC: launchD<WS_EQ_VISCOELASTIC, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_SH: launchT<WS_EQ_SH, 2, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_VISCOSH: launchT<WS_EQ_VISCOSH, 2, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_TMEM: launchT<WS_EQ_TMEM, 2, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_VISCOTMEM: launchT<WS_EQ_VISCOTMEM, 2, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_EMEM: launchD<WS_EQ_EMEM, Q>(P, pass, prog, nl, st); break;
    case WS_EQ_VISCOEMEM: launchD<WS_EQ_VISCOEMEM, Q>(P, pass, prog, nl, st); break;
    default: break;
    }
}

} // namespace

#define WS_TMA_NAME2(q) wsLaunchTmaQ##q
#define WS_TMA_NAME(q) WS_TMA_NAME2(q)

#ifdef WS_TMA_Q
void WS_TMA_NAME(WS_TMA_Q)(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st) { launchQ<WS_TMA_Q>(P, pass, prog, nl, st); }
#else
void wsLaunchTmaQ2(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);
void wsLaunchTmaQ4(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);
void wsLaunchTmaQ6(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);
void wsLaunchTmaQ8(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);
void wsLaunchTmaQ10(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);
void wsLaunchTmaQ12(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st);

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encodeFn()
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p)
        throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return reinterpret_cast<EncodeTiledFn>(p);
}

// 4-D map over an arena of padded arrays: (x, z, y, array); box = boxX x boxZ x boxY planes x boxA arrays
CUtensorMap makeMap(const float *arena, int pitch, int nzp, int nyp, long long arrayStride, int nArrays, int boxX, int boxZ, int boxY, int boxA)
{
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)nzp, (cuuint64_t)nyp, (cuuint64_t)nArrays};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * nzp * 4, (cuuint64_t)arrayStride * 4};
    const cuuint32_t box[4] = {(cuuint32_t)boxX, (cuuint32_t)boxZ, (cuuint32_t)boxY, (cuuint32_t)boxA};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encodeFn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(arena), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

struct Entry {
    bool halo;
    int arena, pos, dy;
};

// x points per thread of a half-step: four (128-bit accesses, index arithmetic shared by the points) except where the operands
// of a plane are so many that a four-point tile would not fit the stage ring twice: the 3-D viscoelastic stress half-step and
// the 3-D viscoEM E half-step (196+ bytes per point).  Measured on the north-star grid with order-reducing edges
// (3-D elastic 1024^3, profiles/r02_tma_sweep.txt): 31.5 Gpt/s with four points per thread against 22.2 with one.
int lanesFor(const WsParams &P, int pass)
{
    if (P.dim != 3)
        return (P.marchLanes == 1 || P.marchLanes == 2 || P.marchLanes == 4) ? P.marchLanes : 4;
    const bool heavyKernels = P.eq == WS_EQ_ELASTIC || P.eq == WS_EQ_VISCOELASTIC || P.eq == WS_EQ_VISCOEMEM; // instantiated with 1 and 4
    if (P.marchLanes == 4 || (P.marchLanes == 1 && heavyKernels))
        return P.marchLanes;
    if (pass == 1 && P.L > 0 && (P.eq == WS_EQ_VISCOELASTIC || P.eq == WS_EQ_VISCOEMEM))
        return 1;
    return 4;
}

} // namespace

bool wsTmaSupported(const WsParams &P, const WsArenaInfo &A, bool exact)
{
    if (exact || !A.base[0])
        return false;
    if (P.q < 2 || P.q > 12 || (P.q & 1))
        return false;
    if (getenv("WS_NO_TMA_MARCH") && atoi(getenv("WS_NO_TMA_MARCH")) != 0)
        return false;
    return P.dim == 2 || P.dim == 3;
}

// Builds the tensor maps (device buffer, returned; freed with wsTmaRelease) and the per-half-step TMA programs.
void *wsTmaPrepare(WsParams &P, const WsArenaInfo &A, int nyp, wstma::TmaProg prog[2], int nl[2])
{
    using namespace wstma;
    std::vector<CUtensorMap> maps;
    P.marchLanes = getenv("WS_MARCH_LANES") ? atoi(getenv("WS_MARCH_LANES")) : 0;
    P.marchDebug = getenv("WS_MARCH_DEBUG") ? atoi(getenv("WS_MARCH_DEBUG")) : 0;
    P.marchStageR = 1; // the memory variables always travel through the stage ring
    for (int pass = 0; pass < 2; pass++) {
        nl[pass] = lanesFor(P, pass);
        const GeoRT g = geoOf(P.dim, P.q, nl[pass]);
        const Lists S = spec(P.eq, P.dim, pass);
        // entries of a stage in the order the consumers index them (wsmarch::MPt): halo tiles, then plain tiles
        std::vector<Entry> ent;
        auto fld = [&](int slot, bool halo, int dy) {
            if (A.fldPos[slot] < 0)
                throw std::runtime_error("TMA program: wavefield slot " + std::to_string(slot) + " is not in an arena");
            ent.push_back({halo, A.fldArena[slot], A.fldPos[slot], dy});
        };
        auto mat = [&](int slot) {
            if (A.matPos[slot] < 0)
                throw std::runtime_error("TMA program: model slot " + std::to_string(slot) + " is not in an arena");
            ent.push_back({false, A.matArena[slot], A.matPos[slot], 0});
        };
        for (int k = 0; k < S.nt; k++)
            fld(S.t[k], true, 0);
        for (int k = 0; k < S.nq; k++)
            fld(S.qf[k], false, g.H);
        for (int k = 0; k < S.nf; k++)
            fld(S.f[k], false, 0);
        for (int k = 0; k < S.nm; k++)
            mat(S.m[k]);
        for (int l = 0; l < P.L; l++)
            for (int k = 0; k < S.nr; k++)
                fld(F_R0 + 6 * l + S.r[k], false, 0);
        for (int l = 0; l < P.L; l++)
            for (int k = 0; k < S.nc; k++)
                mat(M_CD0 + 3 * l + S.c[k]);
        const int nPlain = (int)ent.size() - S.nt;
        const int stripFloats = S.nt * g.TS + g.PB * nPlain * g.NP;
        TmaProg &pr = prog[pass];
        std::memset(&pr, 0, sizeof(pr));
        pr.stageFloats = g.NS * stripFloats;
        // map of (arena, halo, number of arrays): created on demand
        struct Key { int arena, halo, cnt, idx; };
        std::vector<Key> keys;
        auto mapOf = [&](int arena, bool halo, int cnt) {
            for (auto &k : keys)
                if (k.arena == arena && k.halo == (int)halo && k.cnt == cnt)
                    return k.idx;
            maps.push_back(makeMap(A.base[arena], P.pitch, P.nzp, nyp, A.stride, A.count[arena], halo ? g.LDX : g.TX, halo ? g.NROW : g.TZ, g.PB, cnt));
            keys.push_back({arena, (int)halo, cnt, (int)maps.size() - 1});
            return keys.back().idx;
        };
        unsigned bytes = 0;
        for (int s = 0; s < g.NS; s++) {
            size_t e = 0;
            while (e < ent.size()) {
                const Entry &a = ent[e];
                int cnt = 1;
                const bool mergeable = g.PB == 1 && (!a.halo || g.haloMerge);
                while (mergeable && cnt < CMAX && e + cnt < ent.size()) {
                    const Entry &b = ent[e + cnt];
                    if (b.halo != a.halo || b.arena != a.arena || b.dy != a.dy || b.pos != a.pos + cnt)
                        break;
                    cnt++;
                }
                if (pr.nOps >= MAXOPS)
                    throw std::runtime_error("TMA program exceeds MAXOPS");
                const int off = s * stripFloats + (a.halo ? (int)e * g.TS : S.nt * g.TS + ((int)e - S.nt) * g.PB * g.NP);
                TmaOp &o = pr.op[pr.nOps++];
                o.dst16 = (unsigned short)(off / 4);
                o.map = (unsigned char)mapOf(a.arena, a.halo, cnt);
                o.slot = (unsigned char)a.pos;
                o.dx = (short)(s * g.TX - (a.halo ? g.HX : 0));
                o.dz = (short)(a.halo ? -g.HZ : 0);
                o.dy = (short)a.dy;
                bytes += 4u * cnt * g.PB * (a.halo ? g.BOX : g.NP);
                e += cnt;
            }
        }
        pr.stageBytes = bytes;
        // depth of the ring (measured on BASELINE configs 3-5, profiles/r02_tma_sweep.txt).  The half-steps with few operands
        // per point are bound by the bytes in flight: 4 stages, even when that leaves one thread block per SM (3-D acoustic
        // 1024^3: 86 against 82 Gpt/s with 3 stages and two blocks).  The 3-D elastic / viscoelastic half-steps are bound by
        // the resident warps (~700 instructions per point): 2 stages and as many blocks as fit (3-D viscoelastic 768^3:
        // 16.4 against 11.5 Gpt/s with 4 stages).  2-D strips: 3 stages.
        const size_t stageB = (size_t)pr.stageFloats * 4;
        const bool heavy = P.dim == 3 && nl[pass] == 1;
        int nst = heavy ? 2 : (P.dim == 3 ? 4 : 3);
        while (nst > 2 && nst * stageB > (size_t)kTmaMaxSmem)
            nst--;
        if (nst * stageB > (size_t)kTmaMaxSmem)
            throw std::runtime_error("TMA program: a stage does not fit into shared memory");
        if (const char *e = getenv("WS_TMA_STAGES"))
            if (atoi(e) >= 2 && atoi(e) <= NSTMAX && atoi(e) * stageB <= (size_t)kTmaMaxSmem)
                nst = atoi(e);
        pr.nst = nst;
    }
    // planes per thread block: short chunks keep the thread blocks of one y range in step (shared halo rows stay in L2)
    const GeoRT g0 = geoOf(P.dim, P.q, 4);
    int chunk = 64;
    if (const char *e = getenv("WS_TMA_CHUNK"))
        chunk = atoi(e) > 0 ? atoi(e) : chunk;
    chunk = (chunk + g0.PB - 1) / g0.PB * g0.PB;
    P.tmaChunk = chunk;
    void *dev = nullptr;
    if (cudaMalloc(&dev, sizeof(CUtensorMap) * maps.size()) != cudaSuccess)
        throw std::runtime_error("cudaMalloc for tensor maps failed");
    cudaMemcpy(dev, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice);
    P.tmaMaps = dev;
    return dev;
}

void wsTmaRelease(void *maps)
{
    if (maps)
        cudaFree(maps);
}

int wsLaunchTma(const WsParams &P, int pass, const wstma::TmaProg &prog, int nl, cudaStream_t st)
{
    if (P.yhi <= P.ylo || P.tmaChunk <= 0 || !P.tmaMaps)
        return 0;
    switch (P.q) {
    case 2: wsLaunchTmaQ2(P, pass, prog, nl, st); break;
    case 4: wsLaunchTmaQ4(P, pass, prog, nl, st); break;
    case 6: wsLaunchTmaQ6(P, pass, prog, nl, st); break;
    case 8: wsLaunchTmaQ8(P, pass, prog, nl, st); break;
    case 10: wsLaunchTmaQ10(P, pass, prog, nl, st); break;
    case 12: wsLaunchTmaQ12(P, pass, prog, nl, st); break;
    default: return 0;
    }
    return 1;
}
#endif

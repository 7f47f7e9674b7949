// ws_kernels_tile2d.cu — instantiation, TMA program / tensor-map construction and launch dispatch of the 2-D tile kernels
// (ws_kernels_tile2d.cuh).  Built once per FD order (-DWS_TILE_Q=<q>: the kernels of that order) and once without
// (program builder + dispatcher), so that the orders compile in parallel.
#include "../../include/wavesim.h"
#include "ws_kernels_tile2d.cuh"
#include "ws_launch.hpp"

#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int kTileMaxSmem = 227 * 1024 - 1024;

template <int EQ, int Q, int PASS, int TY> void launchK(const WsParams &P, const wstile::TileProg &prog, cudaStream_t st)
{
    using G = wstile::Geo<Q, TY>;
    auto k = wstile::kTile2D<EQ, Q, PASS, TY>;
    if (wsOptInSmem(reinterpret_cast<const void *>(k), kTileMaxSmem) != cudaSuccess)
        return; // the error stays pending: ws_step / ws_run report it through cudaGetLastError
    const int tilesX = (P.nx + G::TX - 1) / G::TX, tilesY = (P.yhi - P.ylo + TY - 1) / TY;
    // programmatic stream serialization: the thread blocks may be scheduled while the previous kernel of the stream drains (they wait
    // for its completion inside, before their first fetch): takes the launch latency out of the 0.1 ms half-steps
    static const bool pdl = !(getenv("WS_PDL") && atoi(getenv("WS_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)tilesX * (unsigned)tilesY);
    cfg.blockDim = dim3(G::NTHR);
    cfg.dynamicSmemBytes = (size_t)prog.totalFloats * sizeof(float);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k, P, prog, tilesX);
}

template <int EQ, int Q> void launchT(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st)
{
    if (P.tileTY == 16) {
        if (pass == 0)
            launchK<EQ, Q, 0, 16>(P, prog, st);
        else
            launchK<EQ, Q, 1, 16>(P, prog, st);
    } else {
        if (pass == 0)
            launchK<EQ, Q, 0, 8>(P, prog, st);
        else
            launchK<EQ, Q, 1, 8>(P, prog, st);
    }
}

template <int Q> void launchQ(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st)
{
    switch (P.eq) {
    case WS_EQ_ACOUSTIC: launchT<WS_EQ_ACOUSTIC, Q>(P, pass, prog, st); break;
    case WS_EQ_ELASTIC: launchT<WS_EQ_ELASTIC, Q>(P, pass, prog, st); break;
    case WS_EQ_VISCOELASTIC: launchT<WS_EQ_VISCOELASTIC, Q>(P, pass, prog, st); break;
    case WS_EQ_SH: launchT<WS_EQ_SH, Q>(P, pass, prog, st); break;
    case WS_EQ_VISCOSH: launchT<WS_EQ_VISCOSH, Q>(P, pass, prog, st); break;
    case WS_EQ_TMEM: launchT<WS_EQ_TMEM, Q>(P, pass, prog, st); break;
    case WS_EQ_VISCOTMEM: launchT<WS_EQ_VISCOTMEM, Q>(P, pass, prog, st); break;
    case WS_EQ_EMEM: launchT<WS_EQ_EMEM, Q>(P, pass, prog, st); break;
    case WS_EQ_VISCOEMEM: launchT<WS_EQ_VISCOEMEM, Q>(P, pass, prog, st); break;
    default: break;
    }
}

} // namespace

#define WS_TILE_NAME2(q) wsLaunchTileQ##q
#define WS_TILE_NAME(q) WS_TILE_NAME2(q)

#ifdef WS_TILE_Q
void WS_TILE_NAME(WS_TILE_Q)(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st) { launchQ<WS_TILE_Q>(P, pass, prog, st); }
#else
void wsLaunchTileQ2(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);
void wsLaunchTileQ4(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);
void wsLaunchTileQ6(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);
void wsLaunchTileQ8(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);
void wsLaunchTileQ10(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);
void wsLaunchTileQ12(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st);

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

// 3-D map over an arena of padded 2-D arrays: (x, y, array); box = boxX x boxY rows x boxA arrays
CUtensorMap makeMap(const float *arena, int pitch, int nyp, long long arrayStride, int nArrays, int boxX, int boxY, int boxA)
{
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)nyp, (cuuint64_t)nArrays};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)arrayStride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)boxX, (cuuint32_t)boxY, (cuuint32_t)boxA};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encodeFn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(arena), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

struct Entry {
    bool halo;
    int arena, pos;
};

// entries of a tile in the order the consumers index them (wsmarch::MPt, YSM form): halo tiles of the differentiated
// fields, then the plain tiles (updated fields, model parameters, memory variables, Cd coefficients)
std::vector<Entry> entriesOf(const WsParams &P, const WsArenaInfo &A, int pass, int &nHalo)
{
    using namespace wstile;
    const Lists S = spec(P.eq, 2, pass);
    const HaloSet U = haloSet(S);
    std::vector<Entry> ent;
    auto fld = [&](int slot, bool halo) {
        if (A.fldPos[slot] < 0)
            throw std::runtime_error("tile program: wavefield slot " + std::to_string(slot) + " is not in an arena");
        ent.push_back({halo, A.fldArena[slot], A.fldPos[slot]});
    };
    auto mat = [&](int slot) {
        if (A.matPos[slot] < 0)
            throw std::runtime_error("tile program: model slot " + std::to_string(slot) + " is not in an arena");
        ent.push_back({false, A.matArena[slot], A.matPos[slot]});
    };
    for (int k = 0; k < U.n; k++)
        fld(U.f[k], true);
    nHalo = U.n;
    for (int k = 0; k < S.nf; k++)
        fld(S.f[k], false);
    for (int k = 0; k < S.nm; k++)
        mat(S.m[k]);
    for (int l = 0; l < P.L; l++)
        for (int k = 0; k < S.nr; k++)
            fld(F_R0 + 6 * l + S.r[k], false);
    for (int l = 0; l < P.L; l++)
        for (int k = 0; k < S.nc; k++)
            mat(M_CD0 + 3 * l + S.c[k]);
    return ent;
}

size_t tileBytes(const WsParams &P, const WsArenaInfo &A, int pass, int ty)
{
    int nHalo = 0;
    const auto ent = entriesOf(P, A, pass, nHalo);
    const wstile::GeoRT g = wstile::geoOf(P.q, ty);
    return 4 * ((size_t)nHalo * g.TS + (ent.size() - nHalo) * (size_t)ty * g.NP);
}

} // namespace

bool wsTileSupported(const WsParams &P, const WsArenaInfo &A, bool exact)
{
    if (exact || !A.base[0] || P.dim != 2)
        return false;
    if (P.q < 2 || P.q > 12 || (P.q & 1))
        return false;
    if (getenv("WS_NO_TILE2D") && atoi(getenv("WS_NO_TILE2D")) != 0)
        return false;
    return tileBytes(P, A, 0, 8) <= (size_t)kTileMaxSmem && tileBytes(P, A, 1, 8) <= (size_t)kTileMaxSmem;
}

// Builds the tensor maps (device buffer, returned; freed with wsTileRelease) and the per-half-step TMA programs.
void *wsTilePrepare(WsParams &P, const WsArenaInfo &A, int nyp, wstile::TileProg prog[2])
{
    using namespace wstile;
    std::vector<CUtensorMap> maps;
    P.marchDebug = getenv("WS_MARCH_DEBUG") ? atoi(getenv("WS_MARCH_DEBUG")) : 0;
    P.marchStageR = 1; // the memory variables travel with the tile
    // rows per tile: 8 (measured on B200, profiles/r02_tile2d.txt: 2-D elastic 4096^2 65.1 Gpt/s against 62.3 with 16 rows, 2-D
    // viscoTMEz 8192 x 2048 83.7 against 82.2: more resident thread blocks per SM beat fewer halo rows per point)
    int ty = 8;
    if (const char *e = getenv("WS_TILE_TY"))
        if ((atoi(e) == 8 || atoi(e) == 16) && tileBytes(P, A, 0, atoi(e)) <= (size_t)kTileMaxSmem && tileBytes(P, A, 1, atoi(e)) <= (size_t)kTileMaxSmem)
            ty = atoi(e);
    P.tileTY = ty;
    const GeoRT g = geoOf(P.q, ty);
    const bool haloMerge = g.TS == g.LDX * g.NROW;
    for (int pass = 0; pass < 2; pass++) {
        int nHalo = 0;
        const std::vector<Entry> ent = entriesOf(P, A, pass, nHalo);
        TileProg &pr = prog[pass];
        std::memset(&pr, 0, sizeof(pr));
        pr.haloFloats = nHalo * g.TS;
        pr.totalFloats = pr.haloFloats + ((int)ent.size() - nHalo) * ty * g.NP;
        struct Key { int arena, halo, cnt, idx; };
        std::vector<Key> keys;
        auto mapOf = [&](int arena, bool halo, int cnt) {
            for (auto &k : keys)
                if (k.arena == arena && k.halo == (int)halo && k.cnt == cnt)
                    return k.idx;
            maps.push_back(makeMap(A.base[arena], P.pitch, nyp, A.stride, A.count[arena], halo ? g.LDX : g.TX, halo ? g.NROW : ty, cnt));
            keys.push_back({arena, (int)halo, cnt, (int)maps.size() - 1});
            return keys.back().idx;
        };
        unsigned bytes = 0;
        size_t e = 0;
        while (e < ent.size()) {
            const Entry &a = ent[e];
            int cnt = 1;
            const bool mergeable = !a.halo || haloMerge;
            while (mergeable && cnt < wstma::CMAX && e + cnt < ent.size()) {
                const Entry &b = ent[e + cnt];
                if (b.halo != a.halo || b.arena != a.arena || b.pos != a.pos + cnt)
                    break;
                cnt++;
            }
            if (pr.nOps >= MAXOPS)
                throw std::runtime_error("tile program exceeds MAXOPS");
            const size_t off = a.halo ? e * (size_t)g.TS : (size_t)pr.haloFloats + (e - nHalo) * (size_t)ty * g.NP;
            TileOp &o = pr.op[pr.nOps++];
            o.dst = (unsigned)(off * 4);
            o.map = (unsigned char)mapOf(a.arena, a.halo, cnt);
            o.slot = (unsigned char)a.pos;
            o.dx = (short)(a.halo ? -g.HX : 0);
            o.dy = (short)(a.halo ? -g.H : 0);
            bytes += 4u * cnt * (a.halo ? g.LDX * g.NROW : ty * g.NP);
            e += cnt;
        }
        pr.bytes = bytes;
    }
    void *dev = nullptr;
    if (cudaMalloc(&dev, sizeof(CUtensorMap) * maps.size()) != cudaSuccess)
        throw std::runtime_error("cudaMalloc for tensor maps failed");
    cudaMemcpy(dev, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice);
    P.tileMaps = dev;
    return dev;
}

void wsTileRelease(void *maps)
{
    if (maps)
        cudaFree(maps);
}

int wsLaunchTile(const WsParams &P, int pass, const wstile::TileProg &prog, cudaStream_t st)
{
    if (P.yhi <= P.ylo || !P.tileMaps)
        return 0;
    switch (P.q) {
    case 2: wsLaunchTileQ2(P, pass, prog, st); break;
    case 4: wsLaunchTileQ4(P, pass, prog, st); break;
    case 6: wsLaunchTileQ6(P, pass, prog, st); break;
    case 8: wsLaunchTileQ8(P, pass, prog, st); break;
    case 10: wsLaunchTileQ10(P, pass, prog, st); break;
    case 12: wsLaunchTileQ12(P, pass, prog, st); break;
    default: return 0;
    }
    return 1;
}
#endif

// ws_kernels_tile2d.cuh — 2-D tile kernels: every 2-D equation type, FD order, edge policy and boundary condition (FMA
// arithmetic).  The 2-D half-steps last 0.1-0.2 ms on a 4096^2 grid; a marching thread block there is a chain of a few
// dozen dependent row steps of 2-3 us each (profiles/r01_ncu_march_cfg2_4096.txt), so the 2-D grids get a kernel without
// any march:
//
//   * a thread block owns a TX x TY tile of the x-y grid (128 x 8 / 16 / 32 points), one thread per 4 consecutive x points
//     of one row (a warp = one row of the tile);
//   * ONE elected thread fetches every operand of the tile with a handful of TMA boxes (cp.async.bulk.tensor.3d over
//     (x, y, array) maps of the solver's arena): the differentiated fields as halo tiles (TX + 2 HX) x (TY + 2 H) — the x
//     AND the y stencils read them, there are no register queues —, the own-point operands (updated fields, model
//     parameters, memory variables) as plain tiles; arrays that are neighbours in the arena travel in one box.  All of it
//     completes on one mbarrier: a thread meets exactly one wait, and the latency of a tile is hidden by the other
//     resident thread blocks (3-6 per SM), not by a software pipeline;
//   * the arithmetic is wsgen::passA / passB through the point type wsmarch::MPt in its YSM form (y stencils from shared
//     memory): the statement sequence and the accumulation order of the per-point kernels, bit-identical in FMA mode;
//   * tiles that are interior on both axes run the instantiation without boundary code (a thread-block-uniform choice).
// Halo rows / columns shared by neighbouring tiles are fetched once from HBM and again from L2 (the tiles of ~30 tile rows
// are resident at the same time).
#pragma once
#include "ws_kernels_tma.cuh"

namespace wstile {

using wsmarch::findIn;
using wsmarch::HaloSet;
using wsmarch::haloSet;
using wsmarch::Lists;
using wsmarch::spec;

constexpr int MAXOPS = 40; // boxes per tile (2-D viscoelastic, L = 4, nothing merged: 2 + 3 + 6 + 12)
struct TileOp {
    unsigned dst;       // destination inside the tile's shared memory, bytes
    unsigned char map;  // tensor map
    unsigned char slot; // first array of the box (position in the arena)
    short dx, dy;       // box origin relative to the tile origin
};
struct TileProg {
    int nOps;
    unsigned bytes;   // bytes the boxes deliver (expect_tx)
    int haloFloats;   // floats of the halo entries (the plain entries follow)
    int totalFloats;  // floats of a tile's operands
    TileOp op[MAXOPS];
};

// tile geometry in the terms wsmarch::MPt asks for (a "plane" of the marching kernels is a row here)
template <int Q, int TY_> struct Geo {
    static constexpr int NL = 4;
    static constexpr int H = Q / 2;
    static constexpr int HX = H <= 4 ? 4 : 8; // x halo rounded to whole 16-byte units
    static constexpr int TX = 128, TY = TY_, TZ = 1, HZ = 0;
    static constexpr int LDX = TX + 2 * HX;
    static constexpr int NROW = TY + 2 * H;                // rows of a halo tile
    static constexpr int TS = (LDX * NROW + 31) / 32 * 32; // floats between two halo entries (128-byte TMA destinations)
    static constexpr int NP = TX;                          // floats of a row of a plain tile
    static constexpr int LXN = TX / NL;
    static constexpr int NTHR = LXN * TY;
    static_assert(LXN == 32, "a warp per tile row");
};
struct GeoRT {
    int H, HX, TX, TY, LDX, NROW, TS, NP, NTHR;
};
inline GeoRT geoOf(int q, int ty)
{
    GeoRT g;
    g.H = q / 2;
    g.HX = g.H <= 4 ? 4 : 8;
    g.TX = 128;
    g.TY = ty;
    g.LDX = g.TX + 2 * g.HX;
    g.NROW = ty + 2 * g.H;
    g.TS = (g.LDX * g.NROW + 31) / 32 * 32;
    g.NP = g.TX;
    g.NTHR = 32 * ty;
    return g;
}

#ifndef WS_EMULATE
__device__ __forceinline__ void tmaLoad3(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <int EQ, int Q, int PASS, int TY>
__global__ void __launch_bounds__(Geo<Q, TY>::NTHR) kTile2D(const __grid_constant__ WsParams P, const __grid_constant__ TileProg prog, int tilesX)
{
    using G = Geo<Q, TY>;
    using MPG = wsmarch::MPt<EQ, 2, Q, PASS, 4, false, G, TY, true>;
    using MPI = wsmarch::MPt<EQ, 2, Q, PASS, 4, true, G, TY, true>;
    using V = FV<4>;
    constexpr Lists S = spec(EQ, 2, PASS);
    constexpr int H = G::H, NQ = S.nq;
    extern __shared__ __align__(1024) unsigned char wsTileSmem[];
    __shared__ __align__(8) uint64_t bar;
    float *sm = reinterpret_cast<float *>(wsTileSmem);

    const int tid = threadIdx.x;
    const int tyi = blockIdx.x / tilesX, txi = blockIdx.x - tyi * tilesX;
    const int tx0 = txi * G::TX, ty0 = P.ylo + tyi * TY;
    const uint32_t barA = wstma::smemAddr(&bar);
    if (tid == 0) {
        wstma::barInit(barA, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // programmatic dependent launch (ws_kernels_tile2d.cu): the thread blocks of the NEXT kernel of the step may be scheduled
        // while this grid drains, and this thread block may have been scheduled while the previous kernel drained: its first read
        // of that kernel's output is the TMA fetch below, so the fetch waits for the previous grid to complete (no-ops when the
        // kernel was launched without the attribute).  Every later access of the thread block follows the fetch's mbarrier.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        wstma::barExpectTx(barA, prog.bytes);
        const char *maps = reinterpret_cast<const char *>(P.tileMaps);
        const uint32_t smBase = wstma::smemAddr(sm);
        const int cx = WS_PADX + tx0, cy = WS_HALO + ty0;
        for (int k = 0; k < prog.nOps; k++) {
            const TileOp o = prog.op[k];
            tmaLoad3(smBase + o.dst, maps + 128 * o.map, barA, cx + o.dx, cy + o.dy, o.slot);
        }
    }
    // the thread's points: 4 x points of row `row` of the tile
    const int lx = tid & 31, row = tid >> 5;
    const int x0 = tx0 + 4 * lx, ly = ty0 + row;
    const int nAct = ly < P.yhi ? max(0, min(4, P.nx - x0)) : 0;
    const int lo = max(H, P.damping != 0 ? P.W : 0);
    const int gy0 = P.gy0 + ty0;
    const bool tileIn = tx0 >= lo && tx0 + G::TX <= P.nx - lo && gy0 >= lo && gy0 + TY <= P.gny - lo && ty0 + TY <= P.yhi;
    const int op = 4 * lx, so = 4 * lx + G::HX;
    float *st = sm + (row + H) * G::LDX, *sp = sm + prog.haloFloats + row * G::NP;
    const long long i = P.base + x0 + (long long)ly * P.plane;
    V q[NQ][Q + 1]; // no register queues here: never read (YSM)
    __syncthreads(); // the barrier is initialised before anybody polls it
    wstma::barWait(barA, 0);
    if (P.marchDebug == 1) { // developer switch: staging only (memory-side ceiling of the tiling)
    } else if (tileIn) {
        MPI t(P, x0, 0, nAct, so, op, q);
        t.setPlane(ly, i);
        t.st = st;
        t.sp = sp;
        if (PASS == 0)
            wsgen::passA<EQ, 2, false>(P, t);
        else
            wsgen::passB<EQ, 2, false>(P, t);
    } else if (nAct > 0) {
        MPG t(P, x0, 0, nAct, so, op, q);
        t.setPlane(ly, i);
        t.st = st;
        t.sp = sp;
        if (PASS == 0)
            wsgen::passA<EQ, 2, false>(P, t);
        else
            wsgen::passB<EQ, 2, false>(P, t);
    }
}
#endif // WS_EMULATE

} // namespace wstile

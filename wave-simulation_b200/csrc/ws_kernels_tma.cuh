// ws_kernels_tma.cuh — warp-specialised TMA marching kernels for EVERY equation type, FD order, edge policy and boundary
// condition (FMA arithmetic): the producer / consumer skeleton of the 3-D elastic kernels (ws_kernels_fast.cu) driven by
// the operand table `wsmarch::spec` instead of hand-placed tiles.
//
//   * a thread block owns a tile of the x-z plane (3-D: TX x TZ; 2-D: NS strips of TX points, one warp per strip) and
//     marches along y in chunks of <= 64 planes;
//   * ONE PRODUCER WARP streams the operands into a ring of NST shared-memory stages with cp.async.bulk.tensor (TMA)
//     over 4-D tensor maps (x, z, y, array) of the solver's arena; a stage holds the operands of one plane (PB = 1; the
//     layout allows groups of planes) and completes on its "full" mbarrier; which boxes make up a stage is a small
//     program built by the host from the same `spec` lists the consumers index (TmaProg), with arrays that are
//     neighbours in the arena fetched by one box;
//   * the CONSUMER WARPS never wait on DRAM and never meet at a block barrier: they wait on the full barrier of a
//     stage, compute its planes and hand it back through the "empty" barrier (one arrival per warp);
//   * the arithmetic is wsgen::passA / passB through the point type wsmarch::MPt (register queues along y, NL x points
//     per thread, interior instantiation without boundary code): bit-identical to the per-point kernels in FMA mode.
// Off-grid parts of a box (ragged last tiles, pads) are zero-filled by the TMA unit or read the zero pads of the layout.
#pragma once
#include "ws_kernels_march.cuh"

namespace wstma {

using wsmarch::findIn;
using wsmarch::ldv;
using wsmarch::Lists;
using wsmarch::spec;

constexpr int MAXOPS = 112; // boxes per stage (2-D viscoelastic, L = 4: 4 strips x 25 entries)
constexpr int CMAX = 12;    // arrays one box may fetch
struct TmaOp {
    unsigned short dst16; // destination inside a stage, in 16-byte units
    unsigned char map;    // tensor map (ws_kernels_tma.cu: kind x number of arrays)
    unsigned char slot;   // first array of the box (position in the arena)
    short dx, dz;         // box origin relative to the tile origin (minus halo, plus strip offset)
    short dy;             // plane offset (the y windows are fed q/2 planes ahead)
    short pad;
};
struct TmaProg {
    int nOps;
    unsigned stageBytes; // bytes the boxes of one stage deliver (expect_tx)
    int stageFloats;     // floats between two stages of the ring (all strips)
    int nst;             // depth of the ring
    TmaOp op[MAXOPS];
};

// tile geometry; NL = x points per thread (4: 128-bit accesses; 2 / 1: more threads per staged byte)
// rows of a 3-D tile with one x point per thread (the half-steps with ~200 B of operands per point: 3-D viscoelastic stress, 3-D viscoEM
// E): 32 x 4.  These half-steps are bound by the instruction issue of the resident warps, and shared memory decides how many are
// resident: 32 x 8 tiles leave 2 thread blocks (16 consumer warps) per SM, 32 x 4 tiles 5 (20).  BASELINE config 4 (768^3, L = 2): stress
// half-step 20.66 -> 19.37 ms, 17.7 -> 18.7 Gpt/s, although a 4-row tile fetches 3 halo rows per own row of the velocities
// (from L2; profiles/r02_tma_sweep.txt)
#ifndef WS_TMA_TZ1
#define WS_TMA_TZ1 4
#endif
template <int DIM, int Q, int NL> struct Geo {
    static constexpr int H = Q / 2;
    static constexpr int HX = H <= 4 ? 4 : 8; // x halo rounded to whole 16-byte units
    static constexpr int TX = DIM == 3 ? (NL == 4 ? 64 : 32) : 128;
    static constexpr int TZ = DIM == 3 ? (NL == 4 ? 16 : WS_TMA_TZ1) : 1;
    static constexpr int NS = DIM == 3 ? 1 : 4; // strips per thread block (2-D)
    static constexpr int PB = 1;                // planes per stage
    static constexpr int HZ = DIM == 3 ? H : 0;
    static constexpr int LDX = TX + 2 * HX, NROW = TZ + 2 * HZ;
    static constexpr int BOX = LDX * NROW;               // floats of a halo box of one plane
    static constexpr int TILE = BOX;                     // floats between the planes of a halo entry
    static constexpr int TS = (PB * BOX + 31) / 32 * 32; // floats between two halo entries: rounded to 128 bytes (TMA destinations)
    static constexpr bool HALO_MERGE = TS == PB * BOX;   // boxes of several arrays are contiguous only without the rounding
    static constexpr int NP = TX * TZ;
    static constexpr int LXN = TX / NL;
    static constexpr int NCONS = LXN * TZ * NS; // consumer threads
    static constexpr int NTHR = NCONS + 32;     // + the producer warp
    static_assert(DIM == 3 || LXN % 32 == 0, "2-D: whole warps per strip");
    static_assert(NCONS % 32 == 0 && NP % 32 == 0, "whole warps, 128-byte tiles");
};
// the same numbers for the host-side program builder
struct GeoRT {
    int H, HX, TX, TZ, NS, PB, HZ, LDX, NROW, BOX, TS, NP, NCONS, NTHR;
    bool haloMerge;
};
inline GeoRT geoOf(int dim, int q, int nl)
{
    GeoRT g;
    g.H = q / 2;
    g.HX = g.H <= 4 ? 4 : 8;
    g.TX = dim == 3 ? (nl == 4 ? 64 : 32) : 128;
    g.TZ = dim == 3 ? (nl == 4 ? 16 : WS_TMA_TZ1) : 1;
    g.NS = dim == 3 ? 1 : 4;
    g.PB = 1;
    g.HZ = dim == 3 ? g.H : 0;
    g.LDX = g.TX + 2 * g.HX;
    g.NROW = g.TZ + 2 * g.HZ;
    g.BOX = g.LDX * g.NROW;
    g.TS = (g.PB * g.BOX + 31) / 32 * 32;
    g.haloMerge = g.TS == g.PB * g.BOX;
    g.NP = g.TX * g.TZ;
    g.NCONS = (g.TX / nl) * g.TZ * g.NS;
    g.NTHR = g.NCONS + 32;
    return g;
}

#ifndef WS_EMULATE
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void barInit(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void barExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void barArrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void barWait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmaLoad(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

constexpr int NSTMAX = 8;

template <int EQ, int DIM, int Q, int PASS, int NL>
__global__ void __launch_bounds__(Geo<DIM, Q, NL>::NTHR) kTma(const __grid_constant__ WsParams P, const __grid_constant__ TmaProg prog)
{
    using G = Geo<DIM, Q, NL>;
    using MPG = wsmarch::MPt<EQ, DIM, Q, PASS, NL, false, G, G::PB>;
    using MPI = wsmarch::MPt<EQ, DIM, Q, PASS, NL, true, G, G::PB>;
    using V = FV<NL>;
    constexpr Lists S = spec(EQ, DIM, PASS);
    constexpr int H = G::H, NQ = S.nq, NT = S.nt, PB = G::PB;
    extern __shared__ __align__(1024) unsigned char wsTmaSmem[];
    __shared__ __align__(8) uint64_t bars[2 * NSTMAX];
    float *sm = reinterpret_cast<float *>(wsTmaSmem);

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * (G::NS * G::TX), tz0 = blockIdx.y * G::TZ;
    const int yc0 = P.ylo + blockIdx.z * P.tmaChunk;
    const int yc1 = min(P.yhi, yc0 + P.tmaChunk);
    if (yc0 >= yc1)
        return;
    const int nGroups = (yc1 - yc0 + PB - 1) / PB;
    const int nst = prog.nst;
    const uint32_t barFull = smemAddr(&bars[0]), barEmpty = smemAddr(&bars[NSTMAX]);
    if (tid == 0) {
        for (int s = 0; s < nst; s++) {
            barInit(barFull + 8u * s, 1);
            barInit(barEmpty + 8u * s, G::NCONS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= G::NCONS) {
        // ---- producer warp: lane k owns boxes k, k + 32, ... of the program (kept in registers), so the boxes of a stage
        // are issued side by side instead of one after the other by a single thread ----
        constexpr int OPL = (MAXOPS + 31) / 32;
        const int lane = tid - G::NCONS;
        const char *maps = reinterpret_cast<const char *>(P.tmaMaps);
        const int cx = WS_PADX + tx0, cz = (P.nzp > 1 ? WS_HALO : 0) + tz0;
        const uint32_t smBase = smemAddr(sm);
        const int nOps = prog.nOps;
        uint32_t oDst[OPL];
        const void *oMap[OPL];
        int oX[OPL], oZ[OPL], oY[OPL], oS[OPL];
#pragma unroll
        for (int j = 0; j < OPL; j++) {
            const int k = lane + 32 * j;
            const TmaOp o = prog.op[k < nOps ? k : 0];
            oDst[j] = 16u * o.dst16;
            oMap[j] = maps + 128 * o.map;
            oX[j] = cx + o.dx;
            oZ[j] = cz + o.dz;
            oY[j] = WS_HALO + yc0 + o.dy;
            oS[j] = o.slot;
        }
        int stage = 0;
        uint32_t parity = 1; // first pass over the ring: the stages are free
        for (int g = 0; g < nGroups; g++) {
            if (g >= nst)
                barWait(barEmpty + 8u * stage, parity);
            const uint32_t st = smBase + (uint32_t)stage * (uint32_t)prog.stageFloats * 4u, bar = barFull + 8u * stage;
            if (lane == 0)
                barExpectTx(bar, prog.stageBytes);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < OPL; j++)
                if (lane + 32 * j < nOps)
                    tmaLoad(st + oDst[j], oMap[j], bar, oX[j], oZ[j], oY[j] + g * PB, oS[j]);
            if (++stage == nst) {
                stage = 0;
                parity ^= 1u;
            }
        }
        return;
    }

    // ---- consumers ----
    const int strip = DIM == 2 ? tid / G::LXN : 0;
    const int tl = DIM == 2 ? tid % G::LXN : tid;
    const int lx = tl % G::LXN, lz = tl / G::LXN;
    const int sx0 = tx0 + strip * G::TX; // x origin of this thread's strip / tile
    const int x0 = sx0 + NL * lx, z = tz0 + lz;
    const int nAct = z < P.nz ? max(0, min(NL, P.nx - x0)) : 0;
    const bool active = nAct > 0;
    const int lane = tid & 31;
    const int stripFloats = prog.stageFloats / G::NS;
    // interior tile / strip: every point is at least `lo` points away from the x and z faces (no edge rows, no CPML / ABS layer)
    const int lo = max(H, P.damping != 0 ? P.W : 0);
    const bool tileIn = sx0 >= lo && sx0 + G::TX <= P.nx - lo && (DIM == 2 || (tz0 >= lo && tz0 + G::TZ <= P.nz - lo));
    const int op = lz * G::TX + NL * lx, so = (lz + G::HZ) * G::LDX + NL * lx + G::HX;

    // register queues: q[f][j] = plane ly - H + j of queued field f while plane ly is computed
    V q[NQ][Q + 1];
    const long long own0 = P.base + x0 + (long long)z * P.pitch;
#pragma unroll
    for (int f = 0; f < NQ; f++) {
        const float *src = P.fld[S.qf[f]] + own0;
        q[f][0] = V(0.0f);
#pragma unroll
        for (int j = 1; j <= Q; j++)
            q[f][j] = active ? ldv<NL>(src + (long long)(yc0 - 1 - H + j) * P.plane) : V(0.0f);
    }

    MPG tg(P, x0, z, nAct, so, op, q);
    MPI ti(P, x0, z, nAct, so, op, q);
    long long iCur = own0 + (long long)yc0 * P.plane;
    int stage = 0;
    uint32_t parity = 0;
    int ly = yc0;
    for (int g = 0; g < nGroups; g++) {
        barWait(barFull + 8u * stage, parity);
        float *sb = sm + (size_t)stage * prog.stageFloats + strip * stripFloats;
#pragma unroll
        for (int p = 0; p < PB; p++, ly++, iCur += P.plane) {
            if (PB > 1 && ly >= yc1)
                break;
            float *st = sb + p * G::TILE, *sp = sb + NT * G::TS + p * G::NP;
#pragma unroll
            for (int f = 0; f < NQ; f++) {
#pragma unroll
                for (int j = 0; j < Q; j++)
                    q[f][j] = q[f][j + 1];
                q[f][Q] = ldv<NL>(sp + f * PB * G::NP + op);
            }
            const int gy = P.gy0 + ly;
            if (P.marchDebug == 1) { // developer switch: staging only (memory-side ceiling of the skeleton)
            } else if (tileIn && gy >= lo && gy < P.gny - lo) { // uniform over the warp (2-D) / thread block (3-D)
                ti.setPlane(ly, iCur);
                ti.st = st;
                ti.sp = sp;
                if (PASS == 0)
                    wsgen::passA<EQ, DIM, false>(P, ti);
                else
                    wsgen::passB<EQ, DIM, false>(P, ti);
            } else if (active) {
                tg.setPlane(ly, iCur);
                tg.st = st;
                tg.sp = sp;
                if (PASS == 0)
                    wsgen::passA<EQ, DIM, false>(P, tg);
                else
                    wsgen::passB<EQ, DIM, false>(P, tg);
            }
        }
        // hand the stage back: the half-steps with memory variables wrote into it (write-through), so order those
        // generic-proxy writes before the producer's next asynchronous-proxy writes
        if (S.nr > 0)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            barArrive(barEmpty + 8u * stage);
        if (++stage == nst) {
            stage = 0;
            parity ^= 1u;
        }
    }
}
#endif // WS_EMULATE

} // namespace wstma

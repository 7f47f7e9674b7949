// ws_common.cuh — shared device/host definitions of the B200-native FD time-stepping library.
//
// HBM layout (private to the library): every wavefield / model array is a padded 3-D box
//     [nyl + 2*WS_HALO planes][nz + 2*WS_HALO rows (1 row in 2D)][pitch floats]
// with x fastest (reference linear index x + z*NX + y*NX*NZ, Acquisition/Coordinates.cpp:687), WS_PADX zero floats
// left of x = 0 (so row starts are 128-byte aligned) and >= WS_HALO zero floats right of x = nx-1.  The pads are
// never written by the stepping kernels, so off-grid stencil taps read exact zeros: this reproduces LAMA's
// StencilMatrix "drop off-grid taps" behaviour without any branch, and the y pads double as the ghost planes of the
// y-slab domain decomposition.
#pragma once
#include <cstdint>
#include <type_traits>
#ifdef WS_EMULATE
// tests/emu/cuda_emu.hpp: host-side stand-in for the CUDA runtime and the device intrinsics used here, so that the
// kernels and the whole C ABI can be exercised by the CPU-only test suite.  Never defined in the product build.
#include "cuda_emu.hpp"
#define WS_LAUNCH(kern, grid, block, smem, stream, ...) wsemu::launch(kern, dim3(grid), dim3(block), __VA_ARGS__)
#else
#include <cuda_runtime.h>
#define WS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include "../../include/wavesim.h"

// Launch geometry of the per-point kernels: CUDA caps gridDim.z at 65535, so a plane range longer than that is cut into
// chunks of WS_ZCHUNK planes that ride on gridDim.y next to the z blocks (gridDim.y = z blocks x chunks).
#define WS_ZCHUNK 65535
#define WS_POINT_Z(nz) ((int)((blockIdx.y % (((nz) + blockDim.y - 1) / blockDim.y)) * blockDim.y + threadIdx.y))
#define WS_POINT_PLANE(nz) ((int)((blockIdx.y / (((nz) + blockDim.y - 1) / blockDim.y)) * WS_ZCHUNK + blockIdx.z))
#define WS_HALO 6  /* max spatialFDorder / 2 (orders 2..12, Derivatives.cpp:2001-2042) */
#define WS_PADX 32 /* floats left of x = 0: one 128-byte line */
#define WS_MAXQ 12
#define WS_NOPS 8 /* xf xb yf yb zf zb yfFreeSurface ybFreeSurface */

enum { OP_XF = 0, OP_XB = 1, OP_YF = 2, OP_YB = 3, OP_ZF = 4, OP_ZB = 5, OP_YF_FS = 6, OP_YB_FS = 7 };

// wavefield slots
enum {
    F_VX = 0, F_VY, F_VZ, F_SXX, F_SYY, F_SZZ, F_SXY, F_SXZ, F_SYZ, F_P,
    F_HX, F_HY, F_HZ, F_EX, F_EY, F_EZ,
    F_R0, // memory variables: F_R0 + 6*l + c ; c = xx,yy,zz,xy,xz,yz (seismic) or c = x,y,z (EM)
    F_COUNT = F_R0 + 6 * 4
};
enum { RC_XX = 0, RC_YY, RC_ZZ, RC_XY, RC_XZ, RC_YZ, RC_X = 0, RC_Y = 1, RC_Z = 2 };

// model slots
enum {
    M_VP = 0, M_VS, M_RHO, M_TAUP, M_TAUS,                 // raw seismic
    M_PW, M_MU, M_RIX, M_RIY, M_RIZ, M_MUXY, M_MUXZ, M_MUYZ, // derived seismic
    M_TSXY, M_TSXZ, M_TSYZ, M_INVRHO,
    M_EPS, M_SIG, M_MUM, M_TAUEPS, M_TAUSIG,               // raw EM (absolute SI)
    M_MIXY, M_MIXZ, M_MIYZ,                                // inverse magnetic permeability averages
    M_CAX, M_CAY, M_CAZ, M_CBX, M_CBY, M_CBZ,
    M_CD0,                                                 // M_CD0 + 3*l + axis
    M_COUNT = M_CD0 + 3 * 4
};

// CPML memory-variable slots (18 seismic terms, CPML3D.hpp:61-79; EM terms share the table)
enum {
    PSI_SXX_X = 0, PSI_SXY_X, PSI_SXZ_X, PSI_SXY_Y, PSI_SYY_Y, PSI_SYZ_Y, PSI_SXZ_Z, PSI_SYZ_Z, PSI_SZZ_Z,
    PSI_VXX, PSI_VYX, PSI_VZX, PSI_VXY, PSI_VYY, PSI_VZY, PSI_VXZ, PSI_VYZ, PSI_VZZ,
    PSI_COUNT,
    // acoustic aliases (CPML3DAcoustic.cpp:21-56)
    PSI_P_X = PSI_SXX_X, PSI_P_Y = PSI_SYY_Y, PSI_P_Z = PSI_SZZ_Z,
    // EM aliases (CPMLEM3D.cpp:27-104)
    PSI_EYX = PSI_SXX_X, PSI_EZX = PSI_SXY_X, PSI_HYX = PSI_VXX, PSI_HZX = PSI_VYX,
    PSI_EXY = PSI_SXY_Y, PSI_EZY = PSI_SYY_Y, PSI_HXY = PSI_VXY, PSI_HZY = PSI_VYY,
    PSI_EXZ = PSI_SXZ_Z, PSI_EYZ = PSI_SYZ_Z, PSI_HXZ = PSI_VXZ, PSI_HYZ = PSI_VYZ
};

struct WsParams {
    // geometry of this rank's slab
    int nx, nyl, nz;  // local interior extent
    int gny, gy0;     // global NY and global y of local plane 0
    int pitch, nzp;   // padded row length, padded number of rows per plane
    long long plane;  // pitch * nzp
    long long base;   // offset of (x=0, y=0 local, z=0)
    int dim, eq, q, h, L;
    int free_surface, damping, W;
    int ylo, yhi;     // local y range [ylo, yhi) processed by this launch (interior/boundary split for overlap)
    int edge_policy;  // 0 truncate | 1 order-reduce
    int fastChunk, fastChunkEdge; // planes per thread block of the tiled kernels (interior / CPML-layer launch)
    int marchChunk;   // planes per thread block of the marching kernels (ws_kernels_march.cuh)
    int marchDebug;   // developer switch (env WS_MARCH_DEBUG): 1 = stage the planes but skip the arithmetic and the stores
    int marchStageR;  // 1 = the memory variables and EM Cd coefficients go through the stage ring too (0: read from global memory)
    int marchLanes;   // developer switch (env WS_MARCH_LANES): x points per thread where both variants exist, 0 = chosen per half-step
    int marchStages;  // developer switch (env WS_MARCH_STAGES): depth of the stage ring, 0 = chosen from the stage size
    int tmaChunk;     // planes per thread block of the TMA marching kernels (ws_kernels_tma.cuh)
    const void *tmaMaps; // device array of CUtensorMap (TMA marching kernels)
    const void *tileMaps; // device array of CUtensorMap (2-D tile kernels, ws_kernels_tile2d.cuh)
    int tileTY;           // rows per tile of the 2-D tile kernels
    int fastDebug;    // developer switch (env WS_FAST_DEBUG): 1 = consumers skip the arithmetic and the stores (memory-side ceiling of the tiling)
    int fastFlags;    // developer switch (env WS_FAST_FLAGS): bit 0 = L2 eviction-priority hints on the TMA loads, bit 1 = CPML-layer tiles in a launch of their own (default 2)
    const int *fastTiles; // tile list of the tiled kernels ((z tile << 16) | x tile), layer tiles first
    int fastNEdge, fastNTiles, fastTileBase;
    unsigned long long *fastTrace; // developer trace buffer (env WS_FAST_TRACE) or null
    const void *fastMaps; // device array of CUtensorMap (tiled kernels)
    // arenas of the tiled kernels: wavefields / model parameters that one TMA box fetches together are slots of one
    // allocation with a constant stride (`total` floats); null when the arrays are allocated one by one
    const float *fldArena, *matArena;
    long long arenaStride;
    // arenas of the CPML memory variables of the x / z terms (6 slabs each, tiled kernels stage them by TMA); null otherwise
    const float *psiXArena, *psiZArena;
    // x-term slabs [ly][z][psiPitchX]: row entry k' of grid column x is wsPsiXIndex(x, W, psiDX) (low side k' = x, high
    // side k' = x - psiDX with psiDX a multiple of 4, padding in between: ws_kernels_fast.cu); cxTab = coefficient rows
    // {a, b, a half, b half}[psiPitchX] in k' order, zero on the padding (tiled kernels)
    int psiPitchX, psiDX;
    int psiBoxX; // entries of an x row a tile of the tiled kernels stages (its own side of the layer, or the whole row)
    const float *cxTab;
    float cw[WS_MAXQ];  // interior weights of the plain operators, c_j * (DT/DH) (policy 0)
    float cwy[WS_MAXQ]; // interior weights of the y operators of the first half-step (image-method rows with a free surface)
    const float *tab; // derivative weight tables [WS_NOPS][2h+1][q+1], already scaled by DT/DH
    // CPML coefficients, 2W entries per array: k < W low-coordinate side, k >= W high-coordinate side (CPML3D.cpp:297-317)
    const float *cax, *cbx, *caxh, *cbxh, *cay, *cby, *cayh, *cbyh, *caz, *cbz, *cazh, *cbzh;
    const float *absCoeff; // W entries (ABS3D.cpp:174-179)
    float *psi[PSI_COUNT];
    float *fld[F_COUNT];
    const float *mat[M_COUNT];
    // free surface scalings on the y = 0 plane, dense nz*nx (FreeSurfaceElastic.cpp:35-46, FreeSurfaceViscoelastic.cpp:38-95)
    const float *sH, *sV;
    const float *sRH[4], *sRV[4];
    // viscoelastic scalars (ForwardSolver3Dviscoelastic.cpp:60-66)
    float viscoCoeff1[4], viscoCoeff2[4], invRelaxTime[4], DThalf;
    float Cc[4]; // EM relaxation (ForwardSolverEM.cpp:78-93)
    float DT;
    float fL;    // (float) L
};

// N consecutive x points handled by one thread (marching kernels: N = 4, or 1 where shared memory limits the number of
// resident threads); every operation is applied lane by lane, so a lane sees exactly the operation sequence of the
// one-point-per-thread kernels
template <int N> struct alignas(4 * N) FV {
    float v[N];
    FV() = default;
    __host__ __device__ __forceinline__ FV(float a)
    {
        for (int l = 0; l < N; l++)
            v[l] = a;
    }
};
using F4v = FV<4>;
template <class T> struct WsVecN { static constexpr int n = 0; };
template <int N> struct WsVecN<FV<N>> { static constexpr int n = N; };
template <class X, class Y, class Z = float> struct WsVecOf {
    static constexpr int n = WsVecN<X>::n > 0 ? WsVecN<X>::n : (WsVecN<Y>::n > 0 ? WsVecN<Y>::n : WsVecN<Z>::n);
    using type = FV<(n > 0 ? n : 1)>;
};
__host__ __device__ __forceinline__ float wsLane(float a, int) { return a; }
template <int N> __host__ __device__ __forceinline__ float wsLane(const FV<N> &a, int p) { return a.v[p]; }
#define WS_IF_VEC2(X, Y) typename std::enable_if<(WsVecOf<X, Y>::n > 0), int>::type = 0
#define WS_IF_VEC3(X, Y, Z) typename std::enable_if<(WsVecOf<X, Y, Z>::n > 0), int>::type = 0

// lane-wise forms of a scalar policy, for any mix of float (broadcast) and FV<N> arguments
template <class S> struct ArVec : S {
    using S::add;
    using S::madd;
    using S::msub;
    using S::mul;
    using S::sub;
    template <class X, class Y, WS_IF_VEC2(X, Y)> static __device__ __forceinline__ typename WsVecOf<X, Y>::type mul(const X &a, const Y &b)
    {
        typename WsVecOf<X, Y>::type r;
#pragma unroll
        for (int p = 0; p < WsVecOf<X, Y>::n; p++)
            r.v[p] = S::mul(wsLane(a, p), wsLane(b, p));
        return r;
    }
    template <class X, class Y, WS_IF_VEC2(X, Y)> static __device__ __forceinline__ typename WsVecOf<X, Y>::type add(const X &a, const Y &b)
    {
        typename WsVecOf<X, Y>::type r;
#pragma unroll
        for (int p = 0; p < WsVecOf<X, Y>::n; p++)
            r.v[p] = S::add(wsLane(a, p), wsLane(b, p));
        return r;
    }
    template <class X, class Y, WS_IF_VEC2(X, Y)> static __device__ __forceinline__ typename WsVecOf<X, Y>::type sub(const X &a, const Y &b)
    {
        typename WsVecOf<X, Y>::type r;
#pragma unroll
        for (int p = 0; p < WsVecOf<X, Y>::n; p++)
            r.v[p] = S::sub(wsLane(a, p), wsLane(b, p));
        return r;
    }
    template <class X, class Y, class Z, WS_IF_VEC3(X, Y, Z)>
    static __device__ __forceinline__ typename WsVecOf<X, Y, Z>::type madd(const X &a, const Y &b, const Z &c)
    {
        typename WsVecOf<X, Y, Z>::type r;
#pragma unroll
        for (int p = 0; p < WsVecOf<X, Y, Z>::n; p++)
            r.v[p] = S::madd(wsLane(a, p), wsLane(b, p), wsLane(c, p));
        return r;
    }
    template <class X, class Y, class Z, WS_IF_VEC3(X, Y, Z)>
    static __device__ __forceinline__ typename WsVecOf<X, Y, Z>::type msub(const X &a, const Y &b, const Z &c)
    {
        typename WsVecOf<X, Y, Z>::type r;
#pragma unroll
        for (int p = 0; p < WsVecOf<X, Y, Z>::n; p++)
            r.v[p] = S::msub(wsLane(a, p), wsLane(b, p), wsLane(c, p));
        return r;
    }
};

// arithmetic policy: EXACT keeps every rounding of the reference statement sequence (no FMA contraction)
struct ArExact {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(c, __fmul_rn(a, b)); } // c + a*b
    static __device__ __forceinline__ float msub(float a, float b, float c) { return __fsub_rn(c, __fmul_rn(a, b)); } // c - a*b
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
struct ArFma {
    // only the explicit multiply-adds fuse; everything else keeps its own rounding, so the result does not depend on
    // the compiler's contraction choices (general kernels, tiled kernels and the host emulation agree bit for bit)
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float madd(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float msub(float a, float b, float c) { return fmaf(-a, b, c); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <bool EXACT> struct Ar;
template <> struct Ar<true> : ArVec<ArExact> {};
template <> struct Ar<false> : ArVec<ArFma> {};

// row class of coordinate `pos` on an axis of length n: 0..h-1 low edge rows, h interior, h+1..2h high edge rows
__host__ __device__ __forceinline__ int wsRowClass(int pos, int n, int h)
{
    if (pos < h)
        return pos;
    if (pos >= n - h)
        return h + 1 + (pos - (n - h));
    return h;
}
// CPML slab index: 0..W-1 low side, W..2W-1 high side, -1 outside the layer
__host__ __device__ __forceinline__ int wsCpmlIndex(int pos, int n, int W)
{
    if (pos < W)
        return pos;
    if (pos >= n - W)
        return W + (pos - (n - W));
    return -1;
}

// grid of a per-point launch over planes [ylo, yhi) (see WS_POINT_Z / WS_POINT_PLANE)
inline void wsPointGrid(int nx, int nz, int planes, dim3 &grid, dim3 &block)
{
    block = nz > 1 ? dim3(64, 4, 1) : dim3(128, 1, 1);
    const unsigned nbz = (nz + block.y - 1) / block.y, nch = planes > 0 ? (planes + WS_ZCHUNK - 1) / WS_ZCHUNK : 0;
    grid = dim3((nx + block.x - 1) / block.x, nbz * (nch ? nch : 1), planes > WS_ZCHUNK ? WS_ZCHUNK : (planes > 0 ? planes : 0));
}

// entry of grid column x (inside an x layer) in a row of the x-term slabs
__host__ __device__ __forceinline__ int wsPsiXIndex(int x, int W, int D) { return x < W ? x : x - D; }

#define WS_CUDA_CHECK(expr)                                                                                           \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            throw WsError(WS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
    } while (0)

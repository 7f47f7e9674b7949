// ws_acquisition.cuh — source injection and receiver recording of a time step (SourceReceiverImpl.cpp:12-37, FDTD3Delastic.cpp:12-53,
// FDTD2Delastic.cpp, FDTDacoustic.cpp, ForwardSolverEM/SourceReceiverImpl/SourceReceiverImplEM.cpp) as device functions
// of the acquisition kernels of ws_api.cu.  (Doing the acquisition in the tail of the second half-step — the thread block that
// finishes last, found with a fence + counter per thread block — was measured and dropped: the per-block fence costs the 2-D tile
// kernels 16 % (2-D elastic 4096^2 73.7 -> 61.8 Gpt/s), far more than the launch it saves; profiles/r02_tile2d.txt.)
#pragma once
#include "ws_common.cuh"

struct WsAcq {
    int nsrc, nrec, nt;
    const int *srcType;
    const long long *srcOff; // padded offset, -1 if the source is not on this rank
    const float *srcSig;     // nsrc x nt
    const float *srcStep;    // nsrc samples of the current step (ws_step_host) or null
    const int *recType;
    const long long *recOff;
    float *seis;             // nrec x nt
    float *recStep;          // nrec samples of the current step
    int *tdev;               // device-resident time-step counter
};

__device__ __forceinline__ void wsInject(const WsParams &P, int type, long long off, float v)
{
    const int eq = P.eq;
    if (eq <= WS_EQ_VISCOSH) {
        switch (type) {
        case WS_TYPE_P:
            if (eq == WS_EQ_ACOUSTIC)
                P.fld[F_P][off] = __fadd_rn(P.fld[F_P][off], v);
            else {
                P.fld[F_SXX][off] = __fadd_rn(P.fld[F_SXX][off], v);
                P.fld[F_SYY][off] = __fadd_rn(P.fld[F_SYY][off], v);
                if (P.dim == 3)
                    P.fld[F_SZZ][off] = __fadd_rn(P.fld[F_SZZ][off], v);
            }
            break;
        case WS_TYPE_VX: P.fld[F_VX][off] = __fadd_rn(P.fld[F_VX][off], v); break;
        case WS_TYPE_VY: P.fld[F_VY][off] = __fadd_rn(P.fld[F_VY][off], v); break;
        case WS_TYPE_VZ: P.fld[F_VZ][off] = __fadd_rn(P.fld[F_VZ][off], v); break;
        }
    } else {
        const int slot = type == WS_TYPE_EZ ? F_EZ : (type == WS_TYPE_EX ? F_EX : (type == WS_TYPE_EY ? F_EY : F_HZ));
        P.fld[slot][off] = __fadd_rn(P.fld[slot][off], v);
    }
}

// sequential = 1: one thread applies all sources in reference order (types P,VX,VY,VZ; ascending trace) so that
// coincident sources accumulate deterministically; sequential = 0: all (target,index) pairs are distinct -> parallel.
__device__ __forceinline__ void wsSourcesSequential(const WsParams &P, const WsAcq &a, int t)
{
    for (int type = 1; type <= 4; type++)
        for (int s = 0; s < a.nsrc; s++) {
            if (a.srcType[s] != type || a.srcOff[s] < 0)
                continue;
            const float v = a.srcStep ? a.srcStep[s] : a.srcSig[(size_t)s * a.nt + t];
            wsInject(P, type, a.srcOff[s], v);
        }
}
__device__ __forceinline__ void wsSourceOne(const WsParams &P, const WsAcq &a, int t, int s)
{
    if (s >= a.nsrc || a.srcOff[s] < 0)
        return;
    const float v = a.srcStep ? a.srcStep[s] : a.srcSig[(size_t)s * a.nt + t];
    wsInject(P, a.srcType[s], a.srcOff[s], v);
}
__device__ __forceinline__ void wsReceiverOne(const WsParams &P, const WsAcq &a, int t, int r)
{
    if (r >= a.nrec || a.recOff[r] < 0)
        return;
    const long long off = a.recOff[r];
    const int type = a.recType[r];
    float v = 0.0f;
    if (P.eq <= WS_EQ_VISCOSH) {
        switch (type) {
        case WS_TYPE_P:
            if (P.eq == WS_EQ_ACOUSTIC)
                v = __fmul_rn(P.fld[F_P][off], 1.0f);
            else if (P.dim == 3) {
                v = __fadd_rn(P.fld[F_SXX][off], P.fld[F_SYY][off]);
                v = __fadd_rn(v, P.fld[F_SZZ][off]);
                v = __fdiv_rn(v, 3.0f);
            } else {
                v = __fadd_rn(P.fld[F_SXX][off], P.fld[F_SYY][off]);
                v = __fmul_rn(v, 0.5f);
            }
            break;
        case WS_TYPE_VX: v = P.fld[F_VX][off]; break;
        case WS_TYPE_VY: v = P.fld[F_VY][off]; break;
        case WS_TYPE_VZ: v = P.fld[F_VZ][off]; break;
        }
    } else {
        const int slot = type == WS_TYPE_EZ ? F_EZ : (type == WS_TYPE_EX ? F_EX : (type == WS_TYPE_EY ? F_EY : F_HZ));
        v = P.fld[slot][off];
    }
    a.seis[(size_t)r * a.nt + t] = v;
    if (a.recStep)
        a.recStep[r] = v;
}

#ifndef WS_EMULATE
// sources, receivers and the time index of a step by ONE thread block, in the order of the three kernels: all sources (block barrier
// + fence), then the receivers, then the time index
__device__ __forceinline__ void wsAcquisitionBlock(const WsParams &P, const WsAcq &a, int sequential)
{
    const int t = *a.tdev;
    if (a.nsrc > 0) {
        if (sequential) {
            if (threadIdx.x == 0)
                wsSourcesSequential(P, a, t);
        } else {
            for (int s = threadIdx.x; s < a.nsrc; s += blockDim.x)
                wsSourceOne(P, a, t, s);
        }
        __threadfence();
    }
    __syncthreads();
    for (int r = threadIdx.x; r < a.nrec; r += blockDim.x)
        wsReceiverOne(P, a, t, r);
    if (threadIdx.x == 0)
        *a.tdev = t + 1;
}
#endif

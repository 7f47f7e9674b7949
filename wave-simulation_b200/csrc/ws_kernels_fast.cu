// ws_kernels_fast.cu — tiled fast kernels (placeholder until the TMA/register-queue kernels land)
#include "ws_launch.hpp"
bool wsFastSupported(const WsParams &, bool) { return false; }
bool wsLaunchFast(const WsParams &, int, cudaStream_t) { return false; }

// ws_kernels_fast_stub.cpp — used only by the host emulation build of the test suite (tests/emu): the tiled TMA kernels
// cannot be emulated on the host, so that build always takes the general kernels.
#include "ws_launch.hpp"
bool wsFastSupported(const WsParams &, bool, int) { return false; }
void *wsFastPrepare(WsParams &, int) { return nullptr; }
void wsFastRelease(void *) {}
int wsLaunchFast(const WsParams &, int, cudaStream_t) { return 0; }

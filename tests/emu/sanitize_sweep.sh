#!/bin/bash
# TEST INFRASTRUCTURE: AddressSanitizer run of the C ABI's host code and of the per-point / marching kernels under the host emulation
# (tests/emu/cuda_emu.hpp), over every case of tests/cases.py (exact + FMA mode, run / read back / reset / run).  Builds into a scratch
# directory; the in-tree emulation library is not touched.   usage: bash tests/emu/sanitize_sweep.sh [scratch dir]
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="${1:-/tmp/wavesim_asan_emu}"
rm -rf "$OUT" && mkdir -p "$OUT"
cp "$ROOT/tests/emu/Makefile" "$ROOT/tests/emu/cuda_emu.hpp" "$OUT/"
sed -i "s#^SRC := .*#SRC := $ROOT/wave-simulation_b200/csrc#; s#../../include/wavesim.h#$ROOT/include/wavesim.h#; \
s#^CXXFLAGS := -O2#CXXFLAGS := -O1 -g -fsanitize=address -fno-omit-frame-pointer#; s#-shared -fopenmp#-shared -fsanitize=address -fopenmp#" "$OUT/Makefile"
make -s -j"$(nproc)" -C "$OUT"
cat > "$OUT/run_sweep.py" <<PY
import sys
sys.path.insert(0, "$ROOT/tests")
import numpy as np
import wsharness
from wsharness import EmuSolver
from cases import SWEEP, make_case, sweep_id, fields_of
EmuSolver.so_path = "$OUT/libwavesim_emu.so"
wsharness.build_emu = lambda force=False: EmuSolver.so_path
for cfg in SWEEP:
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    for variant, exact in ((1, 1), (1, 0), (2, 0)):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=8, exact=exact, kernel_variant=variant)
        e = case.setup(EmuSolver(case.desc))
        e.run(0, 8)
        assert np.isfinite(e.seismogram()).all()
        for f in fields_of(eq, dim, L):
            e.wavefield(f)
        e.reset()
        e.run(0, 4)
        e.close()
    print("ok", sweep_id(cfg), flush=True)
print("SWEEP DONE")
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0 python "$OUT/run_sweep.py"

"""compute-sanitizer memcheck run of every kernel family on small grids (run under `compute-sanitizer --tool memcheck --error-exitcode 1`)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from cases import make_case
from wsharness import Solver

# (eq, dim, nx, ny, nz, q, edge_policy, free_surface, damping, W, L, kernel_variant)
CASES = [
    ("elastic", 3, 64, 40, 48, 8, 0, 1, 2, 8, 0, 0),        # tiled TMA kernels (north-star configuration)
    ("elastic", 3, 64, 40, 48, 8, 1, 1, 2, 8, 0, 0),        # ... order-reducing edges
    ("elastic", 3, 52, 36, 44, 4, 0, 2, 1, 6, 0, 0),        # ABS frame, vacuum formulation: TMA marching
    ("acoustic", 3, 72, 40, 24, 8, 0, 0, 2, 8, 0, 0),       # TMA marching
    ("viscoelastic", 3, 64, 36, 40, 8, 0, 1, 2, 8, 2, 0),   # tiled velocity half-step + 32 x 4 stress kernel
    ("elastic", 2, 300, 70, 1, 8, 0, 1, 2, 10, 0, 0),       # 2-D tile kernels
    ("viscotmem", 2, 300, 60, 1, 8, 1, 0, 2, 8, 1, 0),
    ("sh", 2, 140, 50, 1, 6, 1, 1, 1, 8, 0, 0),
    ("emem", 3, 40, 30, 28, 4, 0, 0, 2, 6, 0, 2),           # cp.async marching kernels
    ("viscoemem", 3, 36, 30, 28, 8, 1, 0, 2, 6, 2, 1),      # per-point kernels
]
t0 = time.time()
for c in CASES:
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, kv = c
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=6, exact=0, kernel_variant=kv)
    s = case.setup(Solver(case.desc))
    s.run(0, 3)      # direct launches
    s.run(3, 6)
    s.sync()
    assert np.isfinite(s.seismogram()).all() and s.is_finite()
    path = s.kernel_path()
    s.reset()
    s.run(0, 2)
    s.sync()
    s.close()
    print("memcheck case %s%dD q%d pol%d fs%d damp%d L%d: kernel path %d, %.1f s" % (eq, dim, q, pol, fs, damp, L, path, time.time() - t0), flush=True)
print("MEMCHECK CASES DONE")

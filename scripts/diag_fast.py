"""Diagnostic: general vs tiled kernels after ONE step from a random wavefield state; prints mismatch locations."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import fields_of, make_case
from wsharness import Solver

def run(nx, ny, nz, q, fs, damp, W, nsteps=1):
    rng = np.random.default_rng(1)
    res = []
    init = {f: rng.standard_normal(nx * ny * nz).astype(np.float32) for f in fields_of("elastic", 3, 0)}
    for variant in (0, 1):
        case = make_case("elastic", 3, nx, ny, nz, q, 0, fs, damp, W, 0, nt=8, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        for f, a in init.items():
            s.set_wavefield(f, a)
        for t in range(nsteps):
            s.step(t)
        s.sync()
        res.append({f: s.wavefield(f).reshape(ny, nz, nx) for f in init})
        print("variant", variant, "fast", s.uses_fast_kernels())
        s.close()
    for f in init:
        a, b = res[0][f], res[1][f]
        bad = np.argwhere(a != b)
        print("%s: %d mismatches of %d, max abs diff %.3e (max |ref| %.3e)" % (f, len(bad), a.size, np.abs(a - b).max(), np.abs(b).max()))
        if len(bad):
            ys, zs, xs = bad[:, 0], bad[:, 1], bad[:, 2]
            print("   y in", np.unique(ys)[:40], "\n   z in", np.unique(zs)[:40], "\n   x in", np.unique(xs)[:70])

for shape in [(64, 40, 32, 8, 0, 0, 6), (64, 40, 32, 8, 1, 0, 6), (64, 40, 32, 8, 0, 2, 8), (48, 40, 44, 8, 1, 2, 8)]:
    for ns in (2, 6):
        print("=== shape", shape, "steps", ns)
        run(*shape, nsteps=ns)

#!/usr/bin/env python
"""BASELINE config 3 at full size: 3D acoustic FD8 1024^3, CPML 20, 10 000 time steps on one B200 (SURVEY.md 8d item 3,
"full 10k once for stability").  Prints the sum of squares of the pressure every 1000 steps and the timing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from wsharness import Solver, make_desc, ricker_np  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nt = 10000
d = make_desc(3, "acoustic", n, n, n, dh=10.0, dt=1e-3, nt=nt, fd_order=8, edge_policy=0, free_surface=0, damping=2, boundary_width=20,
              vmax_cpml=3500.0, fc_cpml=10.0, npower=4.0)
s = Solver(d)
yi = torch.arange(n, device="cuda", dtype=torch.float32).view(-1, 1, 1)
for name, t in (("velocityP", 2000.0 + 1500.0 * (yi / n)), ("density", torch.full((1, 1, 1), 2000.0, device="cuda"))):
    t = t.expand(n, n, n).contiguous()
    torch.cuda.synchronize()
    s.set_material_device(name, t.data_ptr(), t.numel())
    del t
torch.cuda.empty_cache()
s.prepare()
sig = np.zeros((1, nt), np.float32)
sig[0, :400] = ricker_np(400, 1e-3, 10.0, 1.0e6)
c = n // 2
s.set_sources64([1], np.array([c + c * n + c * n * n], dtype=np.int64), sig)
s.set_receivers64([1] * 4, np.array([c + 50 * (i + 1) + c * n + c * n * n for i in range(4)], dtype=np.int64))
s.reset()
print("3D acoustic FD8 %d^3, CPML 20, %d steps, kernels %d (3 = TMA marching)" % (n, nt, s.kernel_path()))
t0 = time.time()
for k in range(0, nt, 1000):
    s.run(k, k + 1000)
    s.sync()
    seis = s.seismogram()
    print("step %5d  finite %s  max |p| at the receivers over the last 1000 steps %.4e  (%.1f s)" % (k + 1000, s.is_finite(), np.abs(seis[:, k:k + 1000]).max(), time.time() - t0), flush=True)
dt = time.time() - t0
print("10000 steps in %.1f s incl. the per-1000-step checks: %.1f Gpt/s" % (dt, float(n) ** 3 * nt / dt / 1e9))
s.close()

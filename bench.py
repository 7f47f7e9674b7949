#!/usr/bin/env python
"""bench.py — headline benchmark of the FD time-stepping hot path (BASELINE.json: Gpt-updates/s, 3D elastic FD8).

`python bench.py --gpus N --steps K --warmup W` (torchrun for N > 1).  A "step" is one time step (velocity half-step +
stress half-step + source injection + receiver recording) over this rank's y-slab.  N = 1: 3D elastic, FD order 8,
1024^3, image-method free surface + CPML (W = 20) — the north-star configuration; N > 1: weak scaling, 1024 planes per
GPU (global NY = 1024 N).  `--impl reference` times the reference's CPU formulation (oracle: explicit CSR derivative
matrices + one SpMV / vector op per reference statement, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

EPS0, MU0 = 8.8541878176e-12, 1.2566370614e-6
# Workloads = SURVEY.md §8(d) "concrete synthetic inputs".  `bytes` = ALGORITHMIC bytes per grid-point update of the first /
# second half-step (§8(d) table).  The default (north-star) is the configuration BASELINE.json's metric is quoted on; the
# others are BASELINE.json's configs[1..4], selectable with --workload for the numbers in DESIGN.md.
WORKLOADS = {
    "northstar": dict(dim=3, eq="elastic", n=(1024, 1024, 1024), dh=10.0, dt=8e-4, fs=1, damp=2, L=0, relax=(), bytes=(60.0, 80.0),
                      vmax=5000.0, fc=10.0, src_type=3, rec_type=3, name="3D elastic FD8"),
    "cfg2": dict(dim=2, eq="elastic", n=(4096, 4096, 1), dh=5.0, dt=5e-4, fs=1, damp=2, L=0, relax=(), bytes=(36.0, 44.0),
                 vmax=3500.0, fc=10.0, src_type=3, rec_type=3, name="2D elastic FD8"),
    "cfg3": dict(dim=3, eq="acoustic", n=(1024, 1024, 1024), dh=10.0, dt=1e-3, fs=0, damp=2, L=0, relax=(), bytes=(40.0, 24.0),
                 vmax=3500.0, fc=10.0, src_type=1, rec_type=1, name="3D acoustic FD8"),
    "cfg4": dict(dim=3, eq="viscoelastic", n=(768, 768, 768), dh=10.0, dt=8e-4, fs=1, damp=2, L=2, relax=(5.0, 50.0), bytes=(60.0, 196.0),
                 vmax=3500.0, fc=10.0, src_type=3, rec_type=3, name="3D viscoelastic L=2 FD8"),
    "cfg5": dict(dim=2, eq="viscotmem", n=(8192, 2048, 1), dh=0.01, dt=1.5e-11, fs=0, damp=2, L=1, relax=(1.0e8,), bytes=(28.0, 36.0),
                 vmax=3.0e8, fc=1.0e8, src_type=1, rec_type=1, name="2D viscoTMEz L=1 FD8"),
}


def set_model_device(s, torch, wl, gny, nx, nz):
    """Synthetic models of SURVEY.md §8(d), generated on the device slab by slab (global depth index = s.y0 + local y)."""
    shape = (s.nyl, nz, nx)
    yi = torch.arange(s.y0, s.y0 + s.nyl, device="cuda", dtype=torch.float32).view(-1, 1, 1)

    def put(name, t):
        t = t.expand(*shape).contiguous()
        torch.cuda.synchronize()
        s.set_material_device(name, t.data_ptr(), t.numel())
        torch.cuda.synchronize()

    with torch.no_grad():
        if wl == "northstar":  # vp 2000..5000 m/s linear in depth, vs = vp/sqrt(3), rho = 2000 + 0.2 (vp-2000)
            vp = 2000.0 + 3000.0 * (yi / gny)
            put("velocityP", vp)
            put("velocityS", vp / float(np.sqrt(3.0)))
            put("density", 2000.0 + 0.2 * (vp - 2000.0))
        elif wl == "cfg2":  # 8 horizontal layers: vp = 1500 + 250 k, vs = vp/sqrt(3), rho = 1800 + 100 k
            k = torch.clamp(torch.floor(yi * 8.0 / gny), 0, 7)
            vp = 1500.0 + 250.0 * k
            put("velocityP", vp)
            put("velocityS", vp / float(np.sqrt(3.0)))
            put("density", 1800.0 + 100.0 * k)
        elif wl == "cfg3":  # vp = 2000 + 1500 y/NY, rho 2000
            put("velocityP", 2000.0 + 1500.0 * (yi / gny))
            put("density", torch.full((1, 1, 1), 2000.0, device="cuda"))
        elif wl == "cfg4":  # two layers, interface at 0.4 NY; tauP = tauS = 0.1
            top = (yi < 0.4 * gny).float()
            put("velocityP", 3500.0 - 1000.0 * top)
            put("velocityS", 2000.0 - 600.0 * top)
            put("density", 2300.0 - 300.0 * top)
            put("tauP", torch.full((1, 1, 1), 0.1, device="cuda"))
            put("tauS", torch.full((1, 1, 1), 0.1, device="cuda"))
        elif wl == "cfg5":  # eps_r = 4 +- 10 % in 64-cell blocks (seed 20260101), sigma 1e-3, mu_r 1, tau_eps 0.05, tau_sigma 0
            g = torch.Generator(device="cpu").manual_seed(20260101)
            blocks = torch.rand((gny + 63) // 64, (nx + 63) // 64, generator=g)
            er = 4.0 * (0.9 + 0.2 * blocks).repeat_interleave(64, 0).repeat_interleave(64, 1)[s.y0:s.y0 + s.nyl, :nx]
            put("dielectricPermittivity", (EPS0 * er).to("cuda").view(s.nyl, 1, nx))
            put("electricConductivity", torch.full((1, 1, 1), 1.0e-3, device="cuda"))
            put("magneticPermeability", torch.full((1, 1, 1), MU0, device="cuda"))
            put("tauDielectricPermittivity", torch.full((1, 1, 1), 0.05, device="cuda"))
            put("tauElectricConductivity", torch.full((1, 1, 1), 0.0, device="cuda"))
        torch.cuda.empty_cache()


def acquisition(wl, nx, gny, nz):
    """(source index, receiver indices) as 64-bit global linear indices x + z NX + y NX NZ"""
    pl = nx * nz
    if wl == "cfg2":
        return [nx // 2 + 1 * pl], [(nx // 4 + 4 * i) + 1 * pl for i in range(min(512, nx // 8))]
    if wl == "cfg3":
        ry = min(32, gny - 1)
        return [nx // 2 + (nz // 2) * nx + (gny // 2) * pl], [i + (nz // 2) * nx + ry * pl for i in range(min(1024, nx))]
    if wl == "cfg5":
        xs = min(512, nx // 4)
        return [xs + 21 * pl], [min(xs + 8 * i, nx - 1) + 21 * pl for i in range(128)]
    nrec = min(1024, nx)
    return [(nx // 2) + (nz // 2) * nx + 1 * pl], [(nx // 2 - nrec // 2 + i) + (nz // 2) * nx + 1 * pl for i in range(nrec)]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(kernel, nx, ny, nz, damping, free_surface):
    """DRAM bytes per half-step (dram__bytes_read.sum + dram__bytes_write.sum over the launches of that half-step) from the
    committed ncu capture of this very configuration (profiles/traffic_1024.json, written by scripts/exp.sh); None for
    any other configuration."""
    if (nx, ny, nz, damping, free_surface) != (1024, 1024, 1024, 2, 1):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_1024.json")) as f:
            t = json.load(f)["per_half_step"]["str" if kernel == 1 else "vel"]
        return float(t["dram_read_bytes"] + t["dram_write_bytes"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def host_cores():
    """Physical cores this process may use (affinity mask and hyper-threading taken into account)."""
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count() or 1
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or aff
    except Exception:
        phys = aff
    return max(1, min(aff, phys))


def cpu_worker(spec):
    """Child process of cpu_reference_sample: the reference CPU formulation (oracle CSR restatement of the LAMA path) on a
    bounded sample of the north-star workload — 3D elastic FD8, free surface + CPML, n^3 grid.  `repeats` timed blocks of
    `steps` steps after `warm` warm-up steps; the best block counts (pinned threads, best of 3)."""
    from wsharness import Oracle, make_desc, idx1d, ricker_np
    n, steps, warm, repeats = (int(v) for v in spec.split(","))
    cores = int(os.environ.get("OMP_NUM_THREADS", "1"))
    Oracle.num_threads()  # loads the library
    Oracle.lib.wso_set_threads(cores)
    Oracle.lib.wso_set_flush_denormals(1)
    nt = warm + repeats * steps
    d = make_desc(3, "elastic", n, n, n, dh=10.0, dt=8e-4, nt=nt, fd_order=8, edge_policy=0, free_surface=1, damping=2,
                  boundary_width=min(20, n // 4), vmax_cpml=5000.0, fc_cpml=10.0, npower=4.0)
    o = Oracle(d)
    y = np.arange(n, dtype=np.float32)[:, None, None] / n
    vp = np.broadcast_to(2000.0 + 3000.0 * y, (n, n, n)).astype(np.float32).ravel()
    o.set_material("velocityP", vp)
    o.set_material("velocityS", (vp / np.float32(np.sqrt(3.0))).astype(np.float32))
    o.set_material("density", (2000.0 + 0.2 * (vp - 2000.0)).astype(np.float32))
    o.prepare()
    o.set_sources([3], [idx1d(n // 2, 1, n // 2, n, n)], ricker_np(nt, d.dt, 10.0, 1.0)[None, :])
    o.set_receivers([3] * 8, [idx1d(n // 4 + 4 * i, 1, n // 2, n, n) for i in range(8)])
    o.reset()
    o.run(0, warm)
    best = None
    for r in range(repeats):
        t0 = time.perf_counter()
        o.run(warm + r * steps, warm + (r + 1) * steps)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # second CPU number: the fused matrix-free back-end of the oracle on the same sample (bit-identical results)
    best_fused = None
    try:  # (the first number must survive whatever happens to the second)
        o.set_fused(True)
        o.reset()
        o.run(0, warm)
        for r in range(repeats):
            t0 = time.perf_counter()
            o.run(warm + r * steps, warm + (r + 1) * steps)
            dt = time.perf_counter() - t0
            best_fused = dt if best_fused is None else min(best_fused, dt)
    except Exception as exc:
        print("fused back-end not timed: %s" % exc, file=sys.stderr, flush=True)
        best_fused = None
    print(json.dumps({"gpts": float(n) ** 3 * steps / best / 1e9, "sec_step": best / steps, "cores": Oracle.num_threads(), "n": n,
                      "steps": steps, "repeats": repeats, "gpts_fused": float(n) ** 3 * steps / best_fused / 1e9 if best_fused else None}), flush=True)


def cpu_reference_sample(n, steps, warm=1, repeats=3):
    """The CPU arm, in a fresh process whose OpenMP environment is set HERE: all physical cores the process may use, one
    pinned thread per core — whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1).  The same sample
    serves `cpu_baseline` of the GPU line and the `--impl reference` line."""
    env = {k: v for k, v in os.environ.items() if not k.startswith(("OMP_", "GOMP_", "KMP_", "MKL_"))}
    cores = host_cores()
    env.update(OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close", OMP_PLACES="cores", OMP_DYNAMIC="false")
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", "%d,%d,%d,%d" % (n, steps, warm, repeats)], env=env, capture_output=True,
                         text=True, timeout=1700)
    if out.returncode != 0:
        raise RuntimeError("CPU reference sample failed: " + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["fused"] = None if r.get("gpts_fused") is None else {
        "value": r["gpts_fused"], "unit": "Gpt/s", "cores": r["cores"], "kind": "port-fused",
        "sample": "same sample and threads; fused matrix-free back-end of the oracle (1-D coefficient rows, one pass per half-step, "
                  "bit-identical to the matrix formulation): the stronger CPU number of SURVEY.md 8(d), not the reference's formulation"}
    r["sample"] = ("%d^3 grid of the 1024^3 workload, best of %d blocks of %d steps, %d pinned threads (one per physical core), FTZ/DAZ; oracle CSR "
                   "formulation (restatement of the LAMA sparse path, not the LAMA binary)" % (r["n"], r["repeats"], r["steps"], r["cores"]))
    return r


def cpu_sample_size(n, steps, warm, repeats=3):
    """bound the run: the CSR formulation moves ~2 kB per grid point and step (about 8e6 point-steps per second and core pair)"""
    while n > 64 and 1.35 * (repeats * steps + warm) * (n ** 3) / 8.0e6 > 240.0:  # (x 1.35: the fused back-end is timed on the same sample)
        n -= 32
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 20)), max(1, args.warmup)
    n = cpu_sample_size(args.ref_n, steps, warm)
    r = cpu_reference_sample(n, steps, warm)
    gpts = r["gpts"]
    line = {
        "impl": "reference", "metric": "Gpt-updates/s", "value": gpts, "unit": "Gpt/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": r["sec_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D elastic FD8, free surface + CPML(20), synthetic gradient model; CPU sample %d^3 of the 1024^3 workload" % n},
        "cpu_baseline": {"value": gpts, "unit": "Gpt/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": gpts, "unit": "Gpt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if r["fused"] is not None:
        line["cpu_baseline_fused"] = r["fused"]
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="northstar", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0, help="planes PER GPU")
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--ref-n", type=int, default=192)
    ap.add_argument("--cpu-worker", default="", help="internal: child process of the CPU arm (n,steps,warm,repeats)")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of BASELINE configs 2-5 appended to the default line")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --ny is the GLOBAL number of planes, cut into y-slabs over the GPUs")
    ap.add_argument("--variant", type=int, default=0, help="0 auto (TMA kernels, else marching kernels), 1 per-point kernels, 2 marching kernels")
    ap.add_argument("--damping", type=int, default=-1, help="developer switch: 2 = CPML (the benchmark configuration), 0 = none")
    ap.add_argument("--free-surface", type=int, default=-1, help="developer switch: 1 = image method")
    ap.add_argument("--edge-policy", type=int, default=0, help="developer switch: 1 = order-reducing edges (useStencilMatrix=0, the par/ default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.nx, args.ny, args.nz = args.nx or wl["n"][0], args.ny or wl["n"][1], (args.nz or wl["n"][2]) if wl["dim"] == 3 else 1
    args.damping = wl["damp"] if args.damping < 0 else args.damping
    args.free_surface = wl["fs"] if args.free_surface < 0 else args.free_surface
    if args.cpu_worker:
        return cpu_worker(args.cpu_worker)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    W, K = max(3, args.warmup), max(1, args.steps)
    sampler = ClockSampler(local)
    sampler.start()
    m = measure(args.workload, args.nx, args.ny, args.nz, K, W, args.variant, args.damping, args.free_surface, world, rank, local, e2e=True, strong=args.strong, edge_policy=args.edge_policy)
    sampler.stop_flag = True
    sampler.join()
    # BASELINE configs 2-5 (short runs after the headline's timed region: driver-run numbers for the other solvers)
    others = None
    if world == 1 and args.workload == "northstar" and not args.no_others and (args.nx, args.ny, args.nz) == WORKLOADS["northstar"]["n"]:
        others = {}
        for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
            o = WORKLOADS[name]
            try:
                ko = 200 if o["dim"] == 2 else 10  # (a 2-D step lasts 0.2 ms: 200 steps are 50 ms and fill whole graph batches)
                r = measure(name, o["n"][0], o["n"][1], o["n"][2], ko, 3, args.variant, o["damp"], o["fs"], 1, 0, local, e2e=False)
                others[name] = {"workload": r["workload"], "value": r["value"], "unit": "Gpt/s", "steps": ko, "warmup": 3, "ms_per_step": r["ms_per_step"],
                                "kernels": r["kernels"], "frac": r["roofline"]["frac"], "whole_step_frac": r["roofline"]["whole_step_frac"],
                                "ms_first": r["roofline"]["ms_first"], "ms_second": r["roofline"]["ms_second"], "finite": r["finite"]}
            except Exception as exc:  # the headline must survive
                others[name] = {"error": str(exc)[:300]}
    parity = parity_multi(world, rank, local) if world > 1 else None
    if rank == 0:
        cpu = cpu_fused = None
        if not args.no_cpu and world == 1 and args.workload == "northstar":
            ck, cw = max(1, min(K, 20)), max(1, args.warmup)
            r = cpu_reference_sample(cpu_sample_size(args.ref_n, ck, cw), ck, cw)
            cpu = {"value": r["gpts"], "unit": "Gpt/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
            cpu_fused = r["fused"]
        cfg = {"workload": m["workload"], "l2": m["l2"], "kernels": m["kernels"], "finite": m["finite"], "halo": m["halo"]}
        if others is not None:
            cfg["others"] = others
        line = {
            "metric": "Gpt-updates/s", "value": m["value"], "unit": "Gpt/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "roofline": m["roofline"], "cpu_baseline": cpu, "e2e": m["e2e"],
            "gpu_launches": m["launches"], "clocks": sampler.summary(),
        }
        if cpu_fused is not None:
            line["cpu_baseline_fused"] = cpu_fused
        if parity is not None:
            line["parity_multi"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure(wlname, nx, nyl, nz, K, W, variant, damping, free_surface, world, rank, local, e2e=True, strong=False, edge_policy=0):
    """W warm-up steps, exactly K timed steps of workload `wlname` on this rank's slab (nyl planes per GPU); max over ranks."""
    import torch
    import torch.distributed as dist
    from wsharness import Solver, make_desc, ricker_np
    wl = WORKLOADS[wlname]
    nz = nz if wl["dim"] == 3 else 1
    gny = nyl if strong else nyl * world
    nt = W + 3 * K + 8
    dt_, dh = wl["dt"], wl["dh"]
    d = make_desc(wl["dim"], wl["eq"], nx, gny, nz, dh=dh, dt=dt_, nt=nt, fd_order=8, edge_policy=edge_policy, free_surface=free_surface, damping=damping,
                  boundary_width=20, vmax_cpml=wl["vmax"], fc_cpml=wl["fc"], npower=4.0, relax_freq=wl["relax"], exact_arith=0,
                  kernel_variant=variant, rank=rank, nranks=world, device=local)
    s = Solver(d)
    if world > 1:
        ids = [Solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s.comm_init(ids[0])
    set_model_device(s, torch, wlname, gny, nx, nz)
    s.prepare()
    src, recs = acquisition(wlname, nx, gny, nz)
    nrec = len(recs)
    amp = 1.0e6 if not wl["eq"].endswith("mem") else 1.0
    sig = ricker_np(nt, dt_, wl["fc"], amp)
    s.set_sources64([wl["src_type"]], np.array(src, dtype=np.int64), sig[None, :])
    s.set_receivers64([wl["rec_type"]] * nrec, np.array(recs, dtype=np.int64))
    s.reset()

    stream = torch.cuda.ExternalStream(s.stream_ptr())

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident run: W warm-up steps, then exactly K timed steps as the product runs them (ws_run: CUDA-graph batches, no
    # event between the kernels) --------------------------------------------------------------------------------------------------
    # 3-D workloads (29 ms steps): CUDA events around both half-step kernels of every step INSIDE the timed region (direct launches).
    # 2-D workloads (0.2 ms steps): the events and the graph-less launches cost ~10 us per step, so the timed region runs the steps
    # as the product does (ws_run: CUDA-graph batches, no event between the kernels) and the per-kernel durations of the roofline
    # come from K further steps with events, right after.
    split = wl["dim"] == 2
    s.set_timing(False)
    s.run(0, W)
    barrier()
    l0 = s.launch_count()
    s.set_timing(not split)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s.run(W, W + K)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count() - l0
    if split:
        s.set_timing(True)
        s.run(W + K, W + 2 * K)
        barrier()
    msA, msB = s.last_timing(0), s.last_timing(1)
    # ---- end-to-end run through host buffers: per step H2D of the source samples, D2H of the receiver samples --------
    s.set_timing(False)
    e2e_s = 0.0
    if e2e:
        rec = np.zeros(nrec, np.float32)
        t_base = W + 2 * K
        for t in range(t_base, t_base + 3):
            s.step_host(t, sig[t:t + 1], rec)
        barrier()
        t0 = time.perf_counter()
        for t in range(t_base + 3, t_base + 3 + K):
            s.step_host(t, sig[t:t + 1], rec)
        barrier()
        e2e_s = time.perf_counter() - t0
    finite = s.is_finite()

    t_ms = torch.tensor([ms, e2e_s * 1e3, msA, msB], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms, e2e_ms, msA, msB = (float(v) for v in t_ms.tolist())
    npts_local = float(nx) * s.nyl * nz  # (strong scaling: the first ranks hold the remainder planes, i.e. the largest slab)
    ws_bytes = s.estimate_memory()
    npts = float(nx) * gny * nz
    nyl = s.nyl
    peak, which = measured_peak()
    bA, bB = wl["bytes"]
    dom = 1 if msB >= msA else 0
    achieved = (bB if dom else bA) * npts_local / ((msB if dom else msA) * 1e-3) / 1e9
    traffic = measured_traffic(dom, nx, nyl, nz, damping, free_surface) if wlname == "northstar" else None
    out = {
        "workload": "%s %dx%dx%d per GPU (global NY %d), %sCPML(20), y-slab decomposition" % (wl["name"], nx, nyl, nz, gny, "free surface + " if free_surface else ""),
        "l2": "inputs (%.1f GB/GPU of wavefields+model) exceed the 126 MB L2" % (ws_bytes / 1e9),
        "kernels": ["per-point", "marching", "tma-tiled", "tma-marching", "tma-tile2d"][s.kernel_path()], "finite": bool(finite),
        "halo": ["none", "nccl send/recv", "external", "kernels over peer memory"][s.halo_transport()],
        "value": npts * K / (ms * 1e-3) / 1e9, "ms_per_step": ms / K, "launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "committed ncu capture of this configuration (profiles/traffic_1024.json), not measured in this run" if traffic else None,
                     "kernel": "second half-step (stress / E)" if dom else "first half-step (velocity / H)", "peak_source": which,
                     "kernel_timing": ("CUDA events around both half-step kernels of K further steps right after the timed region (2-D: the events cost ~10 us per 0.2 ms step)"
                                       if split else "CUDA events around both half-step kernels of every timed step"),
                     "ms_first": msA, "ms_second": msB, "whole_step_frac": (bA + bB) * npts_local / (ms / K * 1e-3) / 1e9 / peak},
        "e2e": {"value": npts * K / (e2e_ms * 1e-3) / 1e9, "unit": "Gpt/s", "h2d_bytes_per_step": 4 * world, "d2h_bytes_per_step": 4 * nrec} if e2e else None,
    }
    s.close()
    torch.cuda.empty_cache()
    return out


def parity_multi(world, rank, local):
    """Driver-visible multi-GPU correctness: a small 3D elastic case (free surface + CPML, TMA kernels) on the `world` ranks
    of this run against the same case on rank 0 alone; "bit-identical" or the worst relative L2 difference."""
    import torch.distributed as dist
    from cases import make_case
    from wsharness import Solver, rel_l2
    nx, ny, nz, nt = 128, 48 * world, 48, 24
    fields = ["VX", "VY", "VZ", "Sxx", "Syy", "Szz", "Sxy", "Sxz", "Syz"]
    case = make_case("elastic", 3, nx, ny, nz, 8, 0, 1, 2, W=10, L=0, nt=nt, exact=0, kernel_variant=0)
    case.desc.rank, case.desc.nranks, case.desc.device = rank, world, local
    s = Solver(case.desc)
    ids = [Solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s.comm_init(ids[0])
    case.setup(s)
    s.run(0, nt)
    s.sync()
    mine = (s.y0, s.nyl, s.seismogram(), {f: s.wavefield(f) for f in fields}, s.kernel_path())
    s.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    case.desc.rank, case.desc.nranks = 0, 1
    ref = case.setup(Solver(case.desc))
    ref.run(0, nt)
    ref.sync()
    worst, same = 0.0, True
    seis = np.zeros_like(ref.seismogram())
    plane = nx * nz
    for y0, nyl, sg, fl, path in parts:
        seis += sg
        for f in fields:
            a, b = fl[f], ref.wavefield(f)[y0 * plane:(y0 + nyl) * plane]
            if not np.array_equal(a, b):
                same = False
                worst = max(worst, rel_l2(a, b))
    if not np.array_equal(seis, ref.seismogram()):
        same = False
        worst = max(worst, rel_l2(seis, ref.seismogram()))
    ref.close()
    return "bit-identical" if same else worst


if __name__ == "__main__":
    main()

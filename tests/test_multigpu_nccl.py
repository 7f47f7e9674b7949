"""y-slab decomposition on real GPUs (needs >= 2 devices, skipped otherwise), with both halo transports — the library's own kernels
over peer memory (CUDA IPC between the processes here; the default where the GPUs can map each other) and ncclSend / ncclRecv
(WS_P2P=0): every kernel family, result bit-identical to the single-GPU run (every grid point sees the same arithmetic whatever
the decomposition)."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from cases import fields_of, make_case

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, cfg, nt, variant, out):
    from wsharness import Solver
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=0, kernel_variant=variant)
    case.desc.rank, case.desc.nranks, case.desc.device = rank, world, rank
    s = Solver(case.desc)
    s.comm_init(uid)
    case.setup(s)
    s.run(0, nt)
    s.sync()
    assert s.is_finite()
    out.put((rank, s.y0, s.nyl, s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}, s.halo_transport()))
    s.close()


CASES = [
    (("elastic", 3, 128, 96, 48, 8, 0, 1, 2, 10, 0), 0),   # tiled kernels, free surface + CPML
    (("elastic", 3, 64, 40, 24, 8, 1, 1, 2, 6, 0), 1),     # general kernels, order-reducing edges
    (("viscoelastic", 2, 96, 120, 1, 6, 1, 1, 2, 8, 2), 0),
    (("acoustic", 3, 48, 64, 40, 4, 0, 0, 1, 8, 0), 0),
    (("elastic", 3, 128, 96, 48, 8, 0, 1, 2, 10, 0), 2),   # marching kernels forced (4 x points per thread)
    (("viscoelastic", 3, 72, 64, 24, 8, 0, 1, 2, 8, 2), 0),  # marching kernels, one x point per thread in the stress half-step
    (("viscoemem", 3, 40, 64, 24, 4, 1, 0, 2, 6, 1), 0),
    (("elastic", 2, 300, 200, 1, 8, 0, 1, 2, 10, 0), 0),    # 2-D tile kernels
    (("viscotmem", 2, 260, 130, 1, 8, 1, 0, 1, 8, 2), 0),
]


def _peer_access(world):
    return all(torch.cuda.can_device_access_peer(a, b) for a in range(world) for b in range(world) if a != b)


@pytest.mark.parametrize("cfg,variant", CASES, ids=["%s%dD-v%d" % (c[0][0], c[0][1], c[1]) for c in CASES])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_slabs_equal_single_gpu(cfg, variant, world, transport, monkeypatch):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    if transport == "nccl":
        monkeypatch.setenv("WS_P2P", "0")
    else:
        monkeypatch.delenv("WS_P2P", raising=False)
    from wsharness import Solver
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    nt = 30
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=0, kernel_variant=variant)
    ref = case.setup(Solver(case.desc))
    ref.run(0, nt)
    ref.sync()
    ref_seis = ref.seismogram()
    ref_fields = {f: ref.wavefield(f) for f in fields_of(eq, dim, L)}
    ref.close()
    uid = Solver.comm_unique_id()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, cfg, nt, variant, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((out.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plane = nx * (nz if dim == 3 else 1)
    seis = np.zeros_like(ref_seis)
    for rank, y0, nyl, sg, fields, tr in results:
        assert tr == (1 if transport == "nccl" or not _peer_access(world) else 3), tr
        seis += sg  # rows of receivers on other ranks are zero
        for f, a in fields.items():
            assert np.array_equal(a, ref_fields[f][y0 * plane:(y0 + nyl) * plane]), (rank, f)
    assert np.abs(ref_seis).max() > 0
    assert np.array_equal(seis, ref_seis)


def test_two_handles_on_two_devices_in_one_process():
    """One process, two solver handles on two GPUs (what host/Simulation does for NumShotDomains > 1): the dynamic
    shared-memory opt-in of the TMA / marching kernels is a per-device attribute and must reach both devices."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from wsharness import Solver
    for cfg, variant in ((("elastic", 3, 128, 40, 48, 8, 0, 1, 2, 10, 0), 0), (("viscoelastic", 3, 72, 40, 24, 8, 0, 1, 2, 8, 2), 0), (("acoustic", 2, 300, 80, 1, 8, 1, 0, 2, 8, 0), 0)):
        eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
        nt = 20
        res = []
        solvers = []
        for dev in (0, 1):
            case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=0, kernel_variant=variant)
            case.desc.device = dev
            solvers.append(case.setup(Solver(case.desc)))
        for s in solvers:  # both GPUs work concurrently
            s.run(0, nt)
        for s in solvers:
            s.sync()
            assert s.is_finite()
            res.append((s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}, s.kernel_path()))
            s.close()
        assert res[0][2] == res[1][2] and res[0][2] >= 1
        assert np.abs(res[0][0]).max() > 0 and np.array_equal(res[0][0], res[1][0])
        for f in res[0][1]:
            assert np.array_equal(res[0][1][f], res[1][1][f]), f


def _worker_reset(rank, world, uid, cfg, nt, out):
    from wsharness import Solver
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=0, kernel_variant=0)
    case.desc.rank, case.desc.nranks, case.desc.device = rank, world, rank
    s = Solver(case.desc)
    s.comm_init(uid)
    case.setup(s)
    s.run(0, nt)
    s.reset()  # the last halo exchange of the first shot must not land in the freshly zeroed ghost planes
    s.run(0, nt)
    s.sync()
    out.put((rank, s.y0, s.nyl, s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}))
    s.close()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_run_reset_run_on_two_ranks_equals_single_gpu(transport, monkeypatch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    if transport == "nccl":
        monkeypatch.setenv("WS_P2P", "0")
    else:
        monkeypatch.delenv("WS_P2P", raising=False)
    from wsharness import Solver
    cfg = ("elastic", 3, 128, 96, 48, 8, 0, 1, 2, 10, 0)
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    nt = 30
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=0, kernel_variant=0)
    ref = case.setup(Solver(case.desc))
    ref.run(0, nt)
    ref.sync()
    ref_seis, ref_fields = ref.seismogram(), {f: ref.wavefield(f) for f in fields_of(eq, dim, L)}
    ref.close()
    uid = Solver.comm_unique_id()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker_reset, args=(r, 2, uid, cfg, nt, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((out.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plane = nx * nz
    seis = np.zeros_like(ref_seis)
    for rank, y0, nyl, sg, fields in results:
        seis += sg
        for f, a in fields.items():
            assert np.array_equal(a, ref_fields[f][y0 * plane:(y0 + nyl) * plane]), (rank, f)
    assert np.array_equal(seis, ref_seis)

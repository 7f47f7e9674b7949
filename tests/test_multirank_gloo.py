"""N > 1 path on CPU: world_size-2 (and 3) `gloo` runs of the y-slab decomposition — slab ranges, ghost
plane exchange after each half-step, owner-computes acquisition, global model upload — using the host emulation build of
the PRODUCT sources (tests/emu) and the bring-your-own-transport hook of the C ABI (ws_comm_init_external) wired to
torch.distributed send/recv.  Result must be bit-identical to the single-rank run."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from cases import fields_of, make_case

CASES = [
    # eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, world
    ("elastic", 3, 16, 26, 14, 8, 0, 1, 2, 4, 0, 2),
    ("acoustic", 2, 30, 40, 1, 4, 1, 1, 2, 6, 0, 2),
    ("viscoelastic", 2, 24, 37, 1, 6, 1, 1, 1, 5, 2, 3),
    ("viscotmem", 2, 26, 36, 1, 4, 0, 0, 2, 6, 1, 2),
    ("elastic", 2, 30, 40, 1, 8, 1, 2, 2, 6, 0, 2),  # FreeSurface = 2: no frame at the top, plain operators on the first slab
]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg, nt, out):
    import torch
    import torch.distributed as dist
    from wsharness import EmuSolver
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=1)
        case.desc.rank, case.desc.nranks = rank, world
        e = EmuSolver(case.desc)

        def sendrecv(send, recv, peer):
            ts, tr = torch.from_numpy(send), torch.from_numpy(recv)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, ts, peer), dist.P2POp(dist.irecv, tr, peer)])
            for r in reqs:
                r.wait()

        e.comm_init_external(sendrecv)
        case.setup(e)
        e.run(0, nt)
        seis = torch.from_numpy(e.seismogram())  # rows of receivers on other ranks stay zero
        dist.all_reduce(seis)
        fields = {f: e.wavefield(f) for f in fields_of(eq, dim, L)}
        out.put((rank, e.y0, e.nyl, seis.numpy().copy(), fields))
        e.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", CASES, ids=["%s%dD-q%d-fs%d-w%d" % (c[0], c[1], c[5], c[7], c[11]) for c in CASES])
def test_yslab_decomposition_equals_single_rank(cfg):
    from wsharness import EmuSolver
    world, cfg = cfg[11], cfg[:11]
    nt = 24
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=nt, exact=1)
    ref = case.setup(EmuSolver(case.desc))
    ref.run(0, nt)
    ref_seis = ref.seismogram()
    ref_fields = {f: ref.wavefield(f) for f in fields_of(eq, dim, L)}
    assert np.abs(ref_seis).max() > 0

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, nt, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    plane = nx * (nz if dim == 3 else 1)
    covered = 0
    for rank, y0, nyl, seis, fields in results:
        assert y0 == covered  # block distribution, contiguous
        covered += nyl
        assert np.array_equal(seis, ref_seis)
        for f, a in fields.items():
            assert np.array_equal(a, ref_fields[f][y0 * plane:(y0 + nyl) * plane]), (rank, f)
    assert covered == ny

"""GPU parity tests proper: the CUDA path, called through the C ABI (include/wavesim.h), against the CPU oracle on the
same seeded inputs, plus the reference's golden seismograms and size-independent properties at larger sizes.
Bar (BASELINE.json north_star): relative L2 seismogram misfit <= 1e-5 over the full trace; in exact-arithmetic mode the
CUDA kernels keep the reference's operation order and must be bit-identical to the oracle."""
import numpy as np
import pytest

from cases import SWEEP, fields_of, make_case, sweep_id
from wsharness import Oracle, Solver, ci_case, golden, reference_gate, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1.0e-5  # relative L2, fp32


def run_pair(case, nt):
    o = case.setup(Oracle(case.desc))
    s = case.setup(Solver(case.desc))
    o.run(0, nt)
    s.run(0, nt)
    s.sync()
    return o, s


@pytest.mark.parametrize("cfg", SWEEP, ids=[sweep_id(c) for c in SWEEP])
def test_general_kernels_exact_mode_bit_identical(cfg):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=30, exact=1, kernel_variant=1)
    o, s = run_pair(case, 30)
    so, ss = o.seismogram(), s.seismogram()
    assert np.abs(so).max() > 0
    assert np.array_equal(so, ss)
    for f in fields_of(eq, dim, L):
        assert np.array_equal(o.wavefield(f), s.wavefield(f)), f
    assert s.launch_count() > 0


@pytest.mark.parametrize("cfg", SWEEP, ids=[sweep_id(c) for c in SWEEP])
def test_default_mode_within_tolerance(cfg):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=40, exact=0, kernel_variant=0)
    o, s = run_pair(case, 40)
    assert rel_l2(s.seismogram(), o.seismogram()) <= TOL
    for f in fields_of(eq, dim, L):
        a, b = o.wavefield(f), s.wavefield(f)
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-30), f


@pytest.mark.parametrize("name", ["2D.acoustic", "2D.sh", "2D.elastic", "2D.visco", "3D.acoustic", "3D.elastic", "3D.visco"])
def test_reference_ci_cases_full_trace(name):
    """par/ci cases, all 1000 samples, CUDA vs the reference's golden trace (and the reference's own CI gate)."""
    case = ci_case(name)
    case.desc.edge_policy = 0
    s = case.setup(Solver(case.desc))
    s.run(0, 1000)
    s.sync()
    g = golden(case.golden)
    ss = s.seismogram()
    assert rel_l2(ss, g) <= 1.0e-5  # golden files carry 6 significant digits
    assert reference_gate(ss, g) <= 5.0e-7
    assert s.is_finite()


@pytest.mark.parametrize("name", ["2D.elastic", "2D.visco", "2D.sh"])
def test_reference_ci_cases_vs_oracle(name):
    case = ci_case(name)
    o, s = run_pair(case, 1000)
    assert rel_l2(s.seismogram(), o.seismogram()) <= TOL


def test_3d_elastic_ci_case_vs_oracle_prefix():
    case = ci_case("3D.elastic", nt=300)
    case.desc.exact_arith = 1
    case.desc.kernel_variant = 1
    o, s = run_pair(case, 300)
    assert np.array_equal(o.seismogram(), s.seismogram())


def test_graph_replay_equals_single_steps():
    case = make_case("elastic", 3, 24, 26, 22, 8, 0, 1, 2, 6, 0, nt=37, exact=1, kernel_variant=1)
    a = case.setup(Solver(case.desc))
    b = case.setup(Solver(case.desc))
    a.run(0, 37)  # graph-batched
    for t in range(37):
        b.step(t)
    a.sync()
    b.sync()
    assert np.array_equal(a.seismogram(), b.seismogram())


def test_step_host_path():
    case = make_case("acoustic", 3, 24, 26, 22, 4, 0, 1, 2, 6, 0, nt=16)
    a = case.setup(Solver(case.desc))
    a.run(0, 16)
    ref = a.seismogram()
    a.reset()
    rec = np.zeros(4, np.float32)
    got = np.zeros_like(ref)
    for t in range(16):
        a.step_host(t, np.ascontiguousarray(case.src[2][:, t]), rec)
        got[:, t] = rec
    assert np.array_equal(got, ref)


def test_linearity_and_reset_large():
    """Size-independent properties at a size the oracle would not finish quickly: doubling every source doubles the
    seismogram exactly (power-of-two scaling is exact in fp32), reset reproduces the run bit for bit."""
    case = make_case("elastic", 3, 160, 128, 144, 8, 0, 1, 2, 12, 0, nt=60)
    s = case.setup(Solver(case.desc))
    s.run(0, 60)
    a = s.seismogram()
    s.reset()
    s.run(0, 60)
    assert np.array_equal(a, s.seismogram())
    st, si, sg = case.src
    s.set_sources(st, si, 2.0 * sg)
    s.reset()
    s.run(0, 60)
    b = s.seismogram()
    big = np.abs(a) > 1e-20  # power-of-two scaling is exact except where intermediates underflow to subnormals
    assert np.array_equal((2.0 * a)[big], b[big])
    assert np.allclose(2.0 * a, b, rtol=0, atol=1e-30)
    assert s.is_finite()


# ---------------------------------------------------------------------------------------------------------------------
# tiled TMA kernels (3D elastic): same arithmetic sequence as the general kernels -> bit-identical in FMA mode
# ---------------------------------------------------------------------------------------------------------------------
FAST_SHAPES = [
    # nx, ny, nz, q, fs, damp, W, edge policy
    (64, 40, 32, 8, 1, 2, 8, 0), (68, 37, 21, 8, 1, 2, 6, 0), (132, 50, 40, 8, 0, 2, 10, 0), (64, 33, 16, 8, 0, 0, 6, 0),
    (96, 44, 36, 4, 1, 2, 8, 0), (72, 70, 50, 8, 1, 0, 6, 0),
    # order-reducing edges (useStencilMatrix=0, the par/ default): own weights within q/2 of every grid face
    (64, 40, 32, 8, 1, 2, 8, 1), (132, 50, 40, 8, 0, 2, 10, 1), (196, 37, 29, 8, 1, 2, 6, 1), (96, 44, 36, 4, 0, 2, 8, 1), (72, 38, 24, 4, 1, 2, 4, 1),
]


@pytest.mark.parametrize("shape", FAST_SHAPES, ids=["%dx%dx%d-q%d-fs%d-damp%d-w%d-pol%d" % s for s in FAST_SHAPES])
def test_fast_kernels_equal_general_kernels(shape):
    nx, ny, nz, q, fs, damp, W, pol = shape
    res = []
    for variant in (0, 1):
        case = make_case("elastic", 3, nx, ny, nz, q, pol, fs, damp, W, 0, nt=40, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        assert s.uses_fast_kernels() == (variant == 0)
        s.run(0, 40)
        s.sync()
        res.append((s.seismogram(), {f: s.wavefield(f) for f in fields_of("elastic", 3, 0)}))
        assert s.is_finite()
    assert np.abs(res[0][0]).max() > 0
    for f in res[0][1]:
        a, b = res[0][1][f].reshape(ny, nz, nx), res[1][1][f].reshape(ny, nz, nx)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s differs at %d points, first (y,z,x) %s, y range %d..%d z %d..%d x %d..%d, max |d| %.3e of %.3e"
                                 % (f, len(bad), bad[0], bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max(),
                                    bad[:, 2].min(), bad[:, 2].max(), np.abs(a - b).max(), np.abs(a).max()))
    assert np.array_equal(res[0][0], res[1][0])


@pytest.mark.parametrize("family", [2, 3], ids=["cpasync", "tma"])
@pytest.mark.parametrize("cfg", SWEEP, ids=[sweep_id(c) for c in SWEEP])
def test_marching_kernels_equal_per_point_kernels(cfg, family):
    """ws_kernels_march.cuh (register queues along y, cp.async-staged planes for x / z) and ws_kernels_tma.cuh (the same
    march fed by a TMA producer through an mbarrier ring) against the per-point kernels: same statement sequence, same
    accumulation order => bit-identical in FMA mode, for every equation type."""
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    res = []
    for variant in (1, family):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=40, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        assert s.kernel_path() == (0 if variant == 1 else (1 if variant == 2 else 3))
        s.run(0, 40)
        s.sync()
        res.append((s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}))
        assert s.is_finite()
        s.close()
    assert np.abs(res[0][0]).max() > 0
    assert np.array_equal(res[0][0], res[1][0])
    for f in res[0][1]:
        assert np.array_equal(res[0][1][f], res[1][1][f]), f


DEFAULT_PATHS = [
    # (cfg, expected ws_kernel_path with kernel_variant = 0)
    (("elastic", 3, 72, 64, 24, 8, 0, 1, 2, 8, 0), 2),       # 3-D elastic TMA kernels
    (("viscoelastic", 3, 72, 64, 24, 8, 0, 1, 2, 8, 2), 3),  # velocity half-step: 3-D elastic TMA kernel, stress half-step: TMA marching kernel
    (("viscoelastic", 3, 70, 64, 24, 6, 1, 1, 2, 8, 1), 3),  # both half-steps on the TMA marching kernels
    (("acoustic", 3, 100, 80, 40, 8, 0, 0, 2, 10, 0), 3),
    (("elastic", 3, 64, 48, 40, 8, 1, 1, 2, 8, 0), 2),       # order-reducing edges (the par/ default) + CPML: 3-D elastic TMA kernels
    (("elastic", 3, 64, 48, 40, 8, 1, 1, 1, 8, 0), 3),       # ... + ABS frame: TMA marching kernels
    (("viscotmem", 2, 900, 300, 1, 8, 0, 0, 2, 20, 1), 4),   # 2-D: tile kernels
    (("elastic", 2, 1000, 300, 1, 8, 0, 1, 2, 20, 0), 4),
]


@pytest.mark.parametrize("cfg,path", DEFAULT_PATHS, ids=[sweep_id(c[0]) for c in DEFAULT_PATHS])
def test_default_kernels_equal_per_point_kernels(cfg, path):
    """kernel_variant = 0 (what a par/ configuration gets) against the per-point kernels: bit-identical in FMA mode."""
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    res = []
    for variant in (1, 0):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=40, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        if variant == 0:
            assert s.kernel_path() >= 1 and (path is None or s.kernel_path() == path)
        s.run(0, 40)
        s.sync()
        res.append((s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}))
        assert s.is_finite()
        s.close()
    assert np.abs(res[0][0]).max() > 0
    assert np.array_equal(res[0][0], res[1][0])
    for f in res[0][1]:
        assert np.array_equal(res[0][1][f], res[1][1][f]), f


MARCH_SHAPES = [
    # eq, dim, nx, ny, nz, q, pol, fs, damp, W, L : ragged tiles, several tiles per axis, several y chunks
    ("acoustic", 3, 100, 150, 70, 8, 0, 0, 2, 10, 0), ("elastic", 3, 70, 90, 50, 8, 0, 1, 2, 8, 0), ("viscoelastic", 3, 67, 75, 35, 8, 0, 1, 2, 8, 2),
    ("elastic", 2, 1000, 700, 1, 8, 0, 1, 2, 20, 0), ("viscotmem", 2, 900, 300, 1, 8, 0, 0, 2, 20, 1), ("viscoemem", 3, 40, 90, 37, 4, 1, 0, 2, 6, 1),
    ("acoustic", 2, 515, 260, 1, 12, 1, 1, 1, 12, 0), ("viscosh", 2, 300, 200, 1, 6, 0, 1, 2, 10, 2),
]


@pytest.mark.parametrize("family", [2, 3], ids=["cpasync", "tma"])
@pytest.mark.parametrize("cfg", MARCH_SHAPES, ids=[sweep_id(c) + "-%dx%dx%d" % c[2:5] for c in MARCH_SHAPES])
def test_marching_kernels_ragged_shapes(cfg, family):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    res = []
    for variant in (1, family):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=25, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        s.run(0, 25)
        s.sync()
        res.append({f: s.wavefield(f) for f in fields_of(eq, dim, L)})
        s.close()
    for f in res[0]:
        assert np.array_equal(res[0][f], res[1][f]), f


# ---------------------------------------------------------------------------------------------------------------------
# 2-D tile kernels (ws_kernels_tile2d.cuh): one thread block per 128 x 8 / 16 tile, x AND y stencils from TMA-fetched halo tiles
# ---------------------------------------------------------------------------------------------------------------------
TILE_SHAPES = [
    # eq, dim, nx, ny, nz, q, pol, fs, damp, W, L : ragged tiles on both axes, several tiles per axis, every 2-D equation type
    ("elastic", 2, 1000, 700, 1, 8, 0, 1, 2, 20, 0), ("elastic", 2, 515, 260, 1, 12, 1, 1, 1, 12, 0), ("elastic", 2, 130, 45, 1, 2, 1, 0, 0, 6, 0),
    ("acoustic", 2, 515, 260, 1, 12, 1, 1, 1, 12, 0), ("acoustic", 2, 300, 129, 1, 8, 0, 1, 2, 10, 0), ("acoustic", 2, 257, 100, 1, 6, 0, 0, 2, 8, 0),
    ("viscoelastic", 2, 400, 150, 1, 8, 0, 1, 2, 12, 2), ("viscoelastic", 2, 260, 97, 1, 4, 1, 0, 1, 8, 4), ("viscoelastic", 2, 300, 120, 1, 10, 1, 1, 2, 10, 1),
    ("sh", 2, 300, 200, 1, 8, 1, 1, 2, 10, 0), ("viscosh", 2, 300, 200, 1, 6, 0, 1, 2, 10, 2), ("viscosh", 2, 140, 77, 1, 4, 1, 0, 1, 8, 3),
    ("tmem", 2, 270, 130, 1, 8, 0, 0, 2, 10, 0), ("viscotmem", 2, 900, 300, 1, 8, 0, 0, 2, 20, 1), ("viscotmem", 2, 200, 90, 1, 4, 1, 0, 1, 8, 3),
    ("emem", 2, 270, 130, 1, 8, 0, 0, 2, 10, 0), ("viscoemem", 2, 333, 111, 1, 6, 1, 0, 2, 9, 2), ("viscoemem", 2, 200, 64, 1, 12, 0, 0, 0, 6, 4),
]


@pytest.mark.parametrize("ty", [8, 16])
@pytest.mark.parametrize("cfg", TILE_SHAPES, ids=[sweep_id(c) + "-%dx%d" % c[2:4] for c in TILE_SHAPES])
def test_tile2d_kernels_equal_per_point_kernels(cfg, ty, monkeypatch):
    """Same statement sequence, same accumulation order as the per-point kernels => bit-identical in FMA mode, for both tile
    heights (WS_TILE_TY is a developer switch read at ws_prepare)."""
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    monkeypatch.setenv("WS_TILE_TY", str(ty))
    res = []
    for variant in (1, 4):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=30, exact=0, kernel_variant=variant)
        s = case.setup(Solver(case.desc))
        assert s.kernel_path() == (0 if variant == 1 else 4)
        s.run(0, 30)
        s.sync()
        res.append((s.seismogram(), {f: s.wavefield(f) for f in fields_of(eq, dim, L)}))
        assert s.is_finite()
        s.close()
    assert np.abs(res[0][0]).max() > 0
    assert np.array_equal(res[0][0], res[1][0])
    for f in res[0][1]:
        assert np.array_equal(res[0][1][f], res[1][1][f]), f


@pytest.mark.parametrize("cfg", [c for c in SWEEP if c[1] == 2], ids=[sweep_id(c) for c in SWEEP if c[1] == 2])
def test_tile2d_kernels_vs_oracle(cfg):
    """The 2-D default path (kernel_variant 0 -> tile kernels) against the CPU oracle on the sweep cases."""
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=40, exact=0, kernel_variant=0)
    o, s = run_pair(case, 40)
    assert s.kernel_path() == 4
    assert rel_l2(s.seismogram(), o.seismogram()) <= TOL


def test_fast_kernels_vs_oracle():
    case = make_case("elastic", 3, 64, 48, 40, 8, 0, 1, 2, 8, 0, nt=60, exact=0, kernel_variant=0)
    o, s = run_pair(case, 60)
    assert s.uses_fast_kernels()
    assert rel_l2(s.seismogram(), o.seismogram()) <= TOL


def test_fast_kernels_3d_elastic_golden():
    """3D elastic CI case has FD order 2 (general kernels); run it at FD order 8 on the tiled kernels against the oracle."""
    case = ci_case("3D.elastic", nt=250)
    case.desc.fd_order = 8
    case.desc.damping = 2
    o, s = run_pair(case, 250)
    assert s.uses_fast_kernels()
    assert rel_l2(s.seismogram(), o.seismogram()) <= TOL


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at (or near) their full sizes: size-independent properties (the oracle would not finish)
# ---------------------------------------------------------------------------------------------------------------------
def _layered_2d(nx, ny):
    k = np.minimum(np.arange(ny) * 8 // ny, 7).astype(np.float32)[:, None]
    vp = np.broadcast_to(1500.0 + 250.0 * k, (ny, nx)).astype(np.float32)
    return dict(velocityP=vp.ravel(), velocityS=(vp / np.float32(np.sqrt(3.0))).ravel().astype(np.float32),
                density=np.broadcast_to(1800.0 + 100.0 * k, (ny, nx)).astype(np.float32).ravel())


BIG = [
    # name, eq, dim, nx, ny, nz, fs, L, dh, dt, fc, vmax, relax
    ("cfg2-2D-elastic-4096", "elastic", 2, 4096, 4096, 1, 1, 0, 5.0, 5e-4, 10.0, 3500.0, ()),
    ("cfg3-3D-acoustic-512", "acoustic", 3, 512, 512, 512, 0, 0, 10.0, 1e-3, 10.0, 3500.0, ()),
    ("cfg4-3D-visco-L2-256", "viscoelastic", 3, 256, 256, 256, 1, 2, 10.0, 8e-4, 10.0, 4550.0, (5.0, 50.0)),
    ("cfg5-2D-viscotmem-8192x2048", "viscotmem", 2, 8192, 2048, 1, 0, 1, 0.01, 1.5e-11, 1e8, 3e8, (1e8,)),
]


@pytest.mark.parametrize("cfg", BIG, ids=[c[0] for c in BIG])
def test_baseline_configs_linearity_reset_finite(cfg):
    """FD order 8 + CPML(20) on the BASELINE.json grids: doubling the source doubles every trace exactly, a reset
    reproduces the run bit for bit, the wavefields stay finite and the direct wave arrives (non-zero traces)."""
    from wsharness import make_desc, idx1d, ricker_np
    from cases import EPS0, MU0
    name, eq, dim, nx, ny, nz, fs, L, dh, dt, fc, vmax, relax = cfg
    nt = 48
    d = make_desc(dim, eq, nx, ny, nz, dh=dh, dt=dt, nt=nt, fd_order=8, edge_policy=0, free_surface=fs, damping=2, boundary_width=20,
                  vmax_cpml=vmax, fc_cpml=fc, npower=4.0, relax_freq=relax)
    s = Solver(d)
    n = nx * ny * nz
    if eq == "viscotmem":
        s.set_material("dielectricPermittivity", np.full(n, 4 * EPS0, np.float32))
        s.set_material("electricConductivity", np.full(n, 1e-3, np.float32))
        s.set_material("magneticPermeability", np.full(n, MU0, np.float32))
        s.set_material("tauDielectricPermittivity", np.full(n, 0.05, np.float32))
        s.set_material("tauElectricConductivity", np.zeros(n, np.float32))
        stype, rtype, amp = 1, 1, 1.0
    else:
        if dim == 2:
            m = _layered_2d(nx, ny)
        else:
            y = (np.arange(ny, dtype=np.float32) / ny)[:, None, None]
            vp = np.broadcast_to(2000.0 + 1500.0 * y, (ny, nz, nx)).astype(np.float32).ravel()
            m = dict(velocityP=vp, velocityS=(vp / np.float32(np.sqrt(3.0))).astype(np.float32), density=np.full(n, 2000.0, np.float32))
        for k in (["velocityP", "density"] if eq == "acoustic" else ["velocityP", "velocityS", "density"]):
            s.set_material(k, m[k])
        if eq == "viscoelastic":
            s.set_material("tauP", np.full(n, 0.1, np.float32))
            s.set_material("tauS", np.full(n, 0.1, np.float32))
        stype, rtype, amp = (1, 1, 1e6) if eq == "acoustic" else (3, 3, 1e6)
    s.prepare()
    ys = 24
    sig = ricker_np(nt, dt, fc * 4, amp)[None, :]  # short wavelet: the direct wave reaches the receivers within nt steps
    s.set_sources([stype], [idx1d(nx // 2, ys, nz // 2, nx, nz)], sig)
    rec = [idx1d(nx // 2 + 2 + 2 * i, ys, nz // 2, nx, nz) for i in range(8)]
    s.set_receivers([rtype] * 8, rec)
    s.reset()
    s.run(0, nt)
    a = s.seismogram()
    assert s.is_finite()
    assert np.abs(a).max() > 0
    s.reset()
    s.run(0, nt)
    assert np.array_equal(a, s.seismogram())
    s.set_sources([stype], [idx1d(nx // 2, ys, nz // 2, nx, nz)], 2.0 * sig)
    s.reset()
    s.run(0, nt)
    b = s.seismogram()
    big = np.abs(a) > 1e-18 * np.abs(a).max()
    assert np.array_equal((2.0 * a)[big], b[big])
    s.close()

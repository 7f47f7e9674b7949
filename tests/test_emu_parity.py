"""CPU-only logic check of the PRODUCT sources: wave-simulation_b200/csrc (C ABI, model preparation, general kernels,
tables) compiled with -DWS_EMULATE against tests/emu/cuda_emu.hpp and compared with the oracle.  In exact-arithmetic
mode the product's statement order equals the reference's, so results must be BIT-IDENTICAL; in the default FMA mode they
must agree to well below the 1e-5 relative-L2 bar of BASELINE.json.  (GPU execution itself is covered by `-m gpu`.)"""
import numpy as np
import pytest

from cases import SWEEP, fields_of, make_case, sweep_id
from wsharness import EmuSolver, Oracle, ci_case, rel_l2


@pytest.mark.parametrize("cfg", SWEEP, ids=[sweep_id(c) for c in SWEEP])
def test_emulated_kernels_bit_exact(cfg):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=30, exact=1)
    o = case.setup(Oracle(case.desc))
    e = case.setup(EmuSolver(case.desc))
    o.run(0, 30)
    e.run(0, 30)
    so, se = o.seismogram(), e.seismogram()
    assert np.abs(so).max() > 0
    assert np.array_equal(so, se)
    for f in fields_of(eq, dim, L):
        assert np.array_equal(o.wavefield(f), e.wavefield(f)), f


@pytest.mark.parametrize("cfg", SWEEP[::3], ids=[sweep_id(c) for c in SWEEP[::3]])
def test_emulated_kernels_fma_mode(cfg):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=30, exact=0)
    o = case.setup(Oracle(case.desc))
    e = case.setup(EmuSolver(case.desc))
    o.run(0, 30)
    e.run(0, 30)
    assert rel_l2(e.seismogram(), o.seismogram()) <= 1.0e-5


# every equation type once (the GPU suite runs the whole sweep)
MARCH_EMU = [c for k, c in enumerate(SWEEP) if k in (0, 2, 3, 6, 8, 9, 11, 12, 14, 15, 17, 18, 19, 20, 22, 23)]


@pytest.mark.parametrize("cfg", MARCH_EMU, ids=[sweep_id(c) for c in MARCH_EMU])
def test_emulated_marching_kernels_equal_per_point_kernels(cfg):
    """The marching kernels (register queues + staged planes, ws_kernels_march.cuh) run the statement sequence of the
    per-point kernels with the weights applied in the same order: bit-identical wavefields in FMA mode."""
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    res = []
    for variant in (1, 2):
        case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=12, exact=0, kernel_variant=variant)
        e = case.setup(EmuSolver(case.desc))
        assert e.kernel_path() == (0 if variant == 1 else 1)
        e.run(0, 12)
        res.append((e.seismogram(), {f: e.wavefield(f) for f in fields_of(eq, dim, L)}))
        e.close()
    assert np.abs(res[0][0]).max() > 0
    assert np.array_equal(res[0][0], res[1][0])
    for f in res[0][1]:
        assert np.array_equal(res[0][1][f], res[1][1][f]), f


def test_emulated_marching_kernels_partial_tiles_and_chunks(monkeypatch):
    """grid sizes that are not multiples of the tile, several y chunks per column (WS_MARCH_CHUNK)"""
    monkeypatch.setenv("WS_MARCH_CHUNK", "7")
    # the last three contain interior tiles (the thread-block-uniform fast path without edge rows / CPML / ABS)
    for cfg in (("elastic", 3, 37, 23, 11, 8, 0, 1, 2, 5, 0), ("viscoelastic", 2, 150, 31, 1, 8, 0, 1, 2, 6, 2), ("acoustic", 3, 34, 20, 19, 8, 0, 0, 2, 5, 0),
                ("acoustic", 3, 100, 24, 30, 4, 0, 0, 2, 5, 0), ("viscoelastic", 2, 300, 40, 1, 8, 0, 1, 2, 6, 2), ("viscoelastic", 3, 72, 20, 26, 4, 1, 1, 1, 4, 1)):
        eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
        res = []
        for variant in (1, 2):
            case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=8, exact=0, kernel_variant=variant)
            e = case.setup(EmuSolver(case.desc))
            e.run(0, 8)
            res.append({f: e.wavefield(f) for f in fields_of(eq, dim, L)})
            e.close()
        for f in res[0]:
            assert np.array_equal(res[0][f], res[1][f]), (cfg, f)


@pytest.mark.parametrize("dim", [2, 3])
def test_div_curl_snapshot(dim):
    """snapType 3 (Wavefields3Delastic.cpp:197-245, Wavefields2Delastic.cpp:217-248): for velocity fields that are linear
    in the coordinates the interior values are known in closed form (FD weights sum to DT/DH per unit slope)."""
    nx, ny, nz = 24, 22, (20 if dim == 3 else 1)
    case = make_case("elastic", dim, nx, ny, nz, 4, 0, 0, 0, 6, 0, nt=4, exact=1)
    e = case.setup(EmuSolver(case.desc))
    y, z, x = np.meshgrid(np.arange(ny), np.arange(nz), np.arange(nx), indexing="ij")
    a, b, c = 3.0, -2.0, 0.5
    e.set_wavefield("VX", (a * y + 2 * x).astype(np.float32))
    e.set_wavefield("VY", (b * x + c * z + 1.5 * y).astype(np.float32))
    if dim == 3:
        e.set_wavefield("VZ", (0.25 * y - 1.0 * z).astype(np.float32))
    s = np.float32(case.desc.dt / case.desc.dh)
    pi, mu = e.get_material("pWaveModulus").reshape(ny, nz, nx), e.get_material("sWaveModulus").reshape(ny, nz, nx)
    div, curl = e.wavefield("DIV").reshape(ny, nz, nx), e.wavefield("CURL").reshape(ny, nz, nx)
    inner = (slice(3, ny - 3), slice(3, nz - 3) if dim == 3 else slice(None), slice(3, nx - 3))
    if dim == 3:
        d = (2 + 1.5 - 1.0) * s  # Dxb vx + Dyb vy + Dzb vz
        want_div = np.sqrt(d * d * pi)
        cx, cy, cz = (0.25 - c) * s, (0.0 - 0.0) * s, (b - a) * s  # (Dyf vz - Dzf vy), (Dzf vx - Dxf vz), (Dxf vy - Dyf vx)
        want_curl = np.sqrt((cx * cx + cy * cy + cz * cz) * mu)
    else:
        want_div = (2 + 1.5) * s * np.sqrt(pi)
        want_curl = (a - b) * s * np.sqrt(mu)  # Dyf vx - Dxf vy
    assert np.allclose(div[inner], want_div[inner], rtol=2e-5)
    assert np.allclose(curl[inner], want_curl[inner], rtol=2e-5)
    with pytest.raises(RuntimeError):
        a_case = make_case("acoustic", 2, 20, 20, 1, 2, 0, 0, 0, 6, 0, nt=2)
        a_case.setup(EmuSolver(a_case.desc)).wavefield("DIV")


@pytest.mark.parametrize("eq,dim", [("tmem", 2), ("viscoemem", 3)])
def test_div_curl_snapshot_em(eq, dim):
    """EM variants (WavefieldsEM/Wavefields2Dtmem.cpp, Wavefields3Dviscoemem.cpp): magnetic field, curl scaled by the
    permittivity, div by the EM velocity 1/sqrt(eps mu) (3-D visco: by the conductivity)."""
    nx, ny, nz = 24, 22, (20 if dim == 3 else 1)
    case = make_case(eq, dim, nx, ny, nz, 4, 0, 0, 0, 6, 1 if eq.startswith("visco") else 0, nt=4, exact=1)
    e = case.setup(EmuSolver(case.desc))
    y, z, x = np.meshgrid(np.arange(ny), np.arange(nz), np.arange(nx), indexing="ij")
    a, b, c = 3.0, -2.0, 0.5
    e.set_wavefield("HX", (a * y + 2 * x).astype(np.float32))
    e.set_wavefield("HY", (b * x + c * z + 1.5 * y).astype(np.float32))
    if dim == 3:
        e.set_wavefield("HZ", (0.25 * y - 1.0 * z).astype(np.float32))
    s = np.float32(case.desc.dt / case.desc.dh)
    eps = case.materials["dielectricPermittivity"].reshape(ny, nz, nx).astype(np.float64)
    mu = case.materials["magneticPermeability"].reshape(ny, nz, nx).astype(np.float64)
    sig = case.materials["electricConductivity"].reshape(ny, nz, nx).astype(np.float64)
    div, curl = e.wavefield("DIV").reshape(ny, nz, nx), e.wavefield("CURL").reshape(ny, nz, nx)
    inner = (slice(3, ny - 3), slice(3, nz - 3) if dim == 3 else slice(None), slice(3, nx - 3))
    if dim == 3:
        d = (2 + 1.5 - 1.0) * s
        want_div = np.sqrt(d * d * sig)  # Wavefields3Dviscoemem.cpp:99: the conductivity
        cx, cz = (0.25 - c) * s, (b - a) * s
        want_curl = np.sqrt((cx * cx + cz * cz) * eps)
    else:
        want_div = (2 + 1.5) * s * np.sqrt(1.0 / np.sqrt(eps * mu))
        want_curl = (a - b) * s * np.sqrt(eps)
    assert np.allclose(div[inner], want_div[inner], rtol=3e-5)
    assert np.allclose(curl[inner], want_curl[inner], rtol=3e-5)


def test_emulated_ci_case_2d_elastic_full_trace():
    case = ci_case("2D.elastic")
    for exact, tol in ((1, 0.0), (0, 1.0e-5)):
        case.desc.exact_arith = exact
        case.desc.kernel_variant = 1  # per-point kernels: 1000 steps of the thread-per-CUDA-thread emulation of the marching kernels take minutes
        o = case.setup(Oracle(case.desc))
        e = case.setup(EmuSolver(case.desc))
        o.run(0, 1000)
        e.run(0, 1000)
        assert rel_l2(e.seismogram(), o.seismogram()) <= tol


def test_derived_model_parameters_match():
    case = make_case("viscoelastic", 3, 14, 15, 13, 4, 1, 1, 2, 4, 2, nt=4)
    o = case.setup(Oracle(case.desc))
    e = case.setup(EmuSolver(case.desc))
    for name in ("pWaveModulus", "sWaveModulus", "inverseDensityAverageX", "inverseDensityAverageY", "inverseDensityAverageZ",
                 "sWaveModulusAverageXY", "sWaveModulusAverageXZ", "sWaveModulusAverageYZ", "tauSAverageXY", "tauSAverageXZ",
                 "tauSAverageYZ"):
        assert np.array_equal(o.get_material(name), e.get_material(name)), name


def test_step_host_and_reset():
    case = make_case("elastic", 2, 30, 28, 1, 4, 0, 1, 1, 6, 0, nt=12)
    e = case.setup(EmuSolver(case.desc))
    e.run(0, 12)
    ref = e.seismogram()
    e.reset()
    sig = case.src[2]
    rec = np.zeros(4, np.float32)
    got = np.zeros_like(ref)
    for t in range(12):
        e.step_host(t, np.ascontiguousarray(sig[:, t]), rec)
        got[:, t] = rec
    assert np.array_equal(got, ref)
    assert np.array_equal(e.seismogram(), ref)
    assert e.is_finite()


def test_error_behaviour():
    from wsharness import make_desc
    with pytest.raises(RuntimeError, match="Unsupported spatialFDorder"):
        EmuSolver(make_desc(2, "acoustic", 30, 30, fd_order=7))
    with pytest.raises(RuntimeError, match="2D only"):
        EmuSolver(make_desc(3, "sh", 30, 30, 30))
    s = EmuSolver(make_desc(2, "sh", 30, 30, nt=5))
    with pytest.raises(RuntimeError, match="SH modeling"):
        s.set_sources([1], [5], np.zeros((1, 5), np.float32))
    with pytest.raises(RuntimeError, match="not set"):
        s.prepare()
    with pytest.raises(RuntimeError, match="ws_prepare must be called"):
        s.step(0)


def test_abi_edge_cases_empty_acquisition_and_error_codes():
    """Edge cases of the C ABI on the emulation build (the host code of the entry points is the product's): no sources, no receivers,
    an empty time range, indices outside the grid, stepping before ws_prepare, a time range beyond NT, a bad descriptor."""
    from wsharness import make_desc
    case = make_case("elastic", 2, 40, 36, 1, 4, 0, 1, 2, 8, 0, nt=12, exact=1)
    o = case.setup(Oracle(case.desc))
    e = case.setup(EmuSolver(case.desc))
    # an empty time range is a no-op
    e.run(5, 5)
    assert not e.wavefield("VX").any()
    # no sources: nothing moves; the receivers record zeros
    types, idx = case.rec
    e.set_sources([], [], np.zeros((0, 12), np.float32))
    e.reset()
    e.run(0, 12)
    assert not e.seismogram().any() and not e.wavefield("Sxx").any() and e.is_finite()
    # no receivers: an empty seismogram, the wavefields as with receivers
    st, sidx, sig = case.src
    e.set_sources(st, sidx, sig)
    e.set_receivers([], [])
    e.reset()
    e.run(0, 12)
    o.run(0, 12)
    assert e.seismogram().shape == (0, 12)
    assert np.array_equal(e.wavefield("VY"), o.wavefield("VY"))
    # indices outside the grid are refused with the reference's kind of message, the acquisition in place stays
    with pytest.raises(RuntimeError, match="(?i)index|coordinate|range|grid"):
        e.set_receivers([1], [40 * 36])
    with pytest.raises(RuntimeError, match="(?i)index|coordinate|range|grid"):
        e.set_sources([1], [-1], np.zeros((1, 12), np.float32))
    with pytest.raises(RuntimeError, match="(?i)type"):
        e.set_receivers([7], [0])
    # time range beyond NT
    with pytest.raises(RuntimeError, match="time range"):
        e.run(0, 13)
    e.close()
    o.close()
    # stepping before ws_prepare
    fresh = EmuSolver(case.desc)
    with pytest.raises(RuntimeError, match="ws_prepare"):
        fresh.run(0, 1)
    fresh.close()
    # descriptors the reference rejects too
    for bad in (dict(fd_order=7), dict(damping=3), dict(free_surface=3)):
        kw = dict(dh=10.0, dt=8e-4, nt=4, fd_order=4, edge_policy=0, free_surface=0, damping=0, boundary_width=0)
        kw.update(bad)
        with pytest.raises(RuntimeError):
            EmuSolver(make_desc(2, "elastic", 20, 20, 1, **kw))
    with pytest.raises(RuntimeError, match="2D only"):
        EmuSolver(make_desc(3, "sh", 20, 20, 20, dh=10.0, dt=8e-4, nt=4, fd_order=4, edge_policy=0, free_surface=0, damping=0, boundary_width=0))

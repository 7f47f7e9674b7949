"""The LAMA-free C++ host layer (wave-simulation_b200/host): the reference's Configuration / Coordinates / Acquisition /
Modelparameter / Wavefields / Derivatives / ForwardSolver classes and the `Simulation` driver on top of the C ABI.

CPU part: unit tests with the reference's known answers (tests/host/test_host.cpp) and the driver run end to end on
par/ci-style inputs against the golden seismograms — linked, FOR THESE TESTS ONLY, against the host emulation build of
the library (tests/emu).  GPU part (`-m gpu`): the product binary `host/Simulation` on the CUDA library."""
import os
import struct
import subprocess

import numpy as np
import pytest

from wsharness import ROOT, EmuSolver, Oracle, build_emu, ci_case, golden, rel_l2, two_layer

HOST = os.path.join(ROOT, "wave-simulation_b200", "host")
EMU_DIR = os.path.join(ROOT, "tests", "emu")

CONFIG_2D_ELASTIC = """# 2D elastic CI case of the reference (par/ci/configuration_ci.2D.elastic.txt), transcribed
dimension=2D
equationType=elastic
NX=100 # horizontal 1
NY=100 # depth
NZ=1
UseVariableGrid=0
useVariableFDoperators=0
useStencilMatrix={stencil}
NumShotDomains=1
DH=50
DT=2.0e-03                   # temporal sampling in seconds
T={T}
spatialFDorder=12
ModelRead=1
ModelFilename=model/model
fileFormat={fmt}
numRelaxationMechanisms=0;
FreeSurface=1
DampingBoundary=1
BoundaryWidth=9
DampingCoeff=8.0
VMaxCPML=3500
CenterFrequencyCPML=5
NPower=4
SourceFilename=acq/sources
ReceiverFilename=acq/receiver
SeismogramFilename=seismograms/seismogram
initSourcesFromSU=0
initReceiverFromSU=0
SeismogramFormat={sfmt}
normalizeTraces={norm}
useReceiversPerShot={rps}
writeSource=0
seismoDT={sdt}
snapType={snap}
WavefieldFileName=wavefields/wavefield
tFirstSnapshot=0
tLastSnapshot=2
tIncSnapshot=0.1
verbose=0
kernelVariant={kvar}
"""


def write_mtx_vector(path, v):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n%d 1\n" % v.size)
        f.write("\n".join("%.9g" % x for x in v))
        f.write("\n")


def write_lmf_vector(path, v):
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", 0x4711E01, 0, 2, 1, v.size))
        f.write(np.asarray(v, "<f4").tobytes())


def read_mtx(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("%")]
    r, c = (int(x) for x in lines[0].split())
    return np.array([float(x) for x in lines[1:1 + r * c]]).reshape(c, r).T


def read_lmf_matrix(path):
    raw = open(path, "rb").read()
    ident, itype, vtype, ndims, r, c = struct.unpack("<6i", raw[:24])
    assert (ident, itype, vtype, ndims) == (0x4711E01, 0, 2, 2)
    return np.frombuffer(raw[24:], "<f4").reshape(r, c)


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    build_emu()
    subprocess.check_call(["make", "-s", "-C", HOST, "libSimulation_host.a", "Simulation.o"])
    subprocess.check_call(["make", "-s", "-C", HOST, "emu"])
    if os.environ.get("WS_HOST_SANITIZE"):
        # WS_HOST_SANITIZE=1 pytest tests/test_host_layer.py -m "not gpu": the same driver tests on a build of the host layer with the address,
        # leak and undefined-behaviour sanitizers (any report ends the run with a non-zero exit code, which the tests check)
        out = str(tmp_path_factory.mktemp("sanitize"))
        srcs = [f for f in sorted(os.listdir(HOST)) if f.endswith(".cpp")]
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-pthread",
                               "-I" + HOST] + [os.path.join(HOST, f) for f in srcs] + ["-L" + EMU_DIR, "-lwavesim_emu", "-Wl,-rpath," + EMU_DIR, "-fopenmp", "-o",
                                                                                      os.path.join(out, "Simulation_sanitize")])
        return os.path.join(out, "Simulation_sanitize")
    return os.path.join(EMU_DIR, "Simulation_emu")


def setup_case(tmp, fmt=1, sources="1 20  0   0   2   1   1   5.0   5.0   0.0\n", receivers="30 0 0 3\n", **kw):
    for d in ("model", "acq", "seismograms", "wavefields"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    m = two_layer(100, 100, 1)
    w = write_mtx_vector if fmt == 1 else write_lmf_vector
    ext = ".mtx" if fmt == 1 else ".lmf"
    for key, suffix in (("velocityP", "vp"), ("velocityS", "vs"), ("density", "density")):
        w(os.path.join(tmp, "model", "model." + suffix + ext), m[key])
    open(os.path.join(tmp, "acq", "sources.txt"), "w").write("# sourceNo X Y Z type wType wShape fc amp tShift\n" + sources)
    open(os.path.join(tmp, "acq", "receiver.txt"), "w").write("# X Y Z type\n" + receivers)
    # kernelVariant=1: per-point kernels (the emulation of the marching kernels runs one OS thread per CUDA thread: slow
    # over 1000 steps); the SU / snapshot test keeps the default (marching kernels) over 150 steps
    par = dict(stencil=1, T=2, fmt=fmt, sfmt=1, norm=0, rps=0, sdt="2.0e-03", snap=0, kvar=1)
    par.update(kw)
    cfg = os.path.join(tmp, "configuration.txt")
    open(cfg, "w").write(CONFIG_2D_ELASTIC.format(**par))
    return cfg


def run(driver, cfg, cwd, expect_ok=True):
    p = subprocess.run([driver, cfg], cwd=cwd, capture_output=True, text=True, timeout=900)
    if expect_ok:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p


def test_host_unit_tests(tmp_path):
    build_emu()  # the model classes reference the solver library (device-side averaging); nothing of it runs in these tests
    subprocess.check_call(["make", "-s", "-C", HOST, "libSimulation_host.a"])
    exe = str(tmp_path / "test_host")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I" + HOST, os.path.join(ROOT, "tests", "host", "test_host.cpp"),
                           os.path.join(HOST, "libSimulation_host.a"), "-L" + EMU_DIR, "-lwavesim_emu", "-Wl,-rpath," + EMU_DIR, "-fopenmp", "-o", exe])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "host unit tests OK" in out.stdout


def test_driver_reproduces_reference_golden_2d_elastic(driver, tmp_path):
    """par/ci 2D elastic case end to end through the driver: config file, source / receiver files, .mtx model files,
    .mtx seismogram; golden = the reference's own par/ci/seismogram.2D.elastic.ref.vy.mtx."""
    cfg = setup_case(str(tmp_path))
    run(driver, cfg, str(tmp_path))
    s = read_mtx(str(tmp_path / "seismograms" / "seismogram.shot_1.vy.mtx"))
    g = golden("seismogram.2D.elastic.ref.vy.mtx")
    assert s.shape == g.shape == (1, 1000)
    assert rel_l2(s, g) <= 1.0e-5


def test_driver_sparse_policy_lmf_and_oracle(driver, tmp_path):
    """useStencilMatrix=0 selects the order-reducing operators (Derivatives.cpp:129-186); .lmf model and seismogram files;
    compared with the oracle on the same case."""
    cfg = setup_case(str(tmp_path), fmt=2, stencil=0, sfmt=2, T=0.8)
    run(driver, cfg, str(tmp_path))
    s = read_lmf_matrix(str(tmp_path / "seismograms" / "seismogram.shot_1.vy.lmf"))
    case = ci_case("2D.elastic", nt=400)
    o = case.setup(Oracle(case.desc))
    o.run(0, 400)
    assert s.shape == (1, 400)
    assert rel_l2(s, o.seismogram()) <= 1.0e-5


def test_driver_free_surface_vacuum_formulation(driver, tmp_path):
    """FreeSurface = 2 (improved vacuum formulation): plain derivative operators and, as with the image method, no absorbing frame at the
    top (ABS2D.cpp:158 tests useFreeSurface == 0); compared with the oracle run with the same setting and told apart from FreeSurface = 1."""
    tmp = str(tmp_path)
    cfg = setup_case(tmp, T=0.8)
    text = open(cfg).read()
    open(cfg, "w").write(text.replace("FreeSurface=1", "FreeSurface=2"))
    run(driver, cfg, tmp)
    s = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    case = ci_case("2D.elastic", nt=400)
    case.desc.edge_policy = 0
    image = case.setup(Oracle(case.desc))
    image.run(0, 400)
    case.desc.free_surface = 2
    o = case.setup(Oracle(case.desc))
    o.run(0, 400)
    assert rel_l2(s, o.seismogram()) <= 1.0e-5
    assert rel_l2(s, image.seismogram()) > 1.0e-2
    open(cfg, "w").write(text.replace("FreeSurface=1", "FreeSurface=3"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "FreeSurface must be" in p.stderr


def test_driver_multishot_receivers_per_shot_resampling_normalisation_snapshots(driver, tmp_path):
    tmp = str(tmp_path)
    src = "1 20 0 0 2 1 1 5.0 5.0 0.0\n2 40 5 0 1 1 1 8.0 1.0 0.0\n-2 60 5 0 3 1 4 6.0 2.0 0.05\n"
    cfg = setup_case(tmp, sources=src, rps=1, sdt="4.0e-03", norm=1, T=0.4, snap=1)
    open(os.path.join(tmp, "acq", "receiver.shot_1.txt"), "w").write("30 0 0 3\n50 2 0 2\n")
    open(os.path.join(tmp, "acq", "receiver.shot_2.txt"), "w").write("35 3 0 1\n36 3 0 3\n37 3 0 3\n")
    run(driver, cfg, tmp)
    files = sorted(os.listdir(os.path.join(tmp, "seismograms")))
    assert files == ["seismogram.shot_1.vx.mtx", "seismogram.shot_1.vy.mtx", "seismogram.shot_2.p.mtx", "seismogram.shot_2.vy.mtx"]
    vy2 = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_2.vy.mtx"))
    assert vy2.shape == (2, 100)  # NT = 200 resampled by 2
    assert np.abs(vy2).max() <= 1.0  # normalizeTraces = 1 acts on the full-rate trace, resampling on the output
    # same shot through the Python binding of the same (emulated) library: two sources fire together in shot 2
    case = ci_case("2D.elastic", nt=200)
    case.desc.edge_policy = 0
    from wsharness import idx1d, ricker
    fg = np.zeros(200, np.float32)
    t = np.arange(200, dtype=np.float32) * np.float32(2e-3)
    tau = (t - np.float32(1.2 / 6.0 + 0.05)) * np.float32(np.pi * 6.0)
    fg = (np.float32(2.0) * (np.float32(-2.0) * tau) * np.exp(-tau * tau)).astype(np.float32)
    case.src = ([1, 3], [idx1d(40, 5, 0, 100, 1), idx1d(60, 5, 0, 100, 1)], np.stack([ricker(200, 2e-3, 8.0, 1.0, 0.0), fg]))
    case.rec = ([1, 3, 3], [idx1d(35, 3, 0, 100, 1), idx1d(36, 3, 0, 100, 1), idx1d(37, 3, 0, 100, 1)])
    e = case.setup(EmuSolver(case.desc))
    e.run(0, 200)
    ref = e.seismogram()[1:]
    ref = (ref / np.abs(ref).max(axis=1, keepdims=True))[:, ::2]
    assert rel_l2(vy2, ref) <= 1.0e-4  # plumbing check: the FGaussian wavelet above is a numpy re-evaluation (libm vs numpy exp)
    snaps = [f for f in os.listdir(os.path.join(tmp, "wavefields")) if f.startswith("wavefield.shot_1.VX.")]
    assert "wavefield.shot_1.VX.0.mtx" in snaps and "wavefield.shot_1.VX.150.mtx" in snaps and len(snaps) == 4


def read_su(path):
    """Seismic Unix traces: 240-byte SEG-Y trace header (native endian) + ns float32 samples per trace."""
    raw = open(path, "rb").read()
    ns = struct.unpack_from("<H", raw, 114)[0]
    rec = 240 + 4 * ns
    assert len(raw) % rec == 0
    ntr = len(raw) // rec
    hdr = [raw[k * rec:k * rec + 240] for k in range(ntr)]
    data = np.stack([np.frombuffer(raw, "<f4", ns, k * rec + 240) for k in range(ntr)])
    return hdr, data


def test_driver_su_seismograms(driver, tmp_path):
    """SeismogramFormat=4: trace headers as the reference fills them (src/IO/SUIO.hpp:196-246: coordinates in mm with
    scalco = -3, offset from the source, dt in microseconds), samples identical to the .mtx output of the same run."""
    tmp = str(tmp_path)
    rec = "30 0 0 3\n44 7 0 3\n"
    cfg = setup_case(tmp, receivers=rec, sfmt=4, T=0.3, sdt="4.0e-03", kvar=0)
    run(driver, cfg, tmp)
    hdr, su = read_su(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.su"))
    cfg = setup_case(tmp, receivers=rec, sfmt=1, T=0.3, sdt="4.0e-03", snap=3, kvar=0)  # + snapType 3: curl / div energy snapshots
    run(driver, cfg, tmp)
    mtx = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    snaps = sorted(os.listdir(os.path.join(tmp, "wavefields")))
    assert snaps == ["wavefield.shot_1.%s.%d.mtx" % (c, t) for c in ("CURL", "DIV") for t in (0, 100, 50)]
    div = read_mtx(os.path.join(tmp, "wavefields", "wavefield.shot_1.DIV.100.mtx"))
    assert div.shape == (10000, 1) and np.isfinite(div).all() and np.abs(div).max() > 0  # signed in 2-D (Wavefields2Delastic.cpp:234-248)
    assert su.shape == (2, 75) and np.abs(su).max() > 0
    assert np.allclose(su, mtx, rtol=2e-6, atol=0)  # the mtx file carries 6-7 digits

    def word(h, off, fmt):
        return struct.unpack_from("<" + fmt, h, off)[0]
    for k, (xr, yr) in enumerate(((30, 0), (44, 7))):
        h = hdr[k]
        assert word(h, 0, "i") == k + 1 and word(h, 204, "i") == 2 and word(h, 28, "h") == 1       # tracl, ntr, trid
        assert word(h, 114, "H") == 75 and word(h, 116, "H") == 4000                              # ns, dt [us]
        assert abs(word(h, 180, "f") - 4.0e-3) < 1e-9                                            # d1
        assert word(h, 70, "h") == -3 and word(h, 68, "h") == -3 and word(h, 88, "h") == 1       # scalco, scalel, counit
        assert word(h, 72, "i") == 20 * 50 * 1000 and word(h, 48, "i") == 0                      # sx, sdepth (source at (20, 0))
        assert word(h, 80, "i") == xr * 50 * 1000 and word(h, 40, "i") == yr * 50 * 1000          # gx, gelev
        assert word(h, 36, "i") == int(round(np.hypot((xr - 20) * 50.0, yr * 50.0) * 1000.0))     # offset
    # initReceiverFromSU=1: the receiver geometry comes from the trace headers of <ReceiverFilename>.<comp>.su (suHandler.cpp)
    os.makedirs(os.path.join(tmp, "acq_su"), exist_ok=True)
    os.rename(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.su"), os.path.join(tmp, "acq_su", "rec.vy.su"))
    open(os.path.join(tmp, "acq", "receiver.txt"), "w").write("# unused\n1 1 0 1\n")
    text = open(cfg).read().replace("initReceiverFromSU=0", "initReceiverFromSU=1").replace("ReceiverFilename=acq/receiver", "ReceiverFilename=acq_su/rec")
    open(cfg, "w").write(text.replace("snapType=3", "snapType=0"))
    run(driver, cfg, tmp)
    again = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    assert np.array_equal(again, mtx)
    # ... and with receivers per shot from <ReceiverFilename>.shot_<n>.<comp>.su (Receivers.cpp:229-246)
    os.rename(os.path.join(tmp, "acq_su", "rec.vy.su"), os.path.join(tmp, "acq_su", "rec.shot_1.vy.su"))
    os.remove(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    text = open(cfg).read().replace("useReceiversPerShot=0", "useReceiversPerShot=1")
    open(cfg, "w").write(text)
    run(driver, cfg, tmp)
    assert np.array_equal(read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx")), mtx)


def test_two_layer_tool_feeds_the_driver(driver, tmp_path):
    """host/Tools/TwoLayer (src/Tools/CreateModel/TwoLayer.cpp): the model it writes equals the fixture the other tests
    build in numpy, in both file formats, and the tool refuses to run without a configuration."""
    subprocess.check_call(["make", "-s", "-C", HOST, "Tools/TwoLayer"])
    tool = os.path.join(HOST, "Tools", "TwoLayer")
    tmp = str(tmp_path)
    m = two_layer(100, 100, 1)
    for fmt, ext in ((1, ".mtx"), (2, ".lmf")):
        cfg = setup_case(tmp, fmt=fmt)
        for suffix in ("vp", "vs", "density"):
            os.remove(os.path.join(tmp, "model", "model." + suffix + ext))
        subprocess.run([tool, cfg], cwd=tmp, check=True)
        assert sorted(os.listdir(os.path.join(tmp, "model"))) == sorted("model." + sfx + e for sfx in ("vp", "vs", "density") for e in ((".mtx", ".lmf") if fmt == 2 else (".mtx",)))
        for key, suffix in (("velocityP", "vp"), ("velocityS", "vs"), ("density", "density")):
            path = os.path.join(tmp, "model", "model." + suffix + ext)
            if fmt == 1:
                v = read_mtx(path).ravel()
            else:
                raw = open(path, "rb").read()
                v = np.frombuffer(raw[20:], "<f4")
            assert np.array_equal(v.astype(np.float32), m[key]), (fmt, key)
    p = subprocess.run([tool], capture_output=True, text=True)
    assert p.returncode == 2 and "No configuration file given" in p.stdout


def test_driver_error_behaviour(driver, tmp_path):
    tmp = str(tmp_path)
    cfg = setup_case(tmp)
    text = open(cfg).read()
    open(cfg, "w").write(text.replace("DT=2.0e-03", "DT=2.0e-02"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "Courant-Friedrichs-Lewy-Criterion is not met" in p.stdout + p.stderr
    open(cfg, "w").write(text.replace("spatialFDorder=12", "spatialFDorder=7"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "Unsupported spatialFDorder" in p.stderr
    open(cfg, "w").write(text.replace("equationType=elastic", "equationType=visco"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "Unkown" in p.stderr
    open(cfg, "w").write(text.replace("NX=100", "#NX=100"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "Parameter NX: Not found in Configuration file!" in p.stderr
    p = subprocess.run([driver], capture_output=True, text=True)
    assert p.returncode != 0 and "No configuration file given" in p.stdout
    open(cfg, "w").write(text.replace("useStencilMatrix=1", "useStencilMatrix=0\nuseHybridFreeSurface=1"))
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "hybrid matrix without stencil matrix" in p.stderr
    open(cfg, "w").write(text + "ShotDomainDefinition=2\n")
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "ShotDomainDefinition" in p.stderr


def test_driver_coordinate_files_and_domain_definition(driver, tmp_path):
    """writeCoordinate (Simulation.cpp:151-154): the grid coordinates of every point as three vectors; ShotDomainDefinition = 1 (one
    domain per node) and useHybridFreeSurface = 1 (the image-method operator as stencil + sparse part) are accepted."""
    tmp = str(tmp_path)
    cfg = with_keys(setup_case(tmp, T=0.05), writeCoordinate=1, coordinateFilename="model/coordinates", ShotDomainDefinition=1, NumShotDomains=4, useHybridFreeSurface=1)
    log = run(driver, cfg, tmp).stdout
    assert "1 shot domain(s)" in log
    x, y, z = (read_mtx(os.path.join(tmp, "model", "coordinates%s.mtx" % a)).ravel() for a in "XYZ")
    k = np.arange(100 * 100)
    assert np.array_equal(x, k % 100) and np.array_equal(y, k // 100) and not z.any()


@pytest.mark.gpu
def test_product_driver_on_gpu(tmp_path):
    """The product binary host/Simulation (CUDA library) on the same par/ci case; golden from the reference."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    cfg = setup_case(str(tmp_path))
    run(os.path.join(HOST, "Simulation"), cfg, str(tmp_path))
    s = read_mtx(str(tmp_path / "seismograms" / "seismogram.shot_1.vy.mtx"))
    assert rel_l2(s, golden("seismogram.2D.elastic.ref.vy.mtx")) <= 1.0e-5


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run_product(cfg, tmp, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([os.path.join(HOST, "Simulation"), cfg], cwd=tmp, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


def _outputs(tmp):
    out = {}
    for d in ("seismograms", "wavefields"):
        for f in sorted(os.listdir(os.path.join(tmp, d))):
            out[d + "/" + f] = open(os.path.join(tmp, d, f), "rb").read()
    return out


MULTISHOT_SRC = "1 20 0 0 2 1 1 5.0 5.0 0.0\n2 40 5 0 1 1 1 8.0 1.0 0.0\n-2 60 5 0 3 1 4 6.0 2.0 0.05\n3 70 8 0 3 1 1 5.0 3.0 0.0\n"


def _multishot_case(tmp, **kw):
    """three shots (shot 2 fires two sources together), receivers per shot, resampled + normalised SU / mtx seismograms,
    snapshots of the first-half-step fields: the driver features of SURVEY.md 8(f) rank 2 on the real library"""
    par = dict(sources=MULTISHOT_SRC, rps=1, sdt="4.0e-03", norm=1, T=0.4, snap=1, kvar=0, stencil=0)
    par.update(kw)
    cfg = setup_case(tmp, **par)
    open(os.path.join(tmp, "acq", "receiver.shot_1.txt"), "w").write("30 0 0 3\n50 2 0 2\n")
    open(os.path.join(tmp, "acq", "receiver.shot_2.txt"), "w").write("35 3 0 1\n36 3 0 3\n37 3 0 3\n")
    open(os.path.join(tmp, "acq", "receiver.shot_3.txt"), "w").write("10 1 0 3\n90 60 0 2\n")
    return cfg


@pytest.mark.gpu
def test_product_driver_multishot_features_on_gpu(tmp_path):
    """Multi-shot run of the product binary: file set, shapes, and the seismogram of shot 2 against the library driven
    through its Python binding (plumbing of sources-per-shot / receivers-per-shot / resampling / normalisation)."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    tmp = str(tmp_path)
    cfg = _multishot_case(tmp)
    _run_product(cfg, tmp, {"WS_NUM_GPUS": "1"})
    files = sorted(os.listdir(os.path.join(tmp, "seismograms")))
    assert files == ["seismogram.shot_1.vx.mtx", "seismogram.shot_1.vy.mtx", "seismogram.shot_2.p.mtx", "seismogram.shot_2.vy.mtx", "seismogram.shot_3.vx.mtx",
                     "seismogram.shot_3.vy.mtx"]
    vy2 = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_2.vy.mtx"))
    assert vy2.shape == (2, 100) and np.abs(vy2).max() <= 1.0
    from wsharness import Solver, idx1d, ricker
    case = ci_case("2D.elastic", nt=200)
    t = np.arange(200, dtype=np.float32) * np.float32(2e-3)
    tau = (t - np.float32(1.2 / 6.0 + 0.05)) * np.float32(np.pi * 6.0)
    fg = (np.float32(2.0) * (np.float32(-2.0) * tau) * np.exp(-tau * tau)).astype(np.float32)
    case.src = ([1, 3], [idx1d(40, 5, 0, 100, 1), idx1d(60, 5, 0, 100, 1)], np.stack([ricker(200, 2e-3, 8.0, 1.0, 0.0), fg]))
    case.rec = ([1, 3, 3], [idx1d(35, 3, 0, 100, 1), idx1d(36, 3, 0, 100, 1), idx1d(37, 3, 0, 100, 1)])
    e = case.setup(Solver(case.desc))
    e.run(0, 200)
    ref = e.seismogram()[1:]
    e.close()
    ref = (ref / np.abs(ref).max(axis=1, keepdims=True))[:, ::2]
    assert rel_l2(vy2, ref) <= 1.0e-4
    snaps = [f for f in os.listdir(os.path.join(tmp, "wavefields")) if f.startswith("wavefield.shot_3.VY.")]
    assert len(snaps) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("domains", [2, 3])
def test_product_driver_shot_domains_on_several_gpus(tmp_path, domains):
    """NumShotDomains > 1 (Simulation.cpp:116-121, 339-369): every shot domain is a GPU of this process working on its
    block of the shots; every output file must equal the single-GPU run byte for byte."""
    if _gpu_count() < domains:
        pytest.skip("needs %d GPUs" % domains)
    subprocess.check_call(["make", "-s", "-C", HOST])
    one, many = str(tmp_path / "one"), str(tmp_path / "many")
    os.makedirs(one), os.makedirs(many)
    _run_product(_multishot_case(one, sfmt=4), one, {"WS_NUM_GPUS": "1"})
    cfg = _multishot_case(many, sfmt=4)
    text = open(cfg).read()
    open(cfg, "w").write(text.replace("NumShotDomains=1", "NumShotDomains=%d" % domains))
    log = _run_product(cfg, many)
    assert "%d shot domain(s) x 1 GPU(s)" % domains in log
    a, b = _outputs(one), _outputs(many)
    assert sorted(a) == sorted(b) and len(a) >= 6 + 12
    for k in a:
        assert a[k] == b[k], k


@pytest.mark.gpu
def test_product_driver_spatial_slabs_on_two_gpus(tmp_path):
    """One shot domain over two GPUs (y-slabs + NCCL halo exchange inside the library, the reference's spatial partitioning
    over commShot): seismograms and snapshots equal the single-GPU run byte for byte (every grid point sees the same
    arithmetic whatever the decomposition)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    subprocess.check_call(["make", "-s", "-C", HOST])
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    os.makedirs(one), os.makedirs(two)
    _run_product(_multishot_case(one, snap=3), one, {"WS_NUM_GPUS": "1"})
    cfg = _multishot_case(two, snap=3)
    text = open(cfg).read()
    open(cfg, "w").write(text + "GPUsPerShotDomain=2\npartitioning=0\n")
    log = _run_product(cfg, two)
    assert "1 shot domain(s) x 2 GPU(s)" in log
    a, b = _outputs(one), _outputs(two)
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k] == b[k], k


def test_driver_sources_from_su(driver, tmp_path):
    """initSourcesFromSU=1 (Sources.cpp:235, 511; suHandler.cpp:14-27, 65-86): source position and type from the trace headers of
    <SourceFilename>.<component>.su, the signal from the traces of the SU file SourceSignalFilename.  Both files are produced by
    a first run of the driver itself (seismogram headers carry the source position, writeSource=1 writes the wavelet), and the
    second run must reproduce the first run's seismogram."""
    tmp = str(tmp_path)
    os.makedirs(os.path.join(tmp, "SourceSignal"), exist_ok=True)
    cfg = setup_case(tmp, receivers="30 0 0 3\n", sfmt=4, T=0.3)
    text = open(cfg).read().replace("writeSource=0", "writeSource=1\nwriteSourceFilename=SourceSignal/Source")
    open(cfg, "w").write(text)
    run(driver, cfg, tmp)
    _, first = read_su(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.su"))
    _, wavelet = read_su(os.path.join(tmp, "SourceSignal", "Source.shot_1.vx.su"))
    assert first.shape == (1, 150) and wavelet.shape == (1, 150) and np.abs(first).max() > 0
    os.makedirs(os.path.join(tmp, "acq_su"), exist_ok=True)
    # the seismogram file of a VY receiver becomes the geometry file of a VX source: only its headers (sx, sdepth) are read
    os.rename(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.su"), os.path.join(tmp, "acq_su", "src.vx.su"))
    text = text.replace("initSourcesFromSU=0", "initSourcesFromSU=1").replace("SourceFilename=acq/sources", "SourceFilename=acq_su/src")
    text = text.replace("writeSource=1", "writeSource=0") + "SourceSignalFilename=SourceSignal/Source.shot_1.vx.su\n"
    open(cfg, "w").write(text)
    run(driver, cfg, tmp)
    # the SU sources carry shot number 0 (the reference value-initialises the settings it does not read from the headers)
    _, again = read_su(os.path.join(tmp, "seismograms", "seismogram.shot_0.vy.su"))
    assert np.array_equal(again, first)


CONFIG_VARGRID = """# par/ci/configuration_ci.{dim}.acoustic.varGrid.txt of the reference, transcribed (+ edgePolicy / exactArithmetic, B200 extensions)
dimension={dim}
equationType=acoustic
useVariableGrid=1
partitioning=2
useVariableFDoperators=1
gridConfigurationFilename=gridConfig.txt
NX={nx}
NY=303
NZ={nz}
NumShotDomains=1
DH=17
DT=2.0e-03
T={T}
spatialFDorder=2
ModelRead=0
ModelFilename=model/model
FileFormat=1
FreeSurface=1
DampingBoundary=2
BoundaryWidth=30
DampingCoeff=8.0
VMaxCPML=3500.0
CenterFrequencyCPML=5.0
NPower=4.0
numRelaxationMechanisms=0
relaxationFrequency=0
velocityP=3500
velocityS=2000
rho=2000
tauP=0.0
tauS=0.0
SourceFilename=acq/sources
ReceiverFilename=acq/receiver
SeismogramFilename=seismograms/seismogram
SeismogramFormat=2
initSourcesFromSU=0
initReceiverFromSU=0
seismoDT=2.0e-03
normalizeTraces=0
useReceiversPerShot=0
writeSource=0
snapType=0
WavefieldFileName=wavefields/wavefield
tFirstSnapshot=0
tLastSnapshot=2
tIncSnapshot=0.05
verbose=0
edgePolicy={pol}
exactArithmetic={exact}
"""


def setup_vargrid_case(tmp, dim, T, pol=0, exact=1):
    for d in ("acq", "seismograms"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    open(os.path.join(tmp, "gridConfig.txt"), "w").write("#interface\tdhfactor\tFDOrder\n0\t\t1\t\t2\n30\t\t3\t\t6\n150\t\t1\t\t2\n200\t\t3\t\t6\t\n")
    if dim == 2:
        src, rec, nx, nz = "0 150 20 0 1 1 1 5.0 5.0 0.0\n", "150 20  0 1\n150 90  0 1\n150 170 0 1\n150 239 0 1\n", 305, 1
    else:
        src, rec, nx, nz = "0 50 20 50 1 1 1 5.0 5.0 0.0\n", "50 20  50 1\n51 90  51 1\n50 170 50 1\n51 239 51 1\n", 104, 104
    open(os.path.join(tmp, "acq", "sources.txt"), "w").write(src)
    open(os.path.join(tmp, "acq", "receiver.txt"), "w").write(rec)
    cfg = os.path.join(tmp, "configuration.txt")
    open(cfg, "w").write(CONFIG_VARGRID.format(dim="%dD" % dim, nx=nx, nz=nz, T=T, pol=pol, exact=exact))
    return cfg


def _oracle_vargrid(dim, nt, pol=0, **kw):
    from test_oracle_golden import vargrid_ci_case
    o, gname = vargrid_ci_case(dim, nt, edge_policy=pol, **kw)
    o.run(0, nt)
    s = o.seismogram()
    o.close()
    return s, gname


def test_driver_variable_grid_2d_reproduces_oracle_and_golden(driver, tmp_path):
    """useVariableGrid=1 + useVariableFDoperators=1: the host layer assembles the operators of the layered
    grid (IrregularGrid.cpp), the library runs them in operator-given mode.  In exact-arithmetic mode the seismogram is
    bit-identical to the oracle's, and with the truncating edge policy it passes the reference's CI gate against
    par/ci/seismogram.2D.acoustic.varGrid.ref.p.mtx (emulated library here; test_product_driver_variable_grid_on_gpu runs the CUDA one)."""
    from wsharness import reference_gate
    tmp = str(tmp_path)
    cfg = setup_vargrid_case(tmp, 2, 2)
    log = run(driver, cfg, tmp).stdout
    # Simulation.cpp:100-108, CheckParameter.hpp:22-49: what the grid fitting did to the configuration
    assert "Number of gridpoints in layer: 3 =" in log and "Percentage of gridpoints" in log
    s = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
    ref, gname = _oracle_vargrid(2, 1000)
    assert s.shape == (4, 1000) and np.abs(s).max() > 0
    assert np.array_equal(s, ref)
    assert reference_gate(s, golden(gname)) <= 5.0e-7
    # the literal order reduction of the reference's current assembly (edgePolicy follows useStencilMatrix = 0 by default)
    cfg = setup_vargrid_case(tmp, 2, 0.6, pol=1)
    run(driver, cfg, tmp)
    s = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
    ref, _ = _oracle_vargrid(2, 300, pol=1)
    assert np.array_equal(s, ref)


def _vargrid_abs_case(tmp, free_surface, T=1.0):
    cfg = setup_vargrid_case(tmp, 2, T)
    text = open(cfg).read().replace("DampingBoundary=2", "DampingBoundary=1").replace("FreeSurface=1", "FreeSurface=%d" % free_surface)
    open(cfg, "w").write(text)
    return cfg


def test_driver_variable_grid_abs_frame(driver, tmp_path):
    """DampingBoundary = 1 on a variable grid: the Cerjan frame as the reference's sparse vector (ABS2D.cpp:110-178 on the coordinates of the
    points: a coarse layer picks every third coefficient), applied before the interpolation of the pressure; bit-identical to the oracle
    in exact-arithmetic mode, with and without a free surface."""
    for fs in (1, 0):
        tmp = str(tmp_path / ("fs%d" % fs))
        os.makedirs(tmp)
        run(driver, _vargrid_abs_case(tmp, fs), tmp)
        s = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
        ref, _ = _oracle_vargrid(2, 500, damping=1, free_surface=fs)
        assert s.shape == (4, 500) and np.abs(s).max() > 0
        assert np.array_equal(s, ref), fs
    undamped, _ = _oracle_vargrid(2, 500, damping=0, free_surface=0)
    assert rel_l2(ref, undamped) > 1.0e-3  # the wave has reached the frame


@pytest.mark.gpu
def test_product_driver_variable_grid_abs_frame_on_gpu(tmp_path):
    """The same on the CUDA library (operator-given mode with ws_set_abs_profile)."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    tmp = str(tmp_path)
    _run_product(_vargrid_abs_case(tmp, 1), tmp, {"WS_NUM_GPUS": "1"})
    s = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
    ref, _ = _oracle_vargrid(2, 500, damping=1, free_surface=1)
    assert np.array_equal(s, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,T", [(2, 2), (3, 2)])
def test_product_driver_variable_grid_on_gpu(tmp_path, dim, T):
    """The product binary on the reference's two variable-grid CI cases: bit-identical to the oracle in exact-arithmetic mode,
    within the reference's CI gate of the goldens; the default (FMA) arithmetic stays within 1e-5 of the exact one."""
    from wsharness import reference_gate
    subprocess.check_call(["make", "-s", "-C", HOST])
    tmp = str(tmp_path)
    nt = 1000
    cfg = setup_vargrid_case(tmp, dim, T)
    _run_product(cfg, tmp, {"WS_NUM_GPUS": "1"})
    s = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
    ref, gname = _oracle_vargrid(dim, nt)
    assert np.array_equal(s, ref)
    assert reference_gate(s, golden(gname)) <= 5.0e-7
    cfg = setup_vargrid_case(tmp, dim, T, exact=0)
    _run_product(cfg, tmp, {"WS_NUM_GPUS": "1"})
    f = read_lmf_matrix(os.path.join(tmp, "seismograms", "seismogram.shot_0.p.lmf"))
    assert rel_l2(f, s) <= 1.0e-5


# ---------------------------------------------------------------------------------------------------------------------
# hooks of the inversion workflow: wavefield operators, compensation
# ---------------------------------------------------------------------------------------------------------------------
CONFIG_2D_TMEM = """# 2-D TMEz modelling, homogeneous lossy medium (ModelRead=0)
dimension=2D
equationType=tmem
NX=60
NY=50
NZ=1
UseVariableGrid=0
useVariableFDoperators=0
useStencilMatrix=1
NumShotDomains=1
DH=0.02
DT=2.0e-11
T={T}
spatialFDorder=4
ModelRead=0
ModelFilename=model/model
fileFormat=1
mur=1
sigma=0.02
epsilonr=4
numRelaxationMechanisms=0;
FreeSurface=0
DampingBoundary=2
BoundaryWidth=8
DampingCoeff=8.0
VMaxCPML=3.0e8
CenterFrequencyCPML=1.0e8
NPower=4
KMaxCPML=1
SourceFilename=acq/sources
ReceiverFilename=acq/receiver
SeismogramFilename=seismograms/seismogram
initSourcesFromSU=0
initReceiverFromSU=0
SeismogramFormat=1
normalizeTraces=0
useReceiversPerShot=0
writeSource=0
seismoDT=2.0e-11
snapType=0
WavefieldFileName=wavefields/wavefield
tFirstSnapshot=0
tLastSnapshot=2
tIncSnapshot=0.1
verbose=0
kernelVariant={kvar}
compensation={comp}
"""


def setup_tmem_case(tmp, T="8.0e-10", comp=1, kvar=1):
    for d in ("model", "acq", "seismograms", "wavefields"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    open(os.path.join(tmp, "acq", "sources.txt"), "w").write("# sourceNo X Y Z type wType wShape fc amp tShift\n1 30 25 0 1 1 1 1.0e8 1.0 0.0\n")
    open(os.path.join(tmp, "acq", "receiver.txt"), "w").write("# X Y Z type\n40 30 0 1\n22 20 0 1\n")
    cfg = os.path.join(tmp, "configuration.txt")
    open(cfg, "w").write(CONFIG_2D_TMEM.format(T=T, comp=comp, kvar=kvar))
    return cfg


def tmem_oracle_with_compensation(nt, comp_on):
    from wsharness import TYPE, idx1d, make_desc, ricker
    nx, ny = 60, 50
    eps0, mu0 = np.float32(8.8541878176e-12), np.float32(1.2566370614e-6)
    d = make_desc(2, "tmem", nx, ny, 1, dh=0.02, dt=2e-11, nt=nt, fd_order=4, edge_policy=0, free_surface=0, damping=2, boundary_width=8,
                  vmax_cpml=3e8, fc_cpml=1e8, exact_arith=0)
    n = nx * ny
    m = dict(magneticPermeability=np.full(n, mu0, np.float32), electricConductivity=np.full(n, 0.02, np.float32),
             dielectricPermittivity=np.full(n, np.float32(4) * eps0, np.float32))
    o = Oracle(d)
    for k, v in m.items():
        o.set_material(k, v)
    o.prepare()
    o.set_sources([TYPE["P"]], [idx1d(30, 25, 0, nx, 1)], ricker(nt, 2e-11, 1e8, 1.0, 0.0)[None, :])
    o.set_receivers([TYPE["P"], TYPE["P"]], [idx1d(40, 30, 0, nx, 1), idx1d(22, 20, 0, nx, 1)])
    o.reset()
    v = (m["electricConductivity"] / m["dielectricPermittivity"]) * np.float32(1 * 2e-11)  # Modelparameter.cpp:135-138
    comp = np.exp(v.astype(np.float32)).astype(np.float32)
    for t in range(nt):
        o.run(t, t + 1)
        if comp_on:
            for f in ("HX", "HY", "EZ"):
                o.set_wavefield(f, o.wavefield(f) * comp)
    return o.seismogram()


def test_driver_compensation_in_the_time_loop(driver, tmp_path):
    """`compensation=1` (Simulation.cpp:441-456): the wavefields are multiplied by exp(sigma/eps DT) after every step; checked
    against the oracle stepped one step at a time with the multiplication in numpy, and against the uncompensated run."""
    tmp = str(tmp_path)
    run(driver, setup_tmem_case(tmp, comp=1), tmp)
    s = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.ez.mtx"))
    want, plain = tmem_oracle_with_compensation(40, True), tmem_oracle_with_compensation(40, False)
    assert s.shape == want.shape == (2, 40)
    assert rel_l2(s, want) <= 1.0e-5
    assert rel_l2(plain, want) > 1.0e-2  # the factor matters in this medium
    run(driver, setup_tmem_case(tmp, comp=0), tmp)
    assert rel_l2(read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.ez.mtx")), plain) <= 1.0e-5


def build_wavefield_ops(tmp, emu):
    subprocess.check_call(["make", "-s", "-C", HOST, "libSimulation_host.a"])
    exe = os.path.join(tmp, "test_wavefield_ops")
    src = os.path.join(ROOT, "tests", "host", "test_wavefield_ops.cpp")
    lib = ["-L" + EMU_DIR, "-lwavesim_emu", "-Wl,-rpath," + EMU_DIR, "-fopenmp"] if emu else \
        ["-L" + os.path.join(ROOT, "wave-simulation_b200", "csrc"), "-lwavesim_cuda", "-Wl,-rpath," + os.path.join(ROOT, "wave-simulation_b200", "csrc")]
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I" + HOST, src, os.path.join(HOST, "libSimulation_host.a")] + lib + ["-o", exe])
    return exe


@pytest.mark.parametrize("which", ["elastic", "tmem"])
def test_host_wavefield_operators(driver, tmp_path, which):
    """Wavefields operator=, -=, +=, *= scalar, *= vector on whole objects and Modelparameter::getCompensation through the host
    classes (tests/host/test_wavefield_ops.cpp), on the emulation build of the library."""
    tmp = str(tmp_path)
    cfg = setup_case(tmp, T=0.1) if which == "elastic" else setup_tmem_case(tmp)
    p = subprocess.run([build_wavefield_ops(tmp, True), cfg], cwd=tmp, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "wavefield operator tests OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["elastic", "tmem"])
def test_host_wavefield_operators_on_gpu(tmp_path, which):
    subprocess.check_call(["make", "-s", "-C", HOST])
    tmp = str(tmp_path)
    cfg = setup_case(tmp, T=0.1, kvar=0) if which == "elastic" else setup_tmem_case(tmp, kvar=0)
    p = subprocess.run([build_wavefield_ops(tmp, False), cfg], cwd=tmp, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "wavefield operator tests OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
def test_product_driver_compensation_on_gpu(tmp_path):
    subprocess.check_call(["make", "-s", "-C", HOST])
    tmp = str(tmp_path)
    run(os.path.join(HOST, "Simulation"), setup_tmem_case(tmp, comp=1, kvar=0), tmp)
    s = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.ez.mtx"))
    assert rel_l2(s, tmem_oracle_with_compensation(40, True)) <= 1.0e-5


# source encoding, random shots, shot increment, per-shot model cut-outs (Sources.cpp:500-714, AcquisitionSettings.hpp:136-483,
# Simulation.cpp:233-282, 339-398)
FOUR_SHOTS = "".join("%d %d 0 0 2 1 1 5.0 %.1f 0.0\n" % (k + 1, 20 + 12 * k, 5.0 + k) for k in range(4))


def seismograms_of(tmp, comp="vy", cop=None):
    """Seismograms by shot number.  cop = the shot numbers of a survey whose shots have one source and record one trace each: the
    reference gathers such traces into ONE common-offset profile `<SeismogramFilename>.<comp>` (row = shot index) and writes no
    per-shot files (Simulation.cpp:356-361, 535-560; Seismogram.cpp:84-96)."""
    out = {}
    for f in os.listdir(os.path.join(tmp, "seismograms")):
        if f.endswith("." + comp + ".mtx"):
            if "shot_" in f:
                out[int(f.split("shot_")[1].split(".")[0])] = read_mtx(os.path.join(tmp, "seismograms", f))
            else:
                assert cop is not None and f == "seismogram.%s.mtx" % comp, f
                profile = read_mtx(os.path.join(tmp, "seismograms", f))
                assert profile.shape[0] == len(cop)
                out.update({no: profile[k:k + 1] for k, no in enumerate(cop)})
    if cop is not None:
        assert not [f for f in os.listdir(os.path.join(tmp, "seismograms")) if "shot_" in f]
    return out


def with_keys(cfg, **kw):
    text = open(cfg).read()
    for k, v in kw.items():
        text = "\n".join(ln for ln in text.split("\n") if not ln.startswith(k + "=")) + "\n%s=%s\n" % (k, v)
    open(cfg, "w").write(text)
    return cfg


def write_mark(tmp, rows, coordinate=False):
    """<ReceiverFilename>.mark.mtx: one row per shot of the source file, [shot number | one mark per receiver of the receiver file]
    (Receivers.cpp:262-276); dense MatrixMarket array or, as LAMA writes a sparse matrix, coordinate format."""
    rows = np.asarray(rows)
    with open(os.path.join(tmp, "acq", "receiver.mark.mtx"), "w") as f:
        if coordinate:
            nz = [(i + 1, j + 1, rows[i, j]) for i in range(rows.shape[0]) for j in range(rows.shape[1]) if rows[i, j] != 0]
            f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (rows.shape[0], rows.shape[1], len(nz)))
            f.write("".join("%d %d %g\n" % e for e in nz))
        else:
            f.write("%%%%MatrixMarket matrix array real general\n%d %d\n" % rows.shape)
            f.write("".join("%g\n" % v for v in rows.T.ravel()))


TWO_RECEIVERS = "30 2 0 3\n70 3 0 3\n"
MARKS = [[1, 1, 1], [2, 1, 0], [3, 0, 1], [4, 1, 1]]  # shots 1 and 4 record both receivers, shot 2 the first, shot 3 the second


def test_driver_source_encoding(driver, tmp_path):
    """useSourceEncode: the shots fire together in NumShotDomains supershots (numbers NumShotDomains*1e4+1+k).  The equations are
    linear, so a supershot's seismogram is the (signed) sum of the seismograms of its shots.  The reference couples the encoding to
    useReceiversPerShot = 2 (Receivers.cpp:365): a supershot records the union of the receivers its shots mark, and is decoded into
    per-shot seismograms (the marked receivers of the shot, polarity undone) after the time loop (Simulation.cpp:531-533)."""
    plain = str(tmp_path / "plain")
    run(driver, setup_case(plain, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.4), plain)
    single = seismograms_of(plain)
    assert sorted(single) == [1, 2, 3, 4]
    for mode, groups in ((2, {20001: [1, 3], 20002: [2, 4]}), (3, {20001: [1, 2], 20002: [3, 4]})):
        tmp = str(tmp_path / ("mode%d" % mode))
        cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.4, rps=2), useSourceEncode=mode, NumShotDomains=2, seedtime=7)
        write_mark(tmp, MARKS, coordinate=(mode == 3))
        run(driver, cfg, tmp)
        enc = seismograms_of(tmp)
        assert sorted(enc) == [1, 2, 3, 4] + sorted(groups)
        for no, members in groups.items():
            assert rel_l2(enc[no], sum(single[m] for m in members)) <= 2.0e-5, (mode, no)
            for m in members:  # the decoded shot = the traces of the supershot at the receivers the shot marks
                assert np.array_equal(enc[m], enc[no][np.array(MARKS[m - 1][1:]) != 0]), (mode, no, m)
            mark = read_mtx(os.path.join(tmp, "acq", "receiver.shot_%d.mark.mtx" % no)).ravel()
            assert list(mark) == [no, 1, 1]
        lines = [ln.split() for ln in open(os.path.join(tmp, "acq", "sources.encode.txt")) if not ln.startswith("#")]
        assert {int(ln[0]): [int(x) for x in ln[1:]] for ln in lines} == groups
    # mode 1: random assignment (every supershot holds numshots / NumShotDomains shots) with random polarity; a given seed repeats
    res = []
    for rep in range(2):
        tmp = str(tmp_path / ("mode1_%d" % rep))
        cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.4, rps=2), useSourceEncode=1, NumShotDomains=2, seedtime=11)
        write_mark(tmp, [[k + 1, 1, 1] for k in range(4)])
        run(driver, cfg, tmp)
        enc = seismograms_of(tmp)
        lines = [ln.split() for ln in open(os.path.join(tmp, "acq", "sources.encode.txt")) if not ln.startswith("#")]
        groups = {int(ln[0]): [int(x) for x in ln[1:]] for ln in lines}
        assert sorted(groups) == [20001, 20002] and sorted(sum(groups.values(), [])) == [1, 2, 3, 4] and all(len(g) == 2 for g in groups.values())
        for no, members in groups.items():
            basis = np.stack([single[m].ravel() for m in members], axis=1)
            coef = np.linalg.lstsq(basis, enc[no].ravel(), rcond=None)[0]
            assert np.allclose(np.abs(coef), 1.0, atol=1e-3), coef  # +-1: the random polarity
            for m, c in zip(members, coef):  # decoding undoes the polarity of the shot
                assert np.array_equal(enc[m], enc[no] if c > 0 else -enc[no])
        res.append((groups, {k: v.copy() for k, v in enc.items()}))
    assert res[0][0] == res[1][0] and all(np.array_equal(res[0][1][k], res[1][1][k]) for k in res[0][1])
    # the reference's assertion: the encoding cannot be decoded without the mark matrix
    tmp = str(tmp_path / "nomark")
    p = run(driver, with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.1), useSourceEncode=2, NumShotDomains=2), tmp, expect_ok=False)
    assert p.returncode != 0 and "useReceiversPerShot != 2" in p.stdout + p.stderr


def test_driver_frequency_encoded_supershots(driver, tmp_path):
    """gradientDomain != 0 with source encoding (Sources.cpp:584-674): every shot of a supershot fires a sine of its own frequency drawn
    from every second FFT bin up to 2 CenterFrequencyCPML; the frequencies go to <SourceFilename>.sourceFC.txt; decoding looks at
    every shot through the one-frequency DFT at its frequency (Receivers.cpp:531-540, Filter.cpp:257-275, 331-341)."""
    tmp = str(tmp_path / "fenc")
    cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.5, rps=2), useSourceEncode=2, NumShotDomains=2, seedtime=5, gradientDomain=1)
    write_mark(tmp, [[k + 1, 1, 1] for k in range(4)])
    run(driver, cfg, tmp)
    nt, dt = 250, 2e-3
    df = 1.0 / (256 * dt)
    lines = [ln.split() for ln in open(os.path.join(tmp, "acq", "sources.sourceFC.txt")) if not ln.startswith("#")]
    fcs = {int(ln[0]): [float(x) for x in ln[1:]] for ln in lines}
    assert sorted(fcs) == [20001, 20002]
    allowed = [2 * df, 4 * df, 6 * df]
    for no, f in fcs.items():
        assert len(f) == 2 and f[0] != f[1] and all(min(abs(x - a) for a in allowed) < 1e-3 for x in f)
    groups = {20001: [1, 3], 20002: [2, 4]}
    enc = seismograms_of(tmp)
    t = np.arange(nt) * dt
    for no, members in groups.items():
        # the supershot = the sum of its shots fired alone with their sines (waveletShape 9)
        alone = str(tmp_path / ("alone%d" % no))
        src = "".join("%d %d 0 0 2 1 9 %.7g %.1f 0.0\n" % (m, 20 + 12 * (m - 1), f, 5.0 + (m - 1)) for m, f in zip(members, fcs[no]))
        run(driver, setup_case(alone, sources=src, receivers=TWO_RECEIVERS, T=0.5), alone)
        parts = seismograms_of(alone)
        assert rel_l2(enc[no], sum(parts[m] for m in members)) <= 1.0e-4  # (the file carries the frequencies with 6 digits)
        for m, f in zip(members, fcs[no]):  # the decoded shot: one-frequency DFT of the supershot at the bin of the shot, scaled to maximum 1
            k = int(np.ceil(np.float32(f) / np.float32(df)))
            ph = 2 * np.pi * k * df * t
            coef = (enc[no] * np.exp(-1j * ph)).sum(axis=1, keepdims=True) * (2.0 / 256)
            want = np.real(coef * np.exp(1j * ph))
            want /= np.abs(want).max()
            assert rel_l2(enc[m], want) <= 1.0e-4, (no, m)


def test_driver_receivers_by_mark_matrix(driver, tmp_path):
    """useReceiversPerShot = 2: one receiver file and a mark matrix; every shot records the receivers its row marks (here also with the
    shot increment: the rows of the mark matrix are those of the source file, not of the selection)."""
    plain = str(tmp_path / "plain")
    run(driver, setup_case(plain, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.12), plain)
    single = seismograms_of(plain)
    tmp = str(tmp_path / "marks")
    write_mark_for = setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.12, rps=2)
    write_mark(tmp, MARKS)
    run(driver, write_mark_for, tmp)
    got = seismograms_of(tmp)
    assert sorted(got) == [1, 2, 3, 4]
    for no in got:
        assert np.array_equal(got[no], single[no][np.array(MARKS[no - 1][1:]) != 0]), no
    assert not [f for f in os.listdir(os.path.join(tmp, "acq")) if f.endswith(".mark.mtx") and "shot_" in f]  # marks are written with the encoding only
    tmp = str(tmp_path / "incr")  # shots 12 grid points = 600 m apart: shotIncr 1200 keeps shots 1 and 3 (rows 0 and 2)
    cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.12, rps=2), shotIncr=1200)
    write_mark(tmp, [[1, 0, 1], [2, 1, 0], [3, 1, 1], [4, 1, 1]], coordinate=True)
    run(driver, cfg, tmp)
    got = seismograms_of(tmp)
    assert sorted(got) == [1, 3]
    assert np.array_equal(got[1], single[1][1:2]) and np.array_equal(got[3], single[3])
    # a mark matrix of the wrong shape is refused
    tmp = str(tmp_path / "bad")
    cfg = setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.1, rps=2)
    write_mark(tmp, [r[:2] for r in MARKS])
    p = run(driver, cfg, tmp, expect_ok=False)
    assert p.returncode != 0 and "mark matrix" in p.stdout + p.stderr


def test_driver_random_shots_and_shot_increment(driver, tmp_path):
    """Four shots of one source each, recorded by two receivers (a file per shot) or by one (the traces are gathered into a
    common-offset profile, summed over the shot domains: SeismogramHandler::sumShotDomain)."""
    plain = str(tmp_path / "plain")
    run(driver, setup_case(plain, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.12), plain)
    single = seismograms_of(plain)
    assert sorted(single) == [1, 2, 3, 4] and single[1].shape == (2, 60)
    for mode in (1, 2, 3):  # two passes of two shots each: every shot exactly once (maxcount = 1)
        tmp = str(tmp_path / ("rand%d" % mode))
        run(driver, with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.12), useRandomSource=mode, NumShotDomains=2, seedtime=3), tmp)
        got = seismograms_of(tmp)
        assert sorted(got) == [1, 2, 3, 4]
        assert all(np.array_equal(got[k], single[k]) for k in got)
    # one receiver: common-offset profile (and the profile of the source signals with writeSource), one and two shot domains
    for domains in (1, 2):
        tmp = str(tmp_path / ("cop%d" % domains))
        cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers="30 2 0 3\n", T=0.12), NumShotDomains=domains, writeSource=1, writeSourceFilename="seismograms/source")
        run(driver, cfg, tmp)
        assert sorted(os.listdir(os.path.join(tmp, "seismograms"))) == ["seismogram.vy.mtx", "source.vx.mtx"]
        got = seismograms_of(tmp, cop=[1, 2, 3, 4])
        assert all(np.array_equal(got[k], single[k][0:1]) for k in got)
        src = read_mtx(os.path.join(tmp, "seismograms", "source.vx.mtx"))
        from wsharness import ricker
        assert src.shape == (4, 60) and all(rel_l2(src[k], ricker(60, 2e-3, 5.0, 5.0 + k, 0.0)) <= 1e-6 for k in range(4))
    # shotIncr = 200 m on a line of shots 2 grid points (100 m) apart: every second shot (Sources.cpp:516-553)
    tmp = str(tmp_path / "incr")
    six = "".join("%d %d 0 0 2 1 1 5.0 5.0 0.0\n" % (k + 1, 20 + 2 * k) for k in range(6))
    run(driver, with_keys(setup_case(tmp, sources=six, T=0.1), shotIncr=200), tmp)
    assert sorted(seismograms_of(tmp, cop=[1, 3, 5])) == [1, 3, 5]
    lines = [ln.split() for ln in open(os.path.join(tmp, "acq", "sources.shotIncr.txt")) if not ln.startswith("#")]
    assert lines == [["1", "1"], ["3", "3"], ["5", "5"]]


def test_driver_stream_config_model_per_shot(driver, tmp_path):
    """useStreamConfig: every shot works on its own NX-wide cut-out of a big model (cut where the shot lies relative to the first
    one), its sources / receivers are given in the big model; compared with the oracle run on the cut-outs themselves."""
    tmp = str(tmp_path)
    nxb, nx, ny, nt = 160, 100, 100, 200
    cfg = setup_case(tmp, sources="1 40 15 0 2 1 1 5.0 5.0 0.0\n2 90 15 0 2 1 1 5.0 5.0 0.0\n", rps=1, T=0.4)  # (a cut-out keeps its acquisition out of the boundary frame)
    big = two_layer(nxb, ny, 1)
    lateral = (1.0 + 0.15 * np.arange(nxb, dtype=np.float32) / nxb)[None, :]
    for key in big:
        big[key] = (big[key].reshape(ny, nxb) * lateral).astype(np.float32).ravel()
    for key, suffix in (("velocityP", "vp"), ("velocityS", "vs"), ("density", "density")):
        write_mtx_vector(os.path.join(tmp, "model", "big." + suffix + ".mtx"), big[key])
    text = open(cfg).read()
    open(os.path.join(tmp, "configBig.txt"), "w").write(text.replace("NX=100", "NX=%d" % nxb).replace("ModelFilename=model/model", "ModelFilename=model/big"))
    with_keys(cfg, useStreamConfig=1, streamConfigFilename="configBig.txt")
    open(os.path.join(tmp, "acq", "receiver.shot_1.txt"), "w").write("50 12 0 3\n")
    open(os.path.join(tmp, "acq", "receiver.shot_2.txt"), "w").write("100 12 0 3\n")
    run(driver, cfg, tmp)
    got = seismograms_of(tmp, cop=[1, 2])  # one source and one trace per shot: a common-offset profile
    assert sorted(got) == [1, 2]
    cut = [ln.split() for ln in open(os.path.join(tmp, "acq", "sources.cut.txt")) if not ln.startswith("#")]
    assert cut == [[str(nx * ny), str(nx), str(ny), "1"], ["1", "0", "0", "0"], ["2", "50", "0", "0"]]
    for shot, x0 in ((1, 0), (2, 50)):
        case = ci_case("2D.elastic", nt=nt)
        case.desc.edge_policy = 0
        case.materials = {k: v.reshape(ny, nxb)[:, x0:x0 + nx].ravel().copy() for k, v in big.items()}
        o = case.setup(Oracle(case.desc))
        from wsharness import TYPE, idx1d, ricker
        o.set_sources([TYPE["VX"]], [idx1d(40, 15, 0, nx, 1)], ricker(nt, 2e-3, 5.0, 5.0, 0.0)[None, :])  # both shots sit at x = 40 of their cut-out
        o.set_receivers([TYPE["VY"]], [idx1d(50, 12, 0, nx, 1)])
        o.reset()
        o.run(0, nt)
        assert rel_l2(got[shot], o.seismogram()) <= 1.0e-5, shot
        # the model of the shot was written next to the big one (Simulation.cpp:393)
        vp = read_mtx(os.path.join(tmp, "model", "model.shot_%d.vp.mtx" % shot)).ravel()
        assert np.array_equal(vp.astype(np.float32), case.materials["velocityP"])
    assert not np.allclose(got[1], got[2], rtol=1e-3)  # the lateral gradient makes the two cut-outs differ


def test_driver_trace_gain_and_instantaneous_outputs(driver, tmp_path):
    """normalizeTraces = 3 (automatic gain control, the gain function written as .inverseAGC) and 4 (envelope with a water level),
    instantaneousTraces = 1 / 2 (a second file with the envelope / the reference's phase) against numpy restatements of
    Seismogram.cpp:215-385 and Common.hpp:272-340."""
    from scipy.signal import hilbert
    raw_dir = str(tmp_path / "raw")
    run(driver, setup_case(raw_dir, receivers=TWO_RECEIVERS, T=0.5), raw_dir)
    raw = seismograms_of(raw_dir)[1].astype(np.float64)
    nt = raw.shape[1]
    unit = raw / np.linalg.norm(raw, axis=1, keepdims=True)

    def envelope(x):  # zero-padded to the next power of two of nt - 1, like Common::calcEnvelope
        n = 1 << int(np.ceil(np.log2(nt - 1)))
        return np.abs(hilbert(np.concatenate([x, np.zeros((x.shape[0], n - nt))], axis=1), axis=1))[:, :nt]

    # automatic gain control: mean square over the window [t - NAGC, t + NAGC - 1] (clipped) of the l2-normalised trace plus a water
    # level, kept as a RUNNING fp32 sum from the end of the trace backwards (its cancellation error is part of the reference's result)
    tmp = str(tmp_path / "agc")
    run(driver, setup_case(tmp, receivers=TWO_RECEIVERS, T=0.5, norm=3), tmp)
    nagc = min(int(round(1.0 / (5.0 * 2e-3))), nt // 2)
    f32 = np.float32
    unit32 = (raw.astype(f32) / np.sqrt((raw ** 2).sum(axis=1, keepdims=True)).astype(f32)).astype(f32)
    gain = np.empty_like(unit32)
    exact = np.empty_like(unit)
    for r in range(unit32.shape[0]):
        x2 = (unit32[r] * unit32[r]).astype(f32)
        norm = f32(np.sqrt((unit32[r].astype(np.float64) ** 2).sum()))
        level = f32(f32(f32(norm * norm) / f32(nt)) * f32(1e-3))
        total, nwin = f32(0), f32(nagc)
        for t in range(nt - nagc, nt):
            total = f32(f32(total + x2[t]) + level)
        for t in range(nt - 1, -1, -1):
            if t >= nt - nagc:
                total = f32(f32(total + x2[t - nagc]) + level)
                nwin += f32(1)
            elif t >= nagc:
                total = f32(f32(total + x2[t - nagc]) - x2[t + nagc])
            else:
                total = f32(f32(total - x2[t + nagc]) - level)
                nwin -= f32(1)
            mean = f32(total / nwin)
            gain[r, t] = f32(1) / np.sqrt(mean) if mean > 0 else 0
            w = unit[r, max(0, t - nagc):min(nt, t + nagc)]
            exact[r, t] = 1.0 / np.sqrt((w ** 2).mean() + 1e-3 * (unit[r] ** 2).sum() / nt)
    assert rel_l2(gain, exact) <= 2.0e-2  # the window formula the running sum stands for
    got_gain = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.inverseAGC.mtx"))
    assert rel_l2(got_gain, gain) <= 1.0e-4
    assert rel_l2(seismograms_of(tmp)[1], unit32 * gain) <= 1.0e-4
    # envelope normalisation, and the envelope of the written traces as a second file
    tmp = str(tmp_path / "env")
    run(driver, with_keys(setup_case(tmp, receivers=TWO_RECEIVERS, T=0.5, norm=4), instantaneousTraces=1), tmp)
    env = envelope(unit)
    want = unit / (env + 1e-3 * env.max(axis=1, keepdims=True))
    got = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    assert rel_l2(got, want) <= 1.0e-4
    got_env = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.envelope.mtx"))
    assert rel_l2(got_env, envelope(got)) <= 1.0e-4
    # instantaneousTraces = 2: atan2(-x, x), the reference leaves the Hilbert transform out (Common.hpp:299-302)
    tmp = str(tmp_path / "phase")
    run(driver, with_keys(setup_case(tmp, receivers=TWO_RECEIVERS, T=0.5), instantaneousTraces=2), tmp)
    got = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.mtx"))
    phase = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_1.vy.instantaneousPhase.mtx"))
    assert np.allclose(phase, np.arctan2(-got, got), atol=1e-5)


@pytest.mark.gpu
def test_product_driver_marks_encoding_profiles_and_gain_on_gpu(tmp_path):
    """The seismogram bookkeeping of the driver on the real library: receivers by mark matrix, a supershot = the signed sum of its
    shots and decoded back into them, the common-offset profile of single-trace shots over two shot domains (when two GPUs are
    there), automatic gain control with its gain function written."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    plain = str(tmp_path / "plain")
    _run_product(setup_case(plain, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.5, kvar=0), plain, {"WS_NUM_GPUS": "1"})
    single = seismograms_of(plain)
    assert sorted(single) == [1, 2, 3, 4]
    tmp = str(tmp_path / "marks")
    cfg = setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.5, rps=2, kvar=0)
    write_mark(tmp, MARKS)
    _run_product(cfg, tmp, {"WS_NUM_GPUS": "1"})
    got = seismograms_of(tmp)
    assert all(np.array_equal(got[no], single[no][np.array(MARKS[no - 1][1:]) != 0]) for no in (1, 2, 3, 4))
    tmp = str(tmp_path / "encode")
    cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.5, rps=2, kvar=0), useSourceEncode=2, NumShotDomains=2, seedtime=7)
    write_mark(tmp, MARKS, coordinate=True)
    _run_product(cfg, tmp)
    enc = seismograms_of(tmp)
    for no, members in {20001: [1, 3], 20002: [2, 4]}.items():
        assert rel_l2(enc[no], sum(single[m] for m in members)) <= 2.0e-5
        assert all(np.array_equal(enc[m], enc[no][np.array(MARKS[m - 1][1:]) != 0]) for m in members)
    tmp = str(tmp_path / "cop")
    cfg = with_keys(setup_case(tmp, sources=FOUR_SHOTS, receivers="30 2 0 3\n", T=0.5, kvar=0), NumShotDomains=2, writeSource=1, writeSourceFilename="seismograms/source")
    _run_product(cfg, tmp)
    assert sorted(os.listdir(os.path.join(tmp, "seismograms"))) == ["seismogram.vy.mtx", "source.vx.mtx"]
    got = seismograms_of(tmp, cop=[1, 2, 3, 4])
    assert all(np.array_equal(got[k], single[k][0:1]) for k in got)
    tmp = str(tmp_path / "agc")
    _run_product(setup_case(tmp, sources=FOUR_SHOTS, receivers=TWO_RECEIVERS, T=0.5, norm=3, kvar=0), tmp, {"WS_NUM_GPUS": "1"})
    gain = read_mtx(os.path.join(tmp, "seismograms", "seismogram.shot_2.vy.inverseAGC.mtx"))
    unit = single[2] / np.linalg.norm(single[2], axis=1, keepdims=True)
    assert rel_l2(seismograms_of(tmp)[2], unit * gain) <= 1.0e-5


# frequency filters and Hilbert transform (Filter/Filter.cpp, Common/HilbertFFT.cpp)
def run_filter_tool(tmp_path, data, dt, *op):
    subprocess.check_call(["make", "-s", "-C", HOST, "libSimulation_host.a"])
    exe = str(tmp_path / "test_filter")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + HOST, os.path.join(ROOT, "tests", "host", "test_filter.cpp"),
                               os.path.join(HOST, "libSimulation_host.a"), "-o", exe])
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    np.ascontiguousarray(data, np.float32).tofile(fin)
    p = subprocess.run([exe, fin, fout, str(data.shape[0]), str(data.shape[1]), repr(dt)] + [str(x) for x in op], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return np.fromfile(fout, np.float32).reshape(data.shape)


def read_mtx_coordinate(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("%")]
    r, c, nnz = (int(x) for x in lines[0].split())
    m = np.zeros((r, c))
    for ln in lines[1:1 + nnz]:
        i, j, v = ln.split()
        m[int(i) - 1, int(j) - 1] = float(v)
    return m


def test_filter_reproduces_the_reference_fixture(tmp_path):
    """The reference's own known answer (Tests/UnitTest/FilterUnitTest.cpp): Butterworth low pass, order 4, 15 Hz, dt 1 ms, 500 samples,
    21 traces; its gate is ||filtered - ref||_2 < 0.005."""
    gold = os.path.join(ROOT, "tests", "golden")
    sig, ref = read_mtx_coordinate(os.path.join(gold, "filterTest_signal.mtx")), read_mtx_coordinate(os.path.join(gold, "filterTest_signalFiltRef.mtx"))
    assert sig.shape == ref.shape == (21, 500)
    for op in ("filter", "filter1"):
        out = run_filter_tool(tmp_path, sig, 1e-3, op, "butterworth", "lp", 4, 15.0)
        assert np.linalg.norm(out - ref) < 0.005, np.linalg.norm(out - ref)
    assert np.linalg.norm(sig - ref) > 0.1  # the filter does something to this signal


def butter_transfer(nt, dt, kind, order, fc):
    n = 1 << int(np.ceil(np.log2(nt - 1)))
    f = np.fft.fftfreq(n, dt)
    f[n // 2] = abs(f[n // 2])  # Filter.cpp:25-31 lists the Nyquist bin with a positive frequency
    k = np.arange(1, order // 2 + 1)
    poly = np.poly1d([1.0])
    for c in -2.0 * np.cos((2.0 * k + order - 1.0) / (2.0 * order) * np.pi):
        poly = poly * np.poly1d([1.0, c, 1.0])
    if order % 2:
        poly = poly * np.poly1d([1.0, 1.0])
    with np.errstate(divide="ignore", invalid="ignore"):
        s = 1j * f / fc if kind == "lp" else -1j * fc / f
        h = 1.0 / poly(s)
    h[0] = 1.0 if kind == "lp" else 0.0
    return n, h


def test_filter_and_hilbert_against_numpy(tmp_path):
    rng = np.random.default_rng(5)
    nt, dt = 700, 2e-3
    t = np.arange(nt) * dt
    sig = np.stack([np.sin(2 * np.pi * 5 * t) + 0.5 * np.sin(2 * np.pi * 60 * t + 1.0), rng.standard_normal(nt), np.exp(-((t - 0.5) / 0.05) ** 2)])
    for kind, order, fc in (("lp", 3, 20.0), ("hp", 4, 30.0), ("lp", 6, 45.0)):
        n, h = butter_transfer(nt, dt, kind, order, fc)
        want = np.real(np.fft.ifft(np.fft.fft(sig, n, axis=1) * h, axis=1))[:, :nt]
        out = run_filter_tool(tmp_path, sig, dt, "filter", "butterworth", kind, order, fc)
        assert np.abs(out - want).max() <= 2e-6 * np.abs(want).max(), (kind, order)
    n, hl = butter_transfer(nt, dt, "lp", 4, 50.0)
    _, hh = butter_transfer(nt, dt, "hp", 4, 10.0)
    want = np.real(np.fft.ifft(np.fft.fft(sig, n, axis=1) * hl * hh, axis=1))[:, :nt]
    out = run_filter_tool(tmp_path, sig, dt, "filter", "butterworth", "bp", 4, 10.0, 50.0)
    assert np.abs(out - want).max() <= 2e-6 * np.abs(want).max()
    # ideal band pass, order 0: only the FFT bin of the frequency passes
    out = run_filter_tool(tmp_path, sig, dt, "filter", "ideal", "bp", 0, 5.0)
    df = 1.0 / (n * dt)
    k = int(np.ceil(5.0 / df))
    spec = np.fft.fft(sig, n, axis=1)
    mask = np.zeros(n)
    mask[[k, n - k]] = 1.0
    want = np.real(np.fft.ifft(spec * mask, axis=1))[:, :nt]
    assert np.abs(out - want).max() <= 2e-6 * np.abs(sig).max()
    # Hilbert transform = imaginary part of the analytic signal of the zero-padded trace (scipy.signal.hilbert)
    from scipy.signal import hilbert
    out = run_filter_tool(tmp_path, sig, dt, "hilbert")
    want = np.imag(hilbert(sig, N=n, axis=1))[:, :nt]
    assert np.abs(out - want).max() <= 2e-6 * np.abs(want).max()

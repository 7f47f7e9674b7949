"""Test harness shared by tests/, bench.py and __graft_entry__.smoke().

Two ctypes front-ends with the same method set:
  * ``Oracle``  -> oracle/_ref/libwave_oracle.so   (CPU restatement of the reference, the CHECKER)
  * ``Solver``  -> wave-simulation_b200/csrc/libwavesim_cuda.so  (the PRODUCT, through its C ABI include/wavesim.h)

plus builders for the reference's CI cases (par/ci/configuration_ci.*.txt, transcribed here because
/root/reference does not exist on the GPU box) and helpers for golden seismograms (tests/golden/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libwave_oracle.so")
PRODUCT_SO = os.path.join(ROOT, "wave-simulation_b200", "csrc", "libwavesim_cuda.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libwavesim_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

EQ = dict(acoustic=0, elastic=1, viscoelastic=2, sh=3, viscosh=4, tmem=5, emem=6, viscotmem=7, viscoemem=8)
TYPE = dict(P=1, VX=2, VY=3, VZ=4, EZ=1, EX=2, EY=3, HZ=4)


class Desc(C.Structure):
    """Mirror of ws_desc (include/wavesim.h)."""
    _fields_ = [
        ("dim", C.c_int32), ("eq", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("dh", C.c_float), ("dt", C.c_float),
        ("nt", C.c_int32), ("fd_order", C.c_int32), ("edge_policy", C.c_int32),
        ("free_surface", C.c_int32), ("damping", C.c_int32), ("boundary_width", C.c_int32),
        ("damping_coeff", C.c_float), ("vmax_cpml", C.c_float), ("fc_cpml", C.c_float), ("npower", C.c_float),
        ("n_relax", C.c_int32), ("relax_freq", C.c_float * 4),
        ("exact_arith", C.c_int32), ("kernel_variant", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("device", C.c_int32),
    ]


def make_desc(dim, eq, nx, ny, nz=1, dh=50.0, dt=2e-3, nt=100, fd_order=2, edge_policy=1, free_surface=0,
              damping=0, boundary_width=10, damping_coeff=8.0, vmax_cpml=3500.0, fc_cpml=5.0, npower=4.0,
              relax_freq=(), exact_arith=0, kernel_variant=0, rank=0, nranks=1, device=0):
    d = Desc()
    d.dim, d.eq = dim, EQ[eq] if isinstance(eq, str) else eq
    d.nx, d.ny, d.nz = nx, ny, (1 if dim == 2 else nz)
    d.dh, d.dt, d.nt = dh, dt, nt
    d.fd_order, d.edge_policy, d.free_surface = fd_order, edge_policy, free_surface
    d.damping, d.boundary_width, d.damping_coeff = damping, boundary_width, damping_coeff
    d.vmax_cpml, d.fc_cpml, d.npower = vmax_cpml, fc_cpml, npower
    d.n_relax = len(relax_freq)
    for i, f in enumerate(relax_freq):
        d.relax_freq[i] = f
    d.exact_arith, d.kernel_variant = exact_arith, kernel_variant
    d.rank, d.nranks, d.device = rank, nranks, device
    return d


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def build_oracle(force=False):
    """Compile oracle/ (plain g++, recipe = oracle/Makefile). Building the checker is not using it."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ROOT, "oracle", "wave_oracle.cpp")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


class _Base:
    """Common method set over a `<prefix>_*` C API."""
    prefix = None
    lib = None

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc, what):
        if rc != 0:
            err = self._fn("last_error")
            err.restype = C.c_char_p
            raise RuntimeError("%s%s failed (%d): %s" % (self.prefix, what, rc, (err() or b"").decode()))

    def set_material(self, name, arr):
        a = _f32(arr).ravel()
        self._check(self._fn("set_material")(self.h, name.encode(), _fp(a), C.c_size_t(a.size)), "set_material")

    def get_material(self, name, n=None):
        out = np.empty(self.n_local if n is None else n, dtype=np.float32)
        self._check(self._fn("get_material")(self.h, name.encode(), _fp(out), C.c_size_t(out.size)), "get_material")
        return out

    def prepare(self):
        self._check(self._fn("prepare")(self.h), "prepare")

    def set_sources(self, types, idx, signals):
        t, i, s = _i32(types), _i32(idx), _f32(signals)
        assert s.shape == (len(t), self.desc.nt), s.shape
        self._check(self._fn("set_sources")(self.h, len(t), _ip(t), _ip(i), _fp(s)), "set_sources")

    def set_receivers(self, types, idx):
        t, i = _i32(types), _i32(idx)
        self.n_rec = len(t)
        self._check(self._fn("set_receivers")(self.h, len(t), _ip(t), _ip(i)), "set_receivers")

    def reset(self):
        self._check(self._fn("reset")(self.h), "reset")

    def step(self, t):
        self._check(self._fn("step")(self.h, t), "step")

    def run(self, t0, t1):
        self._check(self._fn("run")(self.h, t0, t1), "run")

    def seismogram(self):
        out = np.zeros((self.n_rec, self.desc.nt), dtype=np.float32)
        self._check(self._fn("get_seismogram")(self.h, _fp(out)), "get_seismogram")
        return out

    def wavefield(self, comp):
        out = np.empty(self.n_local, dtype=np.float32)
        self._check(self._fn("get_wavefield")(self.h, comp.encode(), _fp(out), C.c_size_t(out.size)), "get_wavefield")
        return out

    def set_wavefield(self, comp, arr):
        a = _f32(arr).ravel()
        self._check(self._fn("set_wavefield")(self.h, comp.encode(), _fp(a), C.c_size_t(a.size)), "set_wavefield")

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle(_Base):
    prefix = "wso_"

    def __init__(self, desc, precision=32):
        if Oracle.lib is None:
            build_oracle()
            Oracle.lib = C.CDLL(ORACLE_SO)
            Oracle.lib.wso_destroy.argtypes = [C.c_void_p]
            for f in ("set_material", "get_material", "prepare", "set_sources", "set_receivers", "reset", "step",
                      "run", "get_seismogram", "get_wavefield", "set_wavefield", "deriv_row"):
                getattr(Oracle.lib, "wso_" + f).argtypes = None
        self.desc = desc
        self.h = C.c_void_p()
        self.n_local = desc.nx * desc.ny * desc.nz
        self.n_rec = 0
        self._check(self.lib.wso_create(C.byref(desc), precision, C.byref(self.h)), "create")

    def deriv_row(self, which, row):
        cols = np.zeros(16, dtype=np.int32)
        vals = np.zeros(16, dtype=np.float32)
        n = self.lib.wso_deriv_row(self.h, which, row, _ip(cols), _fp(vals))
        if n < 0:
            self._check(n, "deriv_row")
        return cols[:n].copy(), vals[:n].copy()

    @staticmethod
    def num_threads():
        build_oracle()
        if Oracle.lib is None:
            Oracle.lib = C.CDLL(ORACLE_SO)
        return Oracle.lib.wso_num_threads()


def ricker(nt, dt, fc, amp, tshift=0.0):
    """Acquisition/SourceSignal/Ricker.cpp:29-53 evaluated by the oracle library in float."""
    build_oracle()
    if Oracle.lib is None:
        Oracle.lib = C.CDLL(ORACLE_SO)
    out = np.zeros(nt, dtype=np.float32)
    rc = Oracle.lib.wso_wavelet(1, nt, C.c_float(dt), C.c_float(fc), C.c_float(amp), C.c_float(tshift), _fp(out))
    assert rc == 0
    return out


def ricker_np(nt, dt, fc, amp, tshift=0.0):
    """Independent numpy statement of the same wavelet (float32 op by op); used by bench/product code paths that must
    not touch oracle/."""
    f = np.float32
    t = np.arange(nt, dtype=np.float32) * f(dt)
    helpv = f(1.5 / fc + tshift)
    tau = (t - helpv) * f(np.pi * fc)
    h2 = tau * tau
    e = np.exp(-h2).astype(np.float32)
    return ((f(amp) * (f(1.0) - f(2.0) * h2)) * e).astype(np.float32)


def _load_ws_lib(path):
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.ws_destroy.argtypes = [C.c_void_p]
    lib.ws_launch_count.restype = C.c_uint64
    lib.ws_launch_count.argtypes = [C.c_void_p]
    lib.ws_estimate_memory.restype = C.c_size_t
    lib.ws_stream.restype = C.c_void_p
    lib.ws_stream.argtypes = [C.c_void_p]
    lib.ws_uses_fast_kernels.argtypes = [C.c_void_p]
    return lib


class Solver(_Base):
    """The product through its C ABI. Fails loudly if the CUDA library is missing: there is no CPU fallback."""
    prefix = "ws_"
    so_path = PRODUCT_SO

    @classmethod
    def _ensure_lib(cls):
        if cls.lib is None:
            if not os.path.exists(cls.so_path):
                raise RuntimeError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                                   % cls.so_path)
            cls.lib = _load_ws_lib(cls.so_path)
        return cls.lib

    def __init__(self, desc):
        self._ensure_lib()
        self.desc = desc
        self.h = C.c_void_p()
        self.n_rec = 0
        self._check(self.lib.ws_create(C.byref(desc), C.byref(self.h)), "create")
        y0, nyl = C.c_int32(), C.c_int32()
        self._check(self.lib.ws_local_range(self.h, C.byref(y0), C.byref(nyl)), "local_range")
        self.y0, self.nyl = y0.value, nyl.value
        self.n_local = desc.nx * desc.nz * self.nyl

    def set_material_device(self, name, dev_ptr, n_local):
        self._check(self.lib.ws_set_material_device(self.h, name.encode(), C.c_void_p(dev_ptr), C.c_size_t(n_local)),
                    "set_material_device")

    def sync(self):
        self._check(self.lib.ws_sync(self.h), "sync")

    def step_host(self, t, src_samples, rec_samples):
        sp = _fp(src_samples) if src_samples is not None else None
        self._check(self.lib.ws_step_host(self.h, t, sp, _fp(rec_samples)), "step_host")

    def set_timing(self, enable):
        self._check(self.lib.ws_set_timing(self.h, int(enable)), "set_timing")

    def uses_fast_kernels(self):
        return bool(self.lib.ws_uses_fast_kernels(self.h))

    def launch_count(self):
        return int(self.lib.ws_launch_count(self.h))

    def last_timing(self, which):
        ms = C.c_float()
        self._check(self.lib.ws_last_timing(self.h, which, C.byref(ms)), "last_timing")
        return ms.value

    def is_finite(self):
        f = C.c_int32()
        self._check(self.lib.ws_is_finite(self.h, C.byref(f)), "is_finite")
        return bool(f.value)

    def comm_init(self, id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        self._check(self.lib.ws_comm_init(self.h, buf), "comm_init")

    @classmethod
    def comm_unique_id(cls):
        lib = cls._ensure_lib()
        buf = (C.c_char * 128)()
        rc = lib.ws_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("ws_comm_unique_id failed")
        return bytes(buf)


def build_emu(force=False):
    """TEST INFRASTRUCTURE: host emulation build (-DWS_EMULATE) of the C ABI and the general kernels, tests/emu/."""
    src = os.path.join(ROOT, "wave-simulation_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".hpp", ".cpp")))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "tests", "emu", "cuda_emu.hpp")),
                 os.path.getmtime(os.path.join(ROOT, "include", "wavesim.h")))
    if force or not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "tests", "emu")])
    return EMU_SO


class EmuSolver(Solver):
    """Same C ABI, host emulation build: checks kernel LOGIC on the CPU-only box. Never used by product or GPU tests."""
    lib = None
    so_path = EMU_SO

    @classmethod
    def _ensure_lib(cls):
        if cls.lib is None:
            build_emu()
            cls.lib = _load_ws_lib(cls.so_path)
        return cls.lib


# ---------------------------------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------------------------------
def idx1d(x, y, z, nx, nz):
    """Acquisition/Coordinates.cpp:687."""
    return x + z * nx + y * nx * nz


def two_layer(nx, ny, nz, depth=40, visco=False):
    """Tools/CreateModel/TwoLayer.cpp:25-62 (vp 3500/4550, vs 2000/2600, rho 2000/2600, tau 0.1; interface at y=depth)."""
    shape = (ny, nz, nx)
    vp = np.full(shape, 3500.0, np.float32)
    vs = np.full(shape, 2000.0, np.float32)
    rho = np.full(shape, 2000.0, np.float32)
    vp[depth:], vs[depth:], rho[depth:] = 4550.0, 2600.0, 2600.0
    m = dict(velocityP=vp.ravel(), velocityS=vs.ravel(), density=rho.ravel())
    if visco:
        m["tauP"] = np.full(vp.size, 0.1, np.float32)
        m["tauS"] = np.full(vp.size, 0.1, np.float32)
    return m


def homogeneous(n, vp=3500.0, vs=2000.0, rho=2000.0):
    return dict(velocityP=np.full(n, vp, np.float32), velocityS=np.full(n, vs, np.float32),
                density=np.full(n, rho, np.float32))


class Case:
    """A fully specified modelling case: descriptor + model + acquisition."""

    def __init__(self, name, desc, materials, src, rec, golden=None):
        self.name, self.desc, self.materials, self.src, self.rec, self.golden = name, desc, materials, src, rec, golden

    def needed_materials(self):
        eq = self.desc.eq
        keys = {0: ["velocityP", "density"], 1: ["velocityP", "velocityS", "density"],
                2: ["velocityP", "velocityS", "density", "tauP", "tauS"], 3: ["velocityS", "density"],
                4: ["velocityS", "density", "tauS"]}.get(eq)
        if keys is None:
            keys = list(self.materials.keys())
        return keys

    def setup(self, solver):
        for k in self.needed_materials():
            solver.set_material(k, self.materials[k])
        solver.prepare()
        st, si, sg = self.src
        solver.set_sources(st, si, sg)
        rt, ri = self.rec
        solver.set_receivers(rt, ri)
        solver.reset()
        return solver


def ci_case(name, nt=None):
    """The regular-grid CI cases of the reference (par/ci/configuration_ci.<name>.txt + sources/receiver files).
    DH 50, DT 2 ms, T 2 s -> NT 1000, Ricker fc 5 Hz amp 5 tshift 0."""
    NT = 1000 if nt is None else nt
    src_sig = ricker(1000, 2e-3, 5.0, 5.0, 0.0)[None, :NT]
    common = dict(dh=50.0, dt=2e-3, nt=NT, boundary_width=9, damping_coeff=8.0, vmax_cpml=3500.0, fc_cpml=5.0, npower=4.0)
    if name == "2D.acoustic":
        d = make_desc(2, "acoustic", 100, 100, fd_order=2, edge_policy=0, **dict(common, boundary_width=10))
        m = homogeneous(100 * 100)
        src = ([TYPE["P"]], [idx1d(49, 49, 0, 100, 1)], src_sig)
        rec = ([TYPE["P"]], [idx1d(69, 69, 0, 100, 1)])
        g = "seismogram.2D.acoustic.ref.p.mtx"
    elif name == "2D.sh":
        d = make_desc(2, "sh", 100, 100, fd_order=2, edge_policy=1, **dict(common, boundary_width=10))
        m = homogeneous(100 * 100)
        src = ([TYPE["VZ"]], [idx1d(49, 49, 0, 100, 1)], src_sig)
        rec = ([TYPE["VZ"]], [idx1d(69, 69, 0, 100, 1)])
        g = "seismogram.2D.sh.ref.vz.mtx"
    elif name == "2D.elastic":
        d = make_desc(2, "elastic", 100, 100, fd_order=12, edge_policy=1, free_surface=1, damping=1, **common)
        m = two_layer(100, 100, 1)
        src = ([TYPE["VX"]], [idx1d(20, 0, 0, 100, 1)], src_sig)
        rec = ([TYPE["VY"]], [idx1d(30, 0, 0, 100, 1)])
        g = "seismogram.2D.elastic.ref.vy.mtx"
    elif name == "3D.acoustic":
        d = make_desc(3, "acoustic", 100, 100, 100, fd_order=4, edge_policy=1, damping=1, **common)
        m = two_layer(100, 100, 100)
        src = ([TYPE["VX"]], [idx1d(20, 20, 20, 100, 100)], src_sig)
        rec = ([TYPE["P"]], [idx1d(30, 20, 30, 100, 100)])
        g = "seismogram.3D.acoustic.ref.p.mtx"
    elif name == "3D.elastic":
        d = make_desc(3, "elastic", 100, 100, 100, fd_order=2, edge_policy=0, free_surface=1, damping=1, **common)
        m = two_layer(100, 100, 100)
        src = ([TYPE["VX"]], [idx1d(20, 0, 20, 100, 100)], src_sig)
        rec = ([TYPE["VY"]], [idx1d(30, 0, 30, 100, 100)])
        g = "seismogram.3D.elastic.ref.vy.mtx"
    elif name == "2D.visco":  # stale config (equationType=visco) -> soft pin, run as viscoelastic
        d = make_desc(2, "viscoelastic", 100, 100, fd_order=8, edge_policy=1, free_surface=1, damping=1,
                      relax_freq=(5.0,), **common)
        m = two_layer(100, 100, 1, visco=True)
        src = ([TYPE["VX"]], [idx1d(20, 0, 0, 100, 1)], src_sig)
        rec = ([TYPE["VY"]], [idx1d(30, 0, 0, 100, 1)])
        g = "seismogram.2D.visco.ref.vy.mtx"
    elif name == "3D.visco":  # stale config -> soft pin
        d = make_desc(3, "viscoelastic", 100, 100, 100, fd_order=2, edge_policy=1, free_surface=1, damping=1,
                      relax_freq=(5.0,), **common)
        m = two_layer(100, 100, 100, visco=True)
        src = ([TYPE["VX"]], [idx1d(20, 0, 20, 100, 100)], src_sig)
        rec = ([TYPE["VY"]], [idx1d(30, 0, 30, 100, 100)])
        g = "seismogram.3D.visco.ref.vy.mtx"
    else:
        raise KeyError(name)
    return Case(name, d, m, src, rec, g)


def read_mtx_array(path):
    """MatrixMarket 'array' format written by lama::DenseMatrix::writeToFile: 2 header lines, column-major values."""
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("%")]
    rows, cols = (int(v) for v in lines[0].split()[:2])
    vals = np.array([float(v) for v in lines[1:1 + rows * cols]], dtype=np.float64)
    return vals.reshape((cols, rows)).T.copy()


def golden(name):
    return read_mtx_array(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a - b))


def reference_gate(test, ref):
    """Test_CompareSeismogram.cpp:57-91: sum of L2 misfits / (max * nSamples * nTraces) must be <= 5e-7
    (evaluated here for one component)."""
    test = np.asarray(test, np.float64)
    ref = np.asarray(ref, np.float64)
    misfit = np.linalg.norm(test - ref)
    return misfit / (np.abs(ref).max() * ref.shape[1] * ref.shape[0])

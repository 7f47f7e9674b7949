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


def load_pkg():
    """Import the hyphenated package directory wave-simulation_b200/ under the module name wave_simulation_b200."""
    import importlib.util
    import sys
    name = "wave_simulation_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "wave-simulation_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = load_pkg()
Desc, make_desc, Solver, SolverBase = _pkg.Desc, _pkg.make_desc, _pkg.Solver, _pkg.SolverBase
EQ, TYPE, idx1d, ricker_np, PRODUCT_SO = _pkg.EQ, _pkg.TYPE, _pkg.idx1d, _pkg.ricker_np, _pkg.PRODUCT_SO
_f32, _i32, _fp, _ip, _load_ws_lib = _pkg._f32, _pkg._i32, _pkg._fp, _pkg._ip, _pkg._load_ws_lib
ORACLE_SO = os.path.join(ROOT, "oracle", "_ref", "libwave_oracle.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libwavesim_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")




def build_oracle(force=False):
    """Compile oracle/ (plain g++, recipe = oracle/Makefile). Building the checker is not using it."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ROOT, "oracle", "wave_oracle.cpp")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


class Oracle(SolverBase):
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

    def set_fused(self, on=True):
        """the fused matrix-free back-end of the 3-D elastic solver (second CPU baseline of bench.py); bit-identical to the matrix formulation"""
        self._check(self.lib.wso_set_fused(self.h, 1 if on else 0), "set_fused")

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


class OracleVarGrid(Oracle):
    """Oracle on a variable grid / with variable FD orders (gridConfig columns: interface, dhFactor, FDorder)."""

    def __init__(self, desc, interfaces, dh_factors, fd_orders=None, precision=32):
        Oracle.num_threads()  # loads the library
        self.desc = desc
        self.h = C.c_void_p()
        self.n_rec = 0
        n = len(interfaces)
        ii, dd = _i32(interfaces), _i32(dh_factors)
        oo = _i32(fd_orders) if fd_orders is not None else None
        self._check(self.lib.wso_create_vargrid(C.byref(desc), precision, n, _ip(ii), _ip(dd), _ip(oo) if oo is not None else None, C.byref(self.h)), "create_vargrid")
        v = [C.c_int32() for _ in range(4)]
        self._check(self.lib.wso_grid_size(self.h, *[C.byref(x) for x in v]), "grid_size")
        self.nx, self.ny, self.nz, self.n_local = (x.value for x in v)

    def index(self, x, y, z=0):
        out = C.c_int32()
        self._check(self.lib.wso_coordinate2index(self.h, int(x), int(y), int(z), C.byref(out)), "coordinate2index")
        return out.value


def ricker(nt, dt, fc, amp, tshift=0.0):
    """Acquisition/SourceSignal/Ricker.cpp:29-53 evaluated by the oracle library in float."""
    build_oracle()
    if Oracle.lib is None:
        Oracle.lib = C.CDLL(ORACLE_SO)
    out = np.zeros(nt, dtype=np.float32)
    rc = Oracle.lib.wso_wavelet(1, nt, C.c_float(dt), C.c_float(fc), C.c_float(amp), C.c_float(tshift), _fp(out))
    assert rc == 0
    return out


def build_emu(force=False):
    """TEST INFRASTRUCTURE: host emulation build (-DWS_EMULATE) of the C ABI and the general kernels, tests/emu/."""
    src = os.path.join(ROOT, "wave-simulation_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".hpp", ".cpp")))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, "tests", "emu", "cuda_emu.hpp")),
                 os.path.getmtime(os.path.join(ROOT, "include", "wavesim.h")))
    if force or not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.check_call(["make", "-s", "-j", str(min(8, os.cpu_count() or 2)), "-C", os.path.join(ROOT, "tests", "emu")])
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

    SENDRECV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_size_t, C.c_int)

    def comm_init_external(self, sendrecv):
        """sendrecv(send: np.ndarray, recv: np.ndarray, peer: int) exchanges ghost planes with a neighbour rank."""
        def thunk(_user, send, recv, count, peer):
            try:
                sendrecv(np.ctypeslib.as_array(send, shape=(count,)), np.ctypeslib.as_array(recv, shape=(count,)), peer)
                return 0
            except Exception as exc:  # surfaces as WS_ECOMM
                print("sendrecv failed:", exc, flush=True)
                return 1
        self._cb = EmuSolver.SENDRECV(thunk)  # keep alive
        self._check(self.lib.ws_comm_init_external(self.h, self._cb, None), "comm_init_external")


# ---------------------------------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------------------------------
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

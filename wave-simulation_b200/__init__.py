"""wave-simulation_b200 — Python ctypes binding of the C ABI (include/wavesim.h) of the B200-native FD time-stepping
library.  This is plumbing for tests/ and bench.py; the reference-facing host layer is C++ (wave-simulation_b200/host/).

There is NO CPU path here: `Solver` loads csrc/libwavesim_cuda.so and fails loudly if it is missing or if no CUDA device
is present.  Nothing in this package imports or loads oracle/.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
# WAVESIM_LIB: developer override used to A/B kernel build variants on the GPU box
PRODUCT_SO = os.environ.get("WAVESIM_LIB") or os.path.join(PKG_DIR, "csrc", "libwavesim_cuda.so")

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


class SolverBase:
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


def _load_ws_lib(path):
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.ws_destroy.argtypes = [C.c_void_p]
    lib.ws_launch_count.restype = C.c_uint64
    lib.ws_launch_count.argtypes = [C.c_void_p]
    lib.ws_estimate_memory.restype = C.c_size_t
    lib.ws_stream.restype = C.c_void_p
    lib.ws_stream.argtypes = [C.c_void_p]
    lib.ws_uses_fast_kernels.argtypes = [C.c_void_p]
    lib.ws_kernel_path.argtypes = [C.c_void_p]
    if hasattr(lib, "ws_halo_transport"):
        lib.ws_halo_transport.argtypes = [C.c_void_p]
    if hasattr(lib, "ws_wavefields_create"):
        lib.ws_wavefields_destroy.argtypes = [C.c_void_p]
        lib.ws_wavefields_destroy.restype = None
        for name in ("assign", "plus_assign", "minus_assign"):
            getattr(lib, "ws_wavefields_" + name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ws_wavefields_times_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        lib.ws_wavefields_times_assign_vector.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_size_t]
        lib.ws_wavefields_get.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_size_t]
        lib.ws_set_step_scaling.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_size_t]
    return lib


class Solver(SolverBase):
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

    def set_sources64(self, types, idx, signals):
        t, i, sg = _i32(types), np.ascontiguousarray(idx, dtype=np.int64), _f32(signals)
        assert sg.shape == (len(t), self.desc.nt), sg.shape
        self._check(self.lib.ws_set_sources64(self.h, len(t), _ip(t), i.ctypes.data_as(C.POINTER(C.c_int64)), _fp(sg)), "set_sources64")

    def set_receivers64(self, types, idx):
        t, i = _i32(types), np.ascontiguousarray(idx, dtype=np.int64)
        self.n_rec = len(t)
        self._check(self.lib.ws_set_receivers64(self.h, len(t), _ip(t), i.ctypes.data_as(C.POINTER(C.c_int64))), "set_receivers64")

    def stream_ptr(self):
        return int(self.lib.ws_stream(self.h))

    def sync(self):
        self._check(self.lib.ws_sync(self.h), "sync")

    def step_host(self, t, src_samples, rec_samples):
        sp = _fp(src_samples) if src_samples is not None else None
        self._check(self.lib.ws_step_host(self.h, t, sp, _fp(rec_samples)), "step_host")

    def set_timing(self, enable):
        self._check(self.lib.ws_set_timing(self.h, int(enable)), "set_timing")

    def uses_fast_kernels(self):
        return bool(self.lib.ws_uses_fast_kernels(self.h))

    def kernel_path(self):
        """0 per-point kernels, 1 marching kernels, 2 warp-specialised TMA kernels"""
        return int(self.lib.ws_kernel_path(self.h))

    def launch_count(self):
        return int(self.lib.ws_launch_count(self.h))

    def estimate_memory(self):
        """bytes of HBM this rank allocates (ForwardSolver::estimateMemory)"""
        return int(self.lib.ws_estimate_memory(C.byref(self.desc)))

    def last_timing(self, which):
        ms = C.c_float()
        self._check(self.lib.ws_last_timing(self.h, which, C.byref(ms)), "last_timing")
        return ms.value

    # wavefield objects (Wavefields/Wavefields.hpp:62-80); None = the solver's own wavefields
    def wavefields_create(self):
        w = C.c_void_p()
        self._check(self.lib.ws_wavefields_create(self.h, C.byref(w)), "wavefields_create")
        return w

    def wavefields_destroy(self, w):
        self.lib.ws_wavefields_destroy(w)

    def wavefields_assign(self, dst, src):
        self._check(self.lib.ws_wavefields_assign(self.h, dst, src), "wavefields_assign")

    def wavefields_plus_assign(self, dst, src):
        self._check(self.lib.ws_wavefields_plus_assign(self.h, dst, src), "wavefields_plus_assign")

    def wavefields_minus_assign(self, dst, src):
        self._check(self.lib.ws_wavefields_minus_assign(self.h, dst, src), "wavefields_minus_assign")

    def wavefields_times_assign(self, dst, rhs):
        if np.ndim(rhs) == 0:
            self._check(self.lib.ws_wavefields_times_assign(self.h, dst, C.c_float(float(rhs))), "wavefields_times_assign")
        else:
            a = _f32(rhs).ravel()
            self._check(self.lib.ws_wavefields_times_assign_vector(self.h, dst, _fp(a), C.c_size_t(a.size)), "wavefields_times_assign_vector")

    def wavefields_get(self, w, comp):
        out = np.empty(self.n_local, dtype=np.float32)
        self._check(self.lib.ws_wavefields_get(self.h, w, comp.encode(), _fp(out), C.c_size_t(out.size)), "wavefields_get")
        return out

    def set_step_scaling(self, vec):
        if vec is None:
            self._check(self.lib.ws_set_step_scaling(self.h, None, C.c_size_t(0)), "set_step_scaling")
        else:
            a = _f32(vec).ravel()
            self._check(self.lib.ws_set_step_scaling(self.h, _fp(a), C.c_size_t(a.size)), "set_step_scaling")

    def is_finite(self):
        f = C.c_int32()
        self._check(self.lib.ws_is_finite(self.h, C.byref(f)), "is_finite")
        return bool(f.value)

    def halo_transport(self):
        """0 single rank, 1 NCCL send / recv, 2 external transport, 3 the library's kernels over peer memory"""
        return int(self.lib.ws_halo_transport(self.h))

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


def idx1d(x, y, z, nx, nz):
    """Acquisition/Coordinates.cpp:687."""
    return x + z * nx + y * nx * nz

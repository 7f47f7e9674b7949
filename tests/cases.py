"""Seeded synthetic cases shared by the CPU (emulation) and GPU parity tests."""
import numpy as np

from wsharness import Case, idx1d, make_desc, ricker_np

EPS0 = 8.8541878176e-12
MU0 = 1.2566370614e-6


def rand_model(rng, n, visco=False):
    vp = (3000 + 1000 * rng.random(n)).astype(np.float32)
    vs = (vp / np.sqrt(3) * (0.9 + 0.2 * rng.random(n))).astype(np.float32)
    rho = (2000 + 500 * rng.random(n)).astype(np.float32)
    m = dict(velocityP=vp, velocityS=vs, density=rho)
    if visco:
        m["tauP"] = (0.05 + 0.1 * rng.random(n)).astype(np.float32)
        m["tauS"] = (0.05 + 0.1 * rng.random(n)).astype(np.float32)
    return m


def em_model(rng, n):
    return dict(dielectricPermittivity=(EPS0 * (4 + 2 * rng.random(n))).astype(np.float32),
                electricConductivity=(1e-3 * (1 + rng.random(n))).astype(np.float32),
                magneticPermeability=(MU0 * (1 + 0.1 * rng.random(n))).astype(np.float32),
                tauDielectricPermittivity=(0.05 + 0.05 * rng.random(n)).astype(np.float32),
                tauElectricConductivity=(1e-10 * rng.random(n)).astype(np.float32))


def seis_fields(dim, L):
    f = ["VX", "VY", "Sxx", "Syy", "Sxy"] + (["VZ", "Szz", "Sxz", "Syz"] if dim == 3 else [])
    for l in range(1, L + 1):
        f += ["Rxx%d" % l, "Ryy%d" % l, "Rxy%d" % l] + (["Rzz%d" % l, "Rxz%d" % l, "Ryz%d" % l] if dim == 3 else [])
    return f


def fields_of(eq, dim, L):
    if eq == "acoustic":
        return ["VX", "VY", "P"] + (["VZ"] if dim == 3 else [])
    if eq in ("elastic", "viscoelastic"):
        return seis_fields(dim, L)
    if eq in ("sh", "viscosh"):
        return ["VZ", "Sxz", "Syz"] + ["Rxz%d" % l for l in range(1, L + 1)] + ["Ryz%d" % l for l in range(1, L + 1)]
    if eq in ("tmem", "viscotmem"):
        return ["HX", "HY", "EZ"] + ["RZ%d" % l for l in range(1, L + 1)]
    f = ["HZ", "EX", "EY"] + (["HX", "HY", "EZ"] if dim == 3 else [])
    for l in range(1, L + 1):
        f += ["RX%d" % l, "RY%d" % l] + (["RZ%d" % l] if dim == 3 else [])
    return f


def make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W=6, L=0, nt=40, exact=1, seed=20260101, kernel_variant=0):
    """Random heterogeneous model, 3 sources (one on the surface/edge), 4 receivers incl. corners."""
    rng = np.random.default_rng(seed)
    nzz = nz if dim == 3 else 1
    n = nx * ny * nzz
    em = eq in ("tmem", "emem", "viscotmem", "viscoemem")
    if em:
        d = make_desc(dim, eq, nx, ny, nz, dh=0.02, dt=2e-11, nt=nt, fd_order=q, edge_policy=pol, free_surface=fs, damping=damp,
                      boundary_width=W, vmax_cpml=3e8, fc_cpml=1e8, relax_freq=tuple([1e8, 3e8, 5e7, 2e8][:L]),
                      exact_arith=exact, kernel_variant=kernel_variant)
        m, fc, amp = em_model(rng, n), 1e8, 1.0
    else:
        d = make_desc(dim, eq, nx, ny, nz, dh=10.0, dt=8e-4, nt=nt, fd_order=q, edge_policy=pol, free_surface=fs, damping=damp,
                      boundary_width=W, vmax_cpml=4000., fc_cpml=20., relax_freq=tuple([20., 50., 5., 80.][:L]),
                      exact_arith=exact, kernel_variant=kernel_variant)
        m, fc, amp = rand_model(rng, n, visco=L > 0), 25., 1e3
    sig = np.stack([ricker_np(nt, d.dt, fc, amp, 0.0) * (1 + k) for k in range(3)])
    if eq == "acoustic":
        st, rt = [1, 2, 3], [1, 2, 3, 1]
    elif eq in ("elastic", "viscoelastic"):
        st, rt = [1, 2, 3], [1, 2, 3, (4 if dim == 3 else 2)]
    elif eq in ("sh", "viscosh"):
        st, rt = [4, 4, 4], [4, 4, 4, 4]
    elif eq in ("tmem", "viscotmem"):
        st, rt = [1, 1, 1], [1, 1, 1, 1]
    else:
        st, rt = [2, 3, 4], [2, 3, 4, (1 if dim == 3 else 2)]
    sidx = [idx1d(nx // 2, 0 if fs else ny // 2, nzz // 2, nx, nzz), idx1d(nx // 3, ny // 3, nzz // 3, nx, nzz),
            idx1d(2, 2, min(2, nzz - 1), nx, nzz)]
    ridx = [idx1d(nx // 2 + 2, 0, nzz // 2, nx, nzz), idx1d(nx - 2, ny - 2, nzz - 1, nx, nzz), idx1d(1, 1, 0, nx, nzz),
            idx1d(nx // 2, ny // 2, nzz // 2, nx, nzz)]
    return Case("%s%dD" % (eq, dim), d, m, (st, sidx, sig), (rt, ridx))


# (eq, dim, nx, ny, nz, q, edge_policy, free_surface, damping, W, L)
SWEEP = [
    ("acoustic", 3, 20, 22, 18, 4, 0, 1, 2, 6, 0), ("acoustic", 3, 20, 22, 18, 8, 1, 0, 1, 6, 0), ("acoustic", 2, 40, 36, 1, 6, 1, 1, 2, 8, 0),
    ("elastic", 3, 20, 22, 18, 8, 0, 1, 2, 6, 0), ("elastic", 3, 20, 22, 18, 4, 1, 1, 1, 6, 0), ("elastic", 3, 26, 26, 26, 12, 1, 0, 2, 5, 0),
    ("elastic", 2, 40, 36, 1, 8, 0, 1, 2, 8, 0), ("elastic", 2, 40, 36, 1, 10, 1, 0, 0, 8, 0),
    ("viscoelastic", 3, 20, 22, 18, 4, 0, 1, 2, 6, 2), ("viscoelastic", 2, 40, 36, 1, 8, 1, 1, 2, 8, 3), ("viscoelastic", 3, 20, 22, 18, 2, 1, 0, 1, 6, 1),
    ("sh", 2, 40, 36, 1, 8, 1, 1, 2, 8, 0), ("viscosh", 2, 40, 36, 1, 4, 0, 1, 2, 8, 2), ("viscosh", 2, 40, 36, 1, 4, 1, 0, 1, 8, 1),
    ("tmem", 2, 40, 36, 1, 8, 0, 0, 2, 8, 0), ("viscotmem", 2, 40, 36, 1, 4, 1, 0, 2, 8, 2), ("tmem", 2, 40, 36, 1, 2, 1, 0, 1, 8, 0),
    ("emem", 2, 40, 36, 1, 8, 0, 0, 2, 8, 0), ("viscoemem", 2, 40, 36, 1, 4, 1, 0, 2, 8, 2), ("emem", 3, 20, 22, 18, 4, 0, 0, 2, 6, 0),
    ("viscoemem", 3, 20, 22, 18, 8, 1, 0, 2, 6, 2), ("emem", 3, 20, 22, 18, 2, 1, 0, 1, 6, 0),
    # FreeSurface = 2 (improved vacuum formulation: no image method, no absorbing frame at the top; ABS*/CPML*::init test useFreeSurface == 0)
    ("elastic", 2, 40, 36, 1, 8, 1, 2, 2, 8, 0), ("elastic", 3, 20, 22, 18, 4, 0, 2, 1, 6, 0), ("acoustic", 2, 40, 36, 1, 4, 0, 2, 2, 8, 0),
    ("viscoelastic", 3, 20, 22, 18, 4, 0, 2, 2, 6, 1),
]


def sweep_id(c):
    return "%s%dD-q%d-pol%d-fs%d-damp%d-L%d" % (c[0], c[1], c[5], c[6], c[7], c[8], c[10])

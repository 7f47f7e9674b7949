"""Wavefield objects and their operators (Wavefields/Wavefields.hpp:62-80, Simulation.cpp:450-461): a copy
of the wavefields before a step, -=, +=, *= scalar, *= vector on whole objects, and `*wavefields *= compensation` inside the
time loop.  Every operator is one fp32 rounding per element, so the results must equal numpy's float32 results bit for bit.
The same tests run on the host emulation build (CPU suite) and on the CUDA library (`-m gpu`)."""
import numpy as np
import pytest

from cases import fields_of, make_case
from wsharness import EmuSolver, Oracle, Solver, rel_l2

BACKENDS = [pytest.param(EmuSolver, id="emulation"), pytest.param(Solver, id="cuda", marks=pytest.mark.gpu)]
CASES = [("viscoelastic", 3, 14, 12, 10, 4, 0, 1, 2, 4, 2), ("viscotmem", 2, 40, 30, 1, 8, 0, 0, 2, 8, 1), ("acoustic", 2, 36, 28, 1, 4, 1, 1, 1, 6, 0)]


def snapshot(s, w, fields):
    return {f: s.wavefields_get(w, f) for f in fields}


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("cfg", CASES, ids=[c[0] + "%dD" % c[1] for c in CASES])
def test_operators_equal_numpy(backend, cfg):
    eq, dim, nx, ny, nz, q, pol, fs, damp, W, L = cfg
    case = make_case(eq, dim, nx, ny, nz, q, pol, fs, damp, W, L, nt=12, exact=0, kernel_variant=1)
    s = case.setup(backend(case.desc))
    fields = fields_of(eq, dim, L)
    tmp, acc = s.wavefields_create(), s.wavefields_create()
    assert all(not snapshot(s, tmp, fields)[f].any() for f in fields)  # Wavefields::init: zero
    s.run(0, 6)
    a = snapshot(s, None, fields)
    s.wavefields_assign(tmp, None)  # *wavefieldsTemp = *wavefields (Simulation.cpp:450)
    s.run(6, 12)
    b = snapshot(s, None, fields)
    assert any(np.abs(a[f]).max() > 0 for f in fields) and any(not np.array_equal(a[f], b[f]) for f in fields)
    got = snapshot(s, tmp, fields)
    for f in fields:
        assert np.array_equal(got[f], a[f]), f
    s.wavefields_minus_assign(tmp, None)  # *wavefieldsTemp -= *wavefields
    dtinv = np.float32(-1.0 / case.desc.dt)
    s.wavefields_times_assign(tmp, dtinv)  # *wavefieldsTemp *= -DTinv (:459)
    got = snapshot(s, tmp, fields)
    for f in fields:
        assert np.array_equal(got[f], (a[f] - b[f]) * dtinv), f
    s.wavefields_plus_assign(acc, tmp)
    s.wavefields_plus_assign(acc, None)
    got = snapshot(s, acc, fields)
    for f in fields:
        assert np.array_equal(got[f], ((a[f] - b[f]) * dtinv) + b[f]), f
    rng = np.random.default_rng(7)
    vec = (0.5 + rng.random(s.n_local)).astype(np.float32)
    s.wavefields_times_assign(None, vec)  # live wavefields *= vector
    got = snapshot(s, None, fields)
    for f in fields:
        assert np.array_equal(got[f], b[f] * vec), f
    s.wavefields_assign(None, acc)  # the live state takes a stored object (checkpoint restore)
    got = snapshot(s, None, fields)
    for f in fields:
        assert np.array_equal(got[f], ((a[f] - b[f]) * dtinv) + b[f]), f
    with pytest.raises(RuntimeError):
        s.wavefields_assign(tmp, tmp)
    s.wavefields_destroy(tmp)
    s.wavefields_destroy(acc)
    s.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_checkpoint_restore_reproduces_the_run(backend):
    """A stored copy put back into the solver continues the run bit for bit (what an adjoint-state checkpoint relies on)."""
    case = make_case("elastic", 2, 40, 36, 1, 8, 0, 1, 0, 6, 0, nt=20, exact=0, kernel_variant=1)
    s = case.setup(backend(case.desc))
    ck = s.wavefields_create()
    s.run(0, 10)
    s.wavefields_assign(ck, None)
    s.run(10, 20)
    want = {f: s.wavefield(f) for f in fields_of("elastic", 2, 0)}
    s.wavefields_assign(None, ck)
    s.run(10, 20)
    for f, v in want.items():
        assert np.array_equal(s.wavefield(f), v), f
    s.wavefields_destroy(ck)
    s.close()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("variant", [1, 0])
def test_step_scaling_equals_stepwise_multiplication(backend, variant):
    """`*wavefields *= compensation` after every step (Simulation.cpp:455-456) inside ws_run (graph-batched on the GPU) against
    the oracle stepped one step at a time with the multiplication done in numpy in between."""
    eq, dim, nx, ny, L = "viscotmem", 2, 44, 32, 1
    case = make_case(eq, dim, nx, ny, 1, 4, 0, 0, 2, 8, L, nt=24, exact=0, kernel_variant=variant)
    fields = fields_of(eq, dim, L)
    sig = case.materials["electricConductivity"] / case.materials["dielectricPermittivity"]
    comp = np.exp((sig * np.float32(case.desc.dt)).astype(np.float32)).astype(np.float32)  # Modelparameter.cpp:128-141, tStep = 1
    comp = (comp * np.float32(1.0005)).astype(np.float32)  # (sigma dt / eps is ~1e-3 here: make the factor matter)
    o = case.setup(Oracle(case.desc))
    for t in range(24):
        o.run(t, t + 1)
        for f in fields:
            o.set_wavefield(f, o.wavefield(f) * comp)
    s = case.setup(backend(case.desc))
    s.set_step_scaling(comp)
    s.run(0, 24)
    if hasattr(s, "sync"):
        s.sync()
    assert rel_l2(s.seismogram(), o.seismogram()) <= 1.0e-5
    for f in fields:
        a, b = o.wavefield(f), s.wavefield(f)
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(a).max(), 1e-30), f
    # switching it off again gives the plain run
    s.set_step_scaling(None)
    s.reset()
    s.run(0, 24)
    p = case.setup(backend(case.desc))
    p.run(0, 24)
    assert np.array_equal(s.seismogram(), p.seismogram())
    s.close()
    p.close()
    o.close()

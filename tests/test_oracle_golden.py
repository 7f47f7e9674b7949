"""The oracle is pinned against the reference's own golden seismograms (par/ci/seismogram.*.ref.*.mtx, copied to
tests/golden/).  Finding recorded in DESIGN.md: all 7 regular-grid goldens are reproduced (to the 6 significant digits
the .mtx files carry) by the *truncating* edge policy (LAMA StencilMatrix behaviour, edge_policy=0), including the two
whose configs name sparse matrices — the goldens predate the order-reducing sparse assembly of Derivatives.cpp:166-175."""
import os

import numpy as np
import pytest

from wsharness import Oracle, ci_case, golden, reference_gate, rel_l2

FULL = os.environ.get("WS_FULL_GOLDEN", "0") == "1"
# (case, number of leading samples checked by default; the time stepping is causal so a prefix is a valid check)
CASES = [("2D.acoustic", 1000), ("2D.sh", 1000), ("2D.elastic", 1000), ("2D.visco", 1000),
         ("3D.acoustic", 450), ("3D.elastic", 450), ("3D.visco", 400)]


@pytest.mark.parametrize("name,nt", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_reference_golden(name, nt):
    nt = 1000 if FULL else nt
    case = ci_case(name, nt=nt)
    case.desc.edge_policy = 0
    o = case.setup(Oracle(case.desc))
    o.run(0, nt)
    s = o.seismogram()
    g = golden(case.golden)[:, :nt]
    assert np.abs(g).max() > 0
    # goldens carry 6 significant digits -> relative L2 floors at ~2e-6
    assert rel_l2(s, g) <= 1.0e-5, rel_l2(s, g)
    # the reference's own CI gate, Test_CompareSeismogram.cpp:84
    assert reference_gate(s, g) <= 5.0e-7


VARGRID = dict(interfaces=[0, 30, 150, 200], dh_factors=[1, 3, 1, 3], fd_orders=[2, 6, 2, 6])  # par/ci/gridConfig.txt


def vargrid_ci_case(dim, nt, edge_policy=0, damping=2, free_surface=1):
    """par/ci/configuration_ci.{2D,3D}.acoustic.varGrid.txt: variable grid (dhFactor 1/3/1/3) with variable FD order (2/6/2/6),
    free surface + CPML(30), DH 17, homogeneous vp 3500 / rho 2000, P source and four P receivers at increasing depth."""
    from wsharness import OracleVarGrid, make_desc, ricker
    common = dict(dh=17.0, dt=2e-3, nt=nt, fd_order=2, edge_policy=edge_policy, free_surface=free_surface, damping=damping, boundary_width=30, damping_coeff=8.0,
                  vmax_cpml=3500.0, fc_cpml=5.0, npower=4.0)
    if dim == 2:
        d = make_desc(2, "acoustic", 305, 303, 1, **common)
        src, recs, g = (150, 20, 0), [(150, 20, 0), (150, 90, 0), (150, 170, 0), (150, 239, 0)], "seismogram.2D.acoustic.varGrid.ref.p.mtx"
    else:
        d = make_desc(3, "acoustic", 104, 303, 104, **common)
        src, recs, g = (50, 20, 50), [(50, 20, 50), (51, 90, 51), (50, 170, 50), (51, 239, 51)], "seismogram.3D.acoustic.varGrid.ref.p.mtx"
    o = OracleVarGrid(d, VARGRID["interfaces"], VARGRID["dh_factors"], VARGRID["fd_orders"])
    o.set_material("velocityP", np.full(o.n_local, 3500.0, np.float32))
    o.set_material("density", np.full(o.n_local, 2000.0, np.float32))
    o.prepare()
    o.set_sources([1], [o.index(*src)], ricker(1000, 2e-3, 5.0, 5.0, 0.0)[None, :nt])
    o.set_receivers([1] * 4, [o.index(*r) for r in recs])
    o.reset()
    return o, g


@pytest.mark.parametrize("dim,nt", [(2, 1000), (3, 1000)], ids=["2D", "3D"])
def test_oracle_reproduces_variable_grid_goldens(dim, nt):
    """The two variable-grid goldens are the only reference fixtures that exercise CPML (and the acoustic free surface).  With
    the truncating edge policy the oracle passes the reference's own CI gate (Test_CompareSeismogram.cpp:84) on both and
    reproduces the 2-D traces to 3e-5 .. 6e-4 relative L2 each (the source-depth trace of the 3-D case to 6e-5).  What is
    NOT reproduced to the goldens' 6 digits: the transmission through the coarse -> fine interface (3-5e-4) and, in the
    3-D case, the deep traces of the 44-cell-wide tube between the CPML layers (3-6 %, within the gate because they are 200x
    weaker than the shallow trace): like the regular-grid goldens these files predate details of the current assembly
    (the literal order reduction of the calc* loops misses them by 7e-3, test below)."""
    nt = 1000 if FULL else nt
    o, gname = vargrid_ci_case(dim, nt)
    assert (o.nx, o.ny, o.nz) == ((305, 303, 1) if dim == 2 else (104, 303, 104))
    o.run(0, nt)
    s = o.seismogram()
    g = golden(gname)[:, :nt]
    assert reference_gate(s, g) <= 5.0e-7, reference_gate(s, g)
    per_trace = [rel_l2(s[k], g[k]) for k in range(4)]
    print("variable grid %dD: gate %.2e total rel L2 %.2e per trace %s" % (dim, reference_gate(s, g), rel_l2(s, g), ["%.1e" % v for v in per_trace]))
    assert per_trace[0] <= 1.0e-4
    if dim == 2:
        assert rel_l2(s, g) <= 1.0e-4 and max(per_trace) <= 8.0e-4
        # up to the arrival of the wave transmitted through the coarse -> fine interface the traces agree to the goldens' digits
        assert rel_l2(s[1][:450], g[1][:450]) <= 1.0e-5 and rel_l2(s[0][:700], g[0][:700]) <= 1.0e-5
    else:
        assert rel_l2(s, g) <= 5.0e-4 and max(per_trace) <= 8.0e-2


def test_variable_grid_literal_order_reduction_differs_from_golden():
    o, gname = vargrid_ci_case(2, 1000, edge_policy=1)
    o.run(0, 1000)
    assert rel_l2(o.seismogram(), golden(gname)) > 3.0e-3


def test_variable_grid_code_equals_regular_code_on_one_layer():
    """dhFactor 1 and one FD order everywhere: the layered assembly must give the regular sparse assembly bit for bit."""
    from wsharness import OracleVarGrid, idx1d, make_desc, ricker
    nt = 200
    res = []
    for var in (False, True):
        d = make_desc(2, "acoustic", 80, 70, 1, dh=17.0, dt=2e-3, nt=nt, fd_order=4, edge_policy=1, free_surface=1, damping=2, boundary_width=10, vmax_cpml=3500.0,
                      fc_cpml=5.0, npower=4.0)
        o = OracleVarGrid(d, [0, 20, 40], [1, 1, 1], [4, 4, 4]) if var else Oracle(d)
        n = 80 * 70
        o.set_material("velocityP", np.full(n, 3500.0, np.float32))
        o.set_material("density", np.full(n, 2000.0, np.float32))
        o.prepare()
        o.set_sources([1], [idx1d(40, 10, 0, 80, 1)], ricker(nt, 2e-3, 5.0, 5.0, 0.0)[None, :])
        o.set_receivers([1, 1], [idx1d(40, 30, 0, 80, 1), idx1d(10, 60, 0, 80, 1)])
        o.reset()
        o.run(0, nt)
        res.append(o.seismogram())
    assert np.abs(res[0]).max() > 0 and np.array_equal(res[0], res[1])


def test_order_reducing_policy_differs_from_golden():
    """Documents that the literal restatement of the sparse assembly (order reduction) is NOT what produced the goldens."""
    case = ci_case("2D.elastic", nt=600)
    case.desc.edge_policy = 1
    o = case.setup(Oracle(case.desc))
    o.run(0, 600)
    g = golden(case.golden)[:, :600]
    assert rel_l2(o.seismogram(), g) > 1.0e-3


def test_fp64_oracle_brackets_fp32():
    """fp32 oracle vs its own fp64 evaluation: the rounding noise floor that any fp32 implementation sits in."""
    case = ci_case("2D.elastic", nt=600)
    case.desc.edge_policy = 0
    o32 = case.setup(Oracle(case.desc, 32))
    o64 = case.setup(Oracle(case.desc, 64))
    o32.run(0, 600)
    o64.run(0, 600)
    assert rel_l2(o32.seismogram(), o64.seismogram()) < 1.0e-5


@pytest.mark.parametrize("pol,fs,damp", [(0, 1, 2), (1, 1, 2), (0, 0, 0), (1, 2, 2)])
def test_fused_backend_of_the_oracle_is_bit_identical(pol, fs, damp):
    """The second CPU baseline of bench.py: the 3-D elastic step with 1-D coefficient rows and all statements of a
    half-step fused per grid point performs the operations of the matrix formulation in the same order: identical bits."""
    from cases import fields_of, make_case
    res = []
    for fused in (False, True):
        case = make_case("elastic", 3, 26, 24, 22, 8, pol, fs, damp, 6, 0, nt=20, exact=1)
        o = case.setup(Oracle(case.desc))
        if fused:
            o.set_fused(True)
        o.run(0, 20)
        res.append((o.seismogram(), {f: o.wavefield(f) for f in fields_of("elastic", 3, 0)}))
        o.close()
    assert np.abs(res[0][0]).max() > 0
    assert np.array_equal(res[0][0], res[1][0])
    for f in res[0][1]:
        assert np.array_equal(res[0][1][f], res[1][1][f]), f

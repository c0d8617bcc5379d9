"""GPU parity tests of flow-field evaluation (C ABI ludvm_flowfield_*, replacing LUDVM.flowfield,
LUDVM.py:1186-1298) against the reference's golden output and the oracle."""
import numpy as np
import pytest

from conftest import biteq, golden_tables, load_golden

pytestmark = pytest.mark.gpu


def _sim(g):
    from ludvm_b200 import LUDVM
    kw = dict(g["kw"])
    kw["circulation_freevort"], kw["xy_freevort"] = np.array(kw["circulation_freevort"]), np.array(kw["xy_freevort"])
    tb = golden_tables(g)
    tb["sum_free"] = float(np.sum(tb["free_g"]))
    s = LUDVM(**kw, verbose=False, run=False)
    s.time_loop(tables=tb)
    return s


def test_flowfield_bit_equal_to_reference_golden():
    g = load_golden("freevort_tf3")
    s = _sim(g)
    s.flowfield(**g["ff_kw"])
    for k in ("x_ff", "z_ff", "u_ff", "w_ff", "ome_ff"):
        assert biteq(getattr(s, k), g[k]), k


def test_flowfield_fast_mode_and_row_slabs(oracle):
    """Fast mode within tolerance; a grid evaluated in x-row slabs (the multi-GPU decomposition) equals the
    whole-grid evaluation bitwise, and the vorticity stencil matches the oracle's on the assembled field."""
    from ludvm_b200 import ops
    rng = np.random.default_rng(20260102)
    n = 5000
    g, xw, zw = rng.standard_normal(n) * 1e-2, rng.uniform(-20, 0, n), rng.uniform(-4, 4, n)
    x1, z1 = np.arange(-3.0, -1.0, 0.02), np.arange(-1.0, 1.0, 0.02)
    vc4 = 0.065 ** 4
    u, w = ops.flowfield_velocity(g, xw, zw, None, None, None, vc4, x1, z1, mode="exact")
    X, Z = np.meshgrid(x1, z1, indexing="ij")
    uo, wo = oracle.induced_velocity(g, xw, zw, X.ravel(), Z.ravel(), 0.065)
    assert biteq(u.ravel(), uo) and biteq(w.ravel(), wo)
    uf, wf = ops.flowfield_velocity(g, xw, zw, None, None, None, vc4, x1, z1, mode="fast")
    assert np.max(np.abs(uf - u)) <= 1e-12 * np.max(np.abs(u))
    parts = [ops.flowfield_velocity(g, xw, zw, None, None, None, vc4, x1, z1, row0=r0, nrows=nr, mode="exact")
             for r0, nr in ((0, 37), (37, 40), (77, len(x1) - 77))]
    assert biteq(np.concatenate([p[0] for p in parts]), u) and biteq(np.concatenate([p[1] for p in parts]), w)
    ome = ops.flowfield_vorticity(x1, z1, u[None], w[None])
    assert biteq(ome, oracle.vorticity(X, Z, u[None], w[None]))


def test_flowfield_large_grid_properties():
    """A bigger case through size-independent properties: the velocity field of one vortex is solenoidal away from
    the core and its vorticity integrates to the circulation (Stokes, the check sketched at LUDVM.py:64-70)."""
    from ludvm_b200 import ops
    dr = 0.01
    x1, z1 = np.arange(-1.0, 1.0, dr), np.arange(-1.0, 1.0, dr)
    G = 2.5
    u, w = ops.flowfield_velocity(np.array([G]), np.array([0.003]), np.array([-0.004]), None, None, None, 0.05 ** 4,
                                  x1, z1, mode="fast")
    ome = ops.flowfield_vorticity(x1, z1, u[None], w[None])[0]
    circ = np.trapezoid(np.trapezoid(ome, dx=dr, axis=0), dx=dr, axis=0)
    # the reference's convention: positive circulation is clockwise, so omega = dw/dx - du/dz integrates to -Gamma
    assert abs(circ + G) < 0.05 * G


def test_flowfield_exact_large_grid_few_sources(oracle):
    """A README-size wake (300 vortices, three source segments' worth of ragged leaves) on a 240 x 240 grid: enough
    rows for the one-thread-per-row exact kernel even with a shallow summation tree.  Bit-equal to the oracle."""
    from ludvm_b200 import ops
    rng = np.random.default_rng(20260103)
    n = 300
    g, xw, zw = rng.standard_normal(n) * 1e-2, rng.uniform(-5, 0, n), rng.uniform(-1, 1, n)
    x1, z1 = np.arange(-6.0, 0.0, 0.025), np.arange(-3.0, 3.0, 0.025)
    assert len(x1) * len(z1) >= 148 * 128 * 3
    u, w = ops.flowfield_velocity(g, xw, zw, None, None, None, 0.065 ** 4, x1, z1, mode="exact")
    X, Z = np.meshgrid(x1, z1, indexing="ij")
    uo, wo = oracle.induced_velocity(g, xw, zw, X.ravel(), Z.ravel(), 0.065)
    assert biteq(u.ravel(), uo) and biteq(w.ravel(), wo)


def test_flowfield_method_row_slabs():
    """`LUDVM.flowfield(rows=(row0, nrows))`: slabs with one halo row per interior side reassemble the whole-grid
    fields bit for bit (velocity rows are independent; the stencil of an interior row only needs its neighbours)."""
    from ludvm_b200.sharded import grid_slab
    g = load_golden("freevort_tf3")
    s = _sim(g)
    kw = dict(g["ff_kw"])
    s.flowfield(**kw)
    full = {k: getattr(s, k).copy() for k in ("x_ff", "z_ff", "u_ff", "w_ff", "ome_ff")}
    nx = full["x_ff"].shape[0]
    parts = {k: [] for k in full}
    for rank in range(3):
        r0, r1, h0, h1 = grid_slab(nx, 3, rank)
        s.flowfield(**kw, rows=(h0, h1 - h0))
        assert s.u_ff.shape[1] == h1 - h0
        for k in full:
            a = getattr(s, k)
            parts[k].append(a[r0 - h0:r1 - h0] if a.ndim == 2 else a[:, r0 - h0:r1 - h0])
    for k in full:
        assert biteq(np.concatenate(parts[k], axis=0 if full[k].ndim == 2 else 1), full[k]), k
    assert biteq(full["u_ff"], g["u_ff"])

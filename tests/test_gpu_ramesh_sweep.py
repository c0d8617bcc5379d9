"""GPU parity tests of the one-CTA-per-case driver: method='Ramesh' (LUDVM.py:683-739, :807-909) and batched
parameter sweeps (BASELINE.json configs[3]) against the reference's golden histories and the oracle."""
import numpy as np
import pytest

from conftest import biteq, golden_tables, load_golden

pytestmark = pytest.mark.gpu

README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")


def test_ramesh_bit_equal_to_reference_golden():
    from ludvm_b200 import LUDVM
    g = load_golden("ramesh_tf2")
    tb = golden_tables(g)
    tb["sum_free"] = float(np.sum(tb["free_g"]))
    s = LUDVM(**g["kw"], verbose=False, run=False)
    s.time_loop(tables=tb)
    s.compute_coefficients()
    for k in ("Fn", "Fs", "L", "D", "M", "LESP", "LESP_prev", "LEV_shed", "fourier", "Cl", "Cd", "Cm"):
        assert biteq(getattr(s, k), g[k]), k
    for k in ("TEV", "LEV", "bound"):
        assert biteq(s.circulation[k], g["circ_" + k]), k
    for k in ("TEV", "LEV", "FREE"):
        assert biteq(s.path[k][g["path_rows"]], g["path_" + k + "_rows"]), k
    assert [s.itev, s.ilev] == list(g["itev_ilev"])


def test_ramesh_vs_oracle_and_vs_faure(oracle):
    from ludvm_b200 import LUDVM
    kw = dict(README, tf=5, method="Ramesh")
    s, o = LUDVM(**kw, verbose=False), oracle.OracleLUDVM(**kw)
    for k in ("Cl", "Cd", "Cm", "LESP", "LEV_shed", "fourier"):
        assert biteq(getattr(s, k), getattr(o, k)), k
    assert biteq(s.path["TEV"], o.path["TEV"]) and biteq(s.path["LEV"], o.path["LEV"])
    f = LUDVM(**dict(README, tf=5), verbose=False)
    assert np.max(np.abs(f.Cl[:41] - s.Cl[:41])) < 1e-12      # the two methods agree before chaos


def test_sweep_matches_reference_golden_and_single_runs(oracle):
    """Five (LESPcrit, k) corner/interior cases of the 64x64 sweep grid: the batched one-CTA-per-case result is
    bit-equal to the reference histories (when the host libm agrees with the fixture's) and to the oracle."""
    from ludvm_b200 import sweep
    combos = [(0.1, 0.1), (0.1, 1.0), (0.4, 0.1), (0.4, 1.0), (0.25, 0.55)]
    cases = [dict(README, LESPcrit=lc, k=k) for lc, k in combos]
    res = sweep.run_sweep(cases, mode="exact")
    for a, kw in enumerate(cases):
        o = oracle.OracleLUDVM(**kw)
        for k in ("Cl", "Cd", "Cm", "LESP", "LEV_shed", "L", "M"):
            assert biteq(res[k][a], getattr(o, k)), (a, k)
        assert biteq(res["circulation_TEV"][a][:-1], o.circulation["TEV"])
        assert biteq(res["circulation_LEV"][a][:-1], o.circulation["LEV"])
        g = load_golden("sweep_%d" % a)
        if biteq(o.alpha, g["alpha"]):
            for k in ("Cl", "Cd", "Cm", "LESP", "LEV_shed"):
                assert biteq(res[k][a], g[k]), (a, k)


def test_sweep_slices_and_fast_mode():
    from ludvm_b200 import sweep
    cases = sweep.lespcrit_k_grid(np.linspace(0.1, 0.4, 3), np.linspace(0.1, 1.0, 4), **dict(README, tf=5))
    full = sweep.run_sweep(cases, mode="exact")
    lo, hi = sweep.run_sweep(cases, case_slice=slice(0, 7)), sweep.run_sweep(cases, case_slice=slice(7, None))
    assert biteq(np.concatenate([lo["Cl"], hi["Cl"]]), full["Cl"])
    fast = sweep.run_sweep(cases, mode="fast")
    assert np.max(np.abs(fast["Cl"][:, :40] - full["Cl"][:, :40])) < 1e-10


def test_sweep_with_many_free_vortices_covers_the_wide_shared_memory_tiles(oracle):
    """700 free vortices from step 0 put > 512 target rows into the one-CTA driver at once, i.e. the 4-rows-per-thread
    variant of its shared-memory convection (fast mode); exact mode of the same sweep stays bit-equal to the oracle."""
    from ludvm_b200 import sweep
    rng = np.random.default_rng(13)
    nf = 700
    xy = np.stack([rng.uniform(-4.0, -0.5, nf), rng.uniform(-0.6, 0.6, nf)])
    gam = rng.standard_normal(nf) * 2e-3
    base = dict(README, tf=1, circulation_freevort=gam, xy_freevort=xy)
    cases = [dict(base, LESPcrit=0.15), dict(base, LESPcrit=0.3, k=0.5)]
    ex, fa = sweep.run_sweep(cases, mode="exact"), sweep.run_sweep(cases, mode="fast")
    o = oracle.OracleLUDVM(**cases[1])
    for k in ("L", "M", "LESP", "LEV_shed"):
        assert biteq(ex[k][1], getattr(o, k)), k
    for k in ("L", "D", "M"):
        assert np.max(np.abs(fa[k] - ex[k])) <= 1e-10 * np.max(np.abs(ex[k])), k
    assert biteq(fa["LEV_shed"], ex["LEV_shed"])

"""CPU tests: the C-ABI library loads and exports every symbol include/ludvm_b200.h declares (no compute without a
GPU), fails loudly without a device, and the host-side logic of the drop-in class matches the oracle's tables."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, biteq


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ludvm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ludvm_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from ludvm_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    assert set(syms) == set(built.SIGNATURES), "ctypes binding and header disagree"
    assert built.load().ludvm_abi_version() == 1


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.LudvmError, match="no CUDA device"):
        built.Context(0)
    from ludvm_b200 import ops
    with pytest.raises(built.LudvmError):
        ops.induced_velocity(np.ones(3), np.zeros(3), np.zeros(3), np.ones(2), np.ones(2), 0.1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ludvm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ludvm_oracle" not in txt, f


@pytest.mark.parametrize("over", [dict(), dict(dt=1e-2, tf=2, Npoints=61, Ncoeffs=12, chord=1.3, Uinf=1.7, alpha_m=3,
                                               alpha_max=20, k=0.7, h_max=0.4, phi=75)])
def test_host_tables_equal_oracle_tables(oracle, over):
    from ludvm_b200 import LUDVM
    kw = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
    kw.update(over)
    s = LUDVM(**kw, verbose=False, run=False)
    o = oracle.OracleLUDVM(**kw, run=False)
    ts, to = s.step_tables(), oracle.tables_from(o)
    for k, v in to.items():
        assert biteq(ts[k], v) if isinstance(v, np.ndarray) else ts[k] == v, k
    assert ts["sum_free"] == float(np.sum(o.circulation_freevort))
    for k in o.airfoil:
        assert biteq(s.airfoil[k], o.airfoil[k]), k


def test_host_tables_equal_golden_tables():
    """Same check against the tables of the real reference run (committed fixture)."""
    from conftest import golden_tables, load_golden
    from ludvm_b200 import LUDVM
    g = load_golden("readme")
    s = LUDVM(**g["kw"], verbose=False, run=False)
    if not biteq(s.alpha, g["alpha"]):
        pytest.skip("this host's libm rounds differently from the fixture's host")
    ts, tg = s.step_tables(), golden_tables(g)
    for k, v in tg.items():
        assert biteq(ts[k], v) if isinstance(v, np.ndarray) else ts[k] == v, k


def test_motion_plunge_fixed_and_cambered_section():
    from ludvm_b200 import LUDVM
    s = LUDVM(tf=3, dt=5e-2, Npoints=41, verbose=False, run=False, Naca="2412")
    assert s.airfoil["eta"].max() > 0.015 and abs(s.airfoil["eta"][0]) < 1e-15
    s.motion_plunge(G=1, T=2)          # the reference raises TypeError here (LUDVM.py:520)
    assert s.alpha_e.shape == (s.nt,) and np.all(np.diff(s.hpiv) <= 1e-15)
    assert s.hpiv[-1] == s.hpiv[np.searchsorted(s.t, 2.0, side="right") - 1]


def test_store_history_argument_is_validated_on_the_host():
    """store_history: True/False/k (strided snapshots); negative values are rejected before any device work."""
    from ludvm_b200 import LUDVM
    kw = dict(t0=0, tf=1, dt=5e-2, Npoints=81, Naca="0012", verbose=False, run=False)
    assert LUDVM(**kw).store_history == 1 and LUDVM(**kw, store_history=False).store_history == 0
    assert LUDVM(**kw, store_history=7).store_history == 7
    with pytest.raises(ValueError):
        LUDVM(**kw, store_history=-2)


def test_sweep_host_structs_per_case(monkeypatch):
    """run_sweep builds the parameter / table structs once per distinct motion and copies them per case: every case
    must still carry its own LESPcrit, the tables of its own reduced frequency, and the requested mode.  The C entry
    point is replaced by a recorder (host logic only; no device work)."""
    import ctypes as C
    import types
    from ludvm_b200 import LUDVM, sweep, _lib
    seen = {}

    class FakeLib:
        def ludvm_sweep_run(self, ctx, n, P, T, out, stride):
            seen["n"], seen["stride"] = n, stride
            seen["p"] = [(P[i].nt, P[i].P, P[i].Nc, P[i].mode, P[i].store_history, P[i].lespcrit, P[i].dt, P[i].vc4) for i in range(n)]
            seen["cos1"] = [T[i].cos_a[1] for i in range(n)]            # first-step kinematics: depends on k only
            seen["ptr"] = [C.cast(T[i].gp, C.c_void_p).value for i in range(n)]
            return 0

    monkeypatch.setattr(sweep, "load", lambda: FakeLib())
    base = dict(t0=0, tf=1, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, Naca="0012")
    lcs, ks = (0.1, 0.25, 0.4), (0.2, 0.9)
    cases = sweep.lespcrit_k_grid(lcs, ks, **base)
    res = sweep.run_sweep(cases, mode="fast", ctx=types.SimpleNamespace(handle=None))
    assert seen["n"] == 6 and seen["stride"] == len(sweep.SW_FIELDS) * res["nt"] and res["Cl"].shape == (6, res["nt"])
    ref = {k: LUDVM(**dict(base, k=k), verbose=False, run=False) for k in ks}
    for i, kw in enumerate(cases):
        nt, P, Nc, mode, hist, lc, dt, vc4 = seen["p"][i]
        r = ref[kw["k"]]
        assert (nt, P, Nc, mode, hist) == (r.nt, 80, 30, _lib.MODES["fast"], 0)
        assert lc == kw["LESPcrit"] and dt == 5e-2 and vc4 == r.v_core ** 4
        assert seen["cos1"][i] == np.cos(r.alpha[1])
    assert len(set(seen["ptr"])) == len(ks)                               # one table set per distinct motion
    assert seen["ptr"][0] == seen["ptr"][2] == seen["ptr"][4] and seen["ptr"][0] != seen["ptr"][1]
    sl = sweep.run_sweep(cases, mode="exact", ctx=types.SimpleNamespace(handle=None), case_slice=slice(3, 5))
    assert seen["n"] == 2 and [p[5] for p in seen["p"]] == [0.25, 0.4] and seen["p"][0][3] == _lib.MODES["exact"]


def test_sweep_table_sharing_keys_long_arrays_by_content(monkeypatch):
    """Cases whose free-vortex arrays differ only in entries that numpy's repr() elides (> 1000 elements) must not share
    one table set; a case that omits LESPcrit gets the constructor default (0.2), not the value of the case that
    happened to create the shared entry."""
    import ctypes as C
    import types
    from ludvm_b200 import sweep
    seen = {}

    class FakeLib:
        def ludvm_sweep_run(self, ctx, n, P, T, out, stride):
            seen["lc"] = [P[i].lespcrit for i in range(n)]
            seen["g600"] = [T[i].free_g[600] for i in range(n)]
            seen["ptr"] = [C.cast(T[i].free_g, C.c_void_p).value for i in range(n)]
            return 0

    monkeypatch.setattr(sweep, "load", lambda: FakeLib())
    rng = np.random.default_rng(0)
    nf = 1500
    xy, g1 = np.stack([rng.uniform(-3, -1, nf), rng.uniform(-0.5, 0.5, nf)]), rng.standard_normal(nf) * 1e-3
    g2 = g1.copy()
    g2[600] += 1e-3                                           # inside the part repr() prints as '...'
    assert repr(g1) == repr(g2)
    base = dict(t0=0, tf=0.5, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, Naca="0012", xy_freevort=xy)
    cases = [dict(base, circulation_freevort=g1, LESPcrit=0.35), dict(base, circulation_freevort=g2),
             dict(base, circulation_freevort=g1.copy())]
    sweep.run_sweep(cases, mode="fast", ctx=types.SimpleNamespace(handle=None))
    assert seen["lc"] == [0.35, 0.2, 0.2]
    assert seen["g600"] == [g1[600], g2[600], g1[600]]
    assert seen["ptr"][0] == seen["ptr"][2] != seen["ptr"][1]

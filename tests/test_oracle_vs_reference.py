"""CPU tests that pin the oracle against the UNMODIFIED reference, run live.  Only possible where
/root/reference exists (the development container); skipped elsewhere."""
import contextlib
import io

import numpy as np
import pytest

from conftest import biteq
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present on this machine")


def _cmp(r, o):
    for k in ("Cl", "Cd", "Cm", "LESP", "LESP_prev", "LEV_shed", "fourier", "Fn", "Fs", "M", "alpha", "h_dot"):
        assert biteq(getattr(r, k), getattr(o, k)), k
    for k in ("TEV", "LEV", "FREE", "airfoil", "airfoil_gamma_points"):
        assert biteq(r.path[k], o.path[k]), k
    for k in ("TEV", "LEV", "bound", "airfoil", "gamma_airfoil", "Gamma_airfoil", "IC"):
        assert biteq(r.circulation[k], o.circulation[k]), k
    assert (r.itev, r.ilev) == (o.itev, o.ilev)


@pytest.mark.parametrize("over", [dict(tf=4), dict(tf=0.5),   # (tf=0.5: fewer wake vortices than chord stations)
                                  dict(tf=2, method="Ramesh", LESPcrit=0.12),
                                  dict(tf=3, dt=2e-2, Npoints=61, Ncoeffs=12, chord=1.3, Uinf=1.7, alpha_m=3,
                                       alpha_max=20, k=0.7)])
def test_time_loop_bit_equal(oracle, over):
    kw = dict(ref_loader.README_KW, **over)
    _cmp(ref_loader.run(**kw), oracle.OracleLUDVM(**kw))


def test_free_vortices_and_flowfield(oracle):
    ref = ref_loader.load()
    xy, g = ref.generate_free_single_vortex()
    kw = dict(ref_loader.README_KW, tf=2, circulation_freevort=g, xy_freevort=xy.T)
    r, o = ref_loader.run(**kw), oracle.OracleLUDVM(**kw)
    _cmp(r, o)
    ff = dict(xmin=-3.0, xmax=0.4, zmin=-1.2, zmax=0.6, dr=0.1, tsteps=[0, 5, 30])
    with contextlib.redirect_stdout(io.StringIO()):
        r.flowfield(**ff)
    o.flowfield(**ff)
    for k in ("u_ff", "w_ff", "ome_ff"):
        assert biteq(getattr(r, k), getattr(o, k)), k


def test_induced_velocity_random(oracle):
    rng = np.random.default_rng(5)
    b = ref_loader.bare(0.03)
    for npnt, nw in [(1, 1), (4, 7), (4, 8), (4, 129), (17, 777), (3, 4100)]:
        g, xw, zw = rng.standard_normal(nw), rng.uniform(-5, 0, nw), rng.uniform(-1, 1, nw)
        xp, zp = rng.uniform(-5, 0, npnt), rng.uniform(-1, 1, npnt)
        u, w = b.induced_velocity(g, xw, zw, xp, zp)
        uo, wo = oracle.induced_velocity(g, xw, zw, xp, zp, 0.03)
        assert biteq(u, uo) and biteq(w, wo)


def test_airfoil_downwash_bit_equal(oracle):
    """LUDVM.airfoil_downwash (LUDVM.py:572-595) on the state of a finished run, at several step indices."""
    kw = dict(ref_loader.README_KW, tf=3)
    r, o = ref_loader.run(**kw), oracle.OracleLUDVM(**kw)
    rng = np.random.default_rng(3)
    for i, nw in [(1, 1), (7, 9), (30, 130), (59, 500)]:
        g, xw, zw = rng.standard_normal(nw) * 1e-2, rng.uniform(-3, 0, nw), rng.uniform(-0.5, 0.5, nw)
        assert biteq(r.airfoil_downwash(g, xw, zw, i), o.airfoil_downwash(g, xw, zw, i)), i


def test_plunge_manoeuvre_bit_equal(oracle, monkeypatch):
    """motion_plunge (LUDVM.py:459-547) + time_loop.  The reference's one-argument np.arctan2 (LUDVM.py:520) raises, so
    the reference is run with arctan2 patched to the documented fix (arctan2(h_dot/Uinf, 1)); alpha_e is a dead local
    there, every table the time loop consumes is untouched by the patch."""
    real = np.arctan2
    monkeypatch.setattr(np, "arctan2", lambda *a, **k: real(*a, **k) if len(a) == 2 else real(a[0], 1.0))
    kw = dict(ref_loader.README_KW, tf=3, alpha_m=4)
    r = ref_loader.run(**kw)                     # the constructor runs the sinusoidal case first (LUDVM.py:282-295)
    r.motion_plunge(G=0.8, T=2, alpha_m=4, h0=0, x0=0.25)
    with contextlib.redirect_stdout(io.StringIO()):
        r.time_loop()
        r.compute_coefficients()
    o = oracle.OracleLUDVM(**kw, run=False)
    o.motion_plunge(G=0.8, T=2, alpha_m=4, h0=0, x0=0.25)
    o.time_loop()
    o.compute_coefficients()
    _cmp(r, o)
    assert np.ptp(o.h_dot) > 0.5 and np.all(o.alpha_dot == 0)

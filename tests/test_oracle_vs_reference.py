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

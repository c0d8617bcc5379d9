"""CPU tests: the oracle (oracle/ludvm_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference (LUDVM.py)."""
import hashlib

import numpy as np
import pytest

from conftest import biteq, golden_tables, load_golden

SIM_FIXTURES = ["readme", "ramesh_tf2", "freevort_tf3", "hires_200"]
HIST = ("Fn", "Fs", "L", "D", "T", "M", "LESP", "LESP_prev", "LEV_shed", "fourier")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_percall_golden(oracle):
    g = load_golden("percall")
    vc = float(g["v_core"])
    for c in range(int(g["ncases"])):
        u, w = oracle.induced_velocity(g["c%d_g" % c], g["c%d_xw" % c], g["c%d_zw" % c],
                                       g["c%d_xp" % c], g["c%d_zp" % c], vc)
        assert biteq(u, g["c%d_u" % c]) and biteq(w, g["c%d_w" % c]), "case %d" % c
    u, w = oracle.induced_velocity(np.array([1]), np.array([0.3]), np.array([-0.2]), g["b_xp"], g["b_zp"], vc)
    assert biteq(u, g["b_u"]) and biteq(w, g["b_w"])
    u, w = oracle.induced_velocity(g["i_g"], g["i_xw"], g["i_zw"], g["b_xp"], g["b_zp"], vc, viscous=False)
    assert biteq(u, g["i_u"]) and biteq(w, g["i_w"])


def test_percall_thread_count_independent(oracle):
    rng = np.random.default_rng(3)
    xw, zw, g = rng.uniform(-20, 0, 3000), rng.uniform(-4, 4, 3000), rng.standard_normal(3000)
    u1, w1 = oracle.induced_velocity(g, xw, zw, xw, zw, 0.065, nthreads=1)
    u2, w2 = oracle.induced_velocity(g, xw, zw, xw, zw, 0.065, nthreads=0)
    assert biteq(u1, u2) and biteq(w1, w2)


@pytest.mark.parametrize("name", SIM_FIXTURES)
def test_sim_from_golden_tables(oracle, name):
    """Run the oracle step loop from the fixture's host tables; every stored output must be bit-equal."""
    g = load_golden(name)
    kw = dict(g["kw"])
    for k in ("circulation_freevort", "xy_freevort"):
        if k in kw:
            kw[k] = np.array(kw[k])
    o = oracle.OracleLUDVM(**kw, run=False)
    o.time_loop(tables=golden_tables(g))
    o.compute_coefficients()
    for k in HIST + ("Cl", "Cd", "Cm", "Cn", "Cs", "Ct"):
        assert biteq(getattr(o, k), g[k]), k
    for k in ("TEV", "LEV", "bound"):
        assert biteq(o.circulation[k], g["circ_" + k]), k
    for k in ("airfoil", "gamma_airfoil", "Gamma_airfoil"):
        assert sha(o.circulation[k]) == str(g["circ_" + k + "_sha256"]), k
    for k in ("TEV", "LEV", "FREE"):
        assert sha(o.path[k]) == str(g["path_" + k + "_sha256"]), k
        assert biteq(o.path[k][g["path_rows"]], g["path_" + k + "_rows"]), k
    assert [o.itev, o.ilev] == list(g["itev_ilev"])


@pytest.mark.parametrize("name", SIM_FIXTURES + ["sweep_%d" % i for i in range(5)])
def test_sim_host_tables_reproduce(oracle, name):
    """Full oracle run including its own host-side geometry/kinematics (numpy on THIS host).  libm may round a
    cos() differently on another CPU; then the histories can only agree until chaos amplifies the ulp."""
    g = load_golden(name)
    kw = dict(g["kw"])
    for k in ("circulation_freevort", "xy_freevort"):
        if k in kw:
            kw[k] = np.array(kw[k])
    o = oracle.OracleLUDVM(**kw)
    same_host = biteq(o.alpha, g["alpha"]) and biteq(o.h_dot, g["h_dot"])
    if same_host:
        for k in ("Cl", "Cd", "Cm", "LESP", "LEV_shed"):
            assert biteq(getattr(o, k), g[k]), k
    else:  # different libm: early history only
        n = min(60, len(g["Cl"]))
        np.testing.assert_allclose(o.Cl[:n], g["Cl"][:n], rtol=1e-9, atol=1e-12)


def test_flowfield_golden(oracle):
    g = load_golden("freevort_tf3")
    kw = dict(g["kw"])
    kw["circulation_freevort"], kw["xy_freevort"] = np.array(kw["circulation_freevort"]), np.array(kw["xy_freevort"])
    o = oracle.OracleLUDVM(**kw, run=False)
    o.time_loop(tables=golden_tables(g))
    o.flowfield(**g["ff_kw"])
    for k in ("x_ff", "z_ff", "u_ff", "w_ff", "ome_ff"):
        assert biteq(getattr(o, k), g[k]), k


def test_numpy_primitives(oracle):
    """Appendix A: pairwise tree, trapz and the 2x2 solve against numpy on this host."""
    rng = np.random.default_rng(11)
    trapz = getattr(np, "trapezoid", None) or np.trapz
    for n in list(range(0, 140)) + [255, 256, 257, 1000, 4097, 20001]:
        a = rng.standard_normal(n)
        assert oracle.np_sum(a) == np.sum(a)
        if n >= 2:
            x = np.sort(rng.uniform(0, 3, n))
            assert oracle.np_trapz(a, x) == trapz(a, x)
    nbad = 0
    for _ in range(2000):
        A = np.array([[1 + rng.normal(0, .3), 1 + rng.normal(0, .3)], [rng.normal(0, .2), rng.normal(0, .2)]])
        b = rng.standard_normal(2)
        nbad += not biteq(oracle.solve2x2(A, b), np.linalg.solve(A, b))
    # LAPACK's 2x2 path is host-BLAS dependent; the recipe is pinned by the golden runs, this
    # only reports drift of the local BLAS.
    if nbad:
        pytest.skip("local BLAS dgesv differs from the pinned recipe on %d/2000 systems" % nbad)


def test_bccheck_reconstruction_and_residual(oracle):
    """BCcheck=True (LUDVM.py:1144-1161, shape bug fixed) is evaluated after the run from the stored history.  The
    delicate part is rebuilding the wake as it stood before each step's convection; check it by redoing the
    convection of every step from the rebuilt wake (LUDVM.py:1095-1127) and landing on the stored next row bit for
    bit.  The residual itself must vanish to rounding.  (Parity of BC against the reference is unpinned: it raises.)"""
    kw = dict(t0=0, tf=3, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
    o = oracle.OracleLUDVM(**kw, run=False)
    o.time_loop(BCcheck=True)
    assert o.BC.shape == (o.nt - 1, o.Npoints) and o.ilev > 3
    assert np.max(np.abs(o.BC)) < 1e-13 and np.any(o.BC != 0) and np.all(o.BC[:, -1] == 0)
    ilev = 0
    gp = o.path["airfoil_gamma_points"]
    for i in range(1, o.nt):
        itev = i - 1
        g, xw, zw = o._wake_before_convection(i, ilev)
        nT, nL = itev + 1, ilev + 1
        cf = o.circulation["airfoil"][itev]
        uw, ww = o.induced_velocity(g, xw, zw, xw, zw)
        uf, wf = o.induced_velocity(cf, gp[i, 0], gp[i, 1], xw, zw)
        xn, zn = xw + o.dt * (uw + uf), zw + o.dt * (ww + wf)
        assert biteq(xn[:nT], o.path["TEV"][i, 0, :nT]) and biteq(zn[:nT], o.path["TEV"][i, 1, :nT]), i
        assert biteq(xn[nT:nT + nL], o.path["LEV"][i, 0, :nL]) and biteq(zn[nT:nT + nL], o.path["LEV"][i, 1, :nL]), i
        assert biteq(xn[nT + nL:], o.path["FREE"][i, 0]), i
        if o.LEV_shed[i] != -1:
            ilev += 1


def test_naca4_mean_line_against_the_published_formula():
    """The camber line the reference takes from the un-vendored `airfoils` package (LUDVM.py:328-335) is the NACA
    4-digit mean line (Abbott & von Doenhoff, Theory of Wing Sections, eq. 6.4): y_c = m/p^2 (2 p x - x^2) ahead of the
    maximum-camber station p, m/(1-p)^2 ((1 - 2 p) + 2 p x - x^2) behind it.  Pin `_naca4_camber` at tabulated stations
    so that a typo cannot hide behind the symmetric sections every BASELINE config uses."""
    from ludvm_b200.ludvm import _naca4_camber
    x = np.array([0.0, 0.1, 0.2, 0.4, 0.7, 1.0])
    assert np.allclose(_naca4_camber("2412", x), [0.0, 0.00875, 0.015, 0.02, 0.015, 0.0], rtol=0, atol=1e-15)
    assert np.allclose(_naca4_camber("4415", x), [0.0, 0.0175, 0.03, 0.04, 0.03, 0.0], rtol=0, atol=1e-15)
    assert np.allclose(_naca4_camber("6309", np.array([0.15, 0.3, 0.65, 1.0])),
                       [0.06 / 0.09 * (0.09 - 0.0225), 0.06, 0.06 / 0.49 * (0.4 + 0.39 - 0.4225), 0.0], rtol=0, atol=1e-15)
    assert not np.any(_naca4_camber("0012", np.linspace(0, 1, 33)))
    for code, m, p in (("2412", 0.02, 0.4), ("6309", 0.06, 0.3), ("9521", 0.09, 0.5)):
        xs = np.linspace(0, 1, 2001)
        yc = _naca4_camber(code, xs)
        assert abs(yc.max() - m) < 1e-12 and abs(xs[np.argmax(yc)] - p) < 1e-3       # maximum camber m at station p
        h = 1e-7                                                                      # slope continuous (zero) at p
        assert abs(_naca4_camber(code, np.array([p + h]))[0] - _naca4_camber(code, np.array([p - h]))[0]) < 1e-12

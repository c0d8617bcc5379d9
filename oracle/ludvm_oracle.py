"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (`oracle/ludvm_oracle.c`).

`OracleLUDVM` reruns the reference's path (LUDVM.py:231-297 constructor -> :299 airfoil_generation ->
:382 motion_sinusoidal -> :597 time_loop -> :1173 compute_coefficients, and :1186 flowfield) with the
arithmetic in scalar C that is bit-identical to the reference's numpy.  It is the checker for the CUDA
product and the timed CPU baseline of bench.py; it must never be imported from `ludvm_b200/`.

Only symmetric NACA 00xx sections are supported (the reference delegates camber to the un-vendored PyPI
package `airfoils`, no version pinned; for 00xx the camber is identically zero so nothing is lost --
SURVEY.md 8c).  Cambered sections: parity unpinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libludvm_oracle.so")
_dp = C.POINTER(C.c_double)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "ludvm_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _SimIn(C.Structure):
    _fields_ = ([(n, C.c_long) for n in ("nt", "P", "Nc", "nfree")] + [("method", C.c_int)] +
                [(n, C.c_double) for n in ("dt", "Uinf", "chord", "rho", "piv", "lespcrit", "vc4", "ic",
                                           "maxerror", "epsilon")] +
                [("maxiter", C.c_long), ("a0_init", C.c_double), ("a1_init", C.c_double)] +
                [(n, _dp) for n in ("cos_a", "sin_a", "alpha_dot", "h_dot", "gp", "le", "te",
                                    "detadx_p", "eta_p", "x_p", "theta_p", "dtheta", "cos_tp", "sin_tp",
                                    "cosn", "sinn", "free_g", "free_xz")])


class _SimOut(C.Structure):
    _fields_ = ([(n, _dp) for n in ("path_tev", "path_lev", "path_free", "g_tev", "g_lev", "g_bound",
                                    "g_airfoil", "gamma_airfoil", "Gamma_airfoil", "fourier",
                                    "lesp", "lesp_prev", "lev_shed", "Fn", "Fs", "L", "D", "T", "M")] +
                [("itev", C.POINTER(C.c_long)), ("ilev", C.POINTER(C.c_long))])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_np_sum.restype = C.c_double
        L.oracle_np_sum.argtypes = [_dp, C.c_long]
        L.oracle_np_trapz.restype = C.c_double
        L.oracle_np_trapz.argtypes = [_dp, _dp, C.c_long]
        L.oracle_solve2x2.restype = None
        L.oracle_solve2x2.argtypes = [_dp, _dp, _dp]
        L.oracle_induced_velocity.restype = C.c_int
        L.oracle_induced_velocity.argtypes = [_dp, C.c_long, _dp, _dp, C.c_long, _dp, _dp, C.c_long,
                                              C.c_double, _dp, _dp, C.c_int]
        L.oracle_sim_run.restype = C.c_int
        L.oracle_sim_run.argtypes = [C.POINTER(_SimIn), C.POINTER(_SimOut)]
        L.oracle_vorticity.restype = None
        L.oracle_vorticity.argtypes = [_dp, _dp, _dp, _dp, C.c_long, C.c_long, C.c_long, _dp]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def np_sum(a):
    a = _f64(a)
    return lib().oracle_np_sum(_p(a), a.size)


def np_trapz(y, x):
    y, x = _f64(y), _f64(x)
    return lib().oracle_np_trapz(_p(y), _p(x), y.size)


def solve2x2(A, b):
    A, b, x = _f64(A), _f64(b), np.zeros(2)
    lib().oracle_solve2x2(_p(A), _p(b), _p(x))
    return x


def induced_velocity(circulation, xw, zw, xp, zp, v_core, viscous=True, nthreads=0):
    """LUDVM.induced_velocity (LUDVM.py:549-570) with the object's `v_core` passed explicitly."""
    g, xw, zw, xp, zp = map(_f64, (np.atleast_1d(circulation), xw, zw, xp, zp))
    vc4 = float(v_core) ** 4 if viscous == True else 0.0  # noqa: E712  (the reference compares with ==)
    u, w = np.empty(xp.size), np.empty(xp.size)
    rc = lib().oracle_induced_velocity(_p(g), g.size, _p(xw), _p(zw), xw.size, _p(xp), _p(zp), xp.size,
                                       vc4, _p(u), _p(w), int(nthreads))
    if rc != 0:
        raise ValueError("oracle_induced_velocity: bad shapes")
    return u, w


def vorticity(x, z, u, w):
    x, z, u, w = map(_f64, (x, z, u, w))
    ome = np.zeros_like(u)
    lib().oracle_vorticity(_p(x), _p(z), _p(u), _p(w), u.shape[0], u.shape[1], u.shape[2], _p(ome))
    return ome


def tables_from(obj):
    """Every host-evaluated table/scalar the step consumes, evaluated with numpy exactly as the reference
    does.  `obj` is a reference `LUDVM` object or an `OracleLUDVM` (same
    attribute names)."""
    self = obj
    af, P, Nc = self.airfoil, self.Npoints - 1, self.Ncoeffs
    tp = af['theta_panel']
    A0 = np.sin(self.alpha_m)                                         # degrees quirk, LUDVM.py:645
    ic = np.sum(self.circulation_freevort) + self.Uinf * self.chord * np.pi * (A0 + 0 / 2)
    return dict(
        nt=self.nt, P=P, Nc=Nc, nfree=self.n_freevort, method=1 if self.method == 'Ramesh' else 0,
        dt=float(self.dt), Uinf=float(self.Uinf), chord=float(self.chord), rho=float(self.rho),
        piv=float(self.piv), lespcrit=float(self.LESPcrit), vc4=float(self.v_core ** 4), ic=float(ic),
        maxerror=self.maxerror, epsilon=self.epsilon, maxiter=self.maxiter,
        a0_init=float(A0), a1_init=0.0,
        cos_a=_f64(np.cos(self.alpha)), sin_a=_f64(np.sin(self.alpha)),
        alpha_dot=_f64(self.alpha_dot), h_dot=_f64(self.h_dot),
        gp=_f64(self.path['airfoil_gamma_points']),
        le=_f64(self.path['airfoil'][:, :, 0]), te=_f64(self.path['airfoil'][:, :, -1]),
        detadx_p=_f64(af['detadx_panel']), eta_p=_f64(af['eta_panel']), x_p=_f64(af['x_panel']),
        theta_p=_f64(tp), dtheta=_f64(af['theta'][1:] - af['theta'][:-1]),
        cos_tp=_f64(np.cos(tp)), sin_tp=_f64(np.sin(tp)),
        cosn=_f64(np.stack([np.cos(n * tp) for n in range(Nc)])),
        sinn=_f64(np.stack([np.sin(n * tp) for n in range(Nc)])),
        free_g=_f64(self.circulation_freevort), free_xz=_f64(self.xy_freevort))


class OracleLUDVM:
    """Same constructor, attributes and methods as the reference class for the path under test."""

    def __init__(self, t0=0, tf=12, dt=1.5e-2, chord=1, rho=1.225, Uinf=1, Npoints=80, Ncoeffs=30,
                 LESPcrit=0.2, Naca='0012', foil_filename=None, G=1, T=2, alpha_m=0, alpha_max=10,
                 k=0.2 * np.pi, phi=90, h_max=1, verbose=True, method='Faure',
                 circulation_freevort=None, xy_freevort=None, run=True, motion='cos', nsteps=None):
        if Naca is None or Naca[:2] != '00':
            raise NotImplementedError("oracle: symmetric NACA 00xx only (airfoils package absent)")
        self.t0, self.tf, self.dt = t0, tf, dt
        self.chord, self.rho, self.Uinf = chord, rho, Uinf
        self.Npoints, self.Ncoeffs, self.LESPcrit = Npoints, Ncoeffs, LESPcrit
        self.piv = 0.25 * chord
        self.maxerror, self.maxiter, self.epsilon, self.xgamma = 1e-10, 50, 1e-4, 0.25
        self.method = method
        self.t = np.arange(t0, tf + dt, dt)                               # LUDVM.py:254
        if nsteps is not None:                                           # prefix runs (tests, baselines)
            self.t = self.t[:nsteps + 1]
        self.nt = len(self.t)
        self.dt_star = dt * Uinf / chord
        self.v_core = 1.3 * self.dt_star * chord                          # LUDVM.py:260
        self.alpha_m = alpha_m
        if circulation_freevort is not None and xy_freevort is not None:  # LUDVM.py:268-277
            self.n_freevort = len(circulation_freevort)
            self.circulation_freevort = np.asarray(circulation_freevort)
            self.xy_freevort = np.asarray(xy_freevort)
        else:
            self.n_freevort = 1
            self.circulation_freevort = np.array([0])
            self.xy_freevort = np.array([0, 0])[:, np.newaxis]
        self._section()
        self._kinematics(alpha_m, alpha_max, h_max, k, phi, 0, 0, motion)
        if run:
            self.time_loop()
            self.compute_coefficients()

    # LUDVM.py:337-372 restricted to a zero camber line ('theta' spacing)
    def _section(self):
        N, c = self.Npoints, self.chord
        theta = np.linspace(0, np.pi, N)
        x = c / 2 * (1 - np.cos(theta))
        xa = c * 0.5 * (np.linspace(0, 1, N) + np.linspace(0, 1, N))
        eta = np.interp(x, xa, np.zeros(N))
        x_panel = x[:-1] + self.xgamma * (x[1:] - x[:-1])
        eta_panel = np.interp(x_panel, x, eta)
        theta_panel = np.arccos(1 - 2 * x_panel / c)
        z, zp = np.zeros(N), np.zeros(N - 1)   # all finite differences of a zero camber line are zero
        self.airfoil = {'x': x, 'theta': theta, 'eta': eta, 'detadx': z.copy(), 'detadtheta': z.copy(),
                        'x_panel': x_panel, 'theta_panel': theta_panel, 'eta_panel': eta_panel,
                        'detadx_panel': zp.copy(), 'detadtheta_panel': zp.copy()}

    # LUDVM.py:382-457
    def _kinematics(self, alpha_m, alpha_max, h_max, k, phi, h0, x0, motion):
        pi, U, nt = np.pi, self.Uinf, self.nt
        f = k * U / (2 * pi * self.chord)
        self.f = f
        alpha_m, alpha_max, phi = alpha_m * pi / 180, alpha_max * pi / 180, phi * pi / 180
        al, ald, h, hd, x = (np.zeros(nt) for _ in range(5))
        for i in range(nt):
            ti = self.t[i]
            if motion == 'cos':
                al[i] = alpha_m + alpha_max * np.cos(2 * pi * f * ti + phi)
                ald[i] = - alpha_max * 2 * pi * f * np.sin(2 * pi * f * ti + phi)
                h[i] = h0 + h_max * np.cos(2 * pi * f * ti)
                hd[i] = - h_max * 2 * pi * f * np.sin(2 * pi * f * ti)
            else:
                al[i] = alpha_m + alpha_max * np.sin(2 * pi * f * ti + phi)
                ald[i] = alpha_max * 2 * pi * f * np.cos(2 * pi * f * ti + phi)
                h[i] = h0 + h_max * np.sin(2 * pi * f * ti)
                hd[i] = - h_max * 2 * pi * f * np.cos(2 * pi * f * ti)
            x[i] = x0 - U * ti
        self.alpha, self.alpha_dot, self.hpiv, self.h_dot, self.xpiv = al, ald, h, hd, x
        self.x_dot = -U * np.ones(nt)
        self.alpha_e = al - np.arctan2(hd, U)
        pa = np.zeros([nt, 2, self.Npoints])
        ax, ae = self.airfoil['x'], self.airfoil['eta']
        for i in range(nt):
            pa[i, 0, 0] = x[i] - self.piv * np.cos(-al[i])
            pa[i, 1, 0] = h[i] + self.piv * np.sin(-al[i])
            pa[i, 0, 1:] = pa[i, 0, 0] + np.cos(-al[i]) * ax[1:] - np.sin(-al[i]) * ae[1:]
            pa[i, 1, 1:] = pa[i, 1, 0] + np.sin(-al[i]) * ax[1:] + np.cos(-al[i]) * ae[1:]
        self.path = {'airfoil': pa,
                     'airfoil_gamma_points': pa[:, :, :-1] + self.xgamma * (pa[:, :, 1:] - pa[:, :, :-1])}

    # LUDVM.py:459-547 (loop form as in the reference; the one-argument arctan2 of LUDVM.py:520 is the documented fix)
    def motion_plunge(self, G=1, T=2, alpha_m=0, h0=0, x0=0.25):
        pi, U, nt = np.pi, self.Uinf, self.nt
        alpha_m = alpha_m * pi / 180
        Vmax = G * U
        T = T * self.chord / U
        self.G, self.T = G, T
        al, ald, h, hd, x = (np.zeros(nt) for _ in range(5))
        for i in range(nt):
            ti = self.t[i]
            al[i] = alpha_m
            if ti <= T:
                h[i] = h0 - Vmax * ti / 2 + Vmax * T / (4 * pi) * np.sin(2 * pi * ti / T)
                hd[i] = - Vmax * np.sin(pi * ti / T) ** 2
            else:
                h[i] = h[i - 1]
                hd[i] = 0
            x[i] = x0 - U * ti
        self.alpha, self.alpha_dot, self.hpiv, self.h_dot, self.xpiv = al, ald, h, hd, x
        self.x_dot = -U * np.ones(nt)
        self.alpha_e = al - np.arctan2(hd, U)
        pa = np.zeros([nt, 2, self.Npoints])
        ax, ae = self.airfoil['x'], self.airfoil['eta']
        for i in range(nt):
            pa[i, 0, 0] = x[i] - self.piv * np.cos(-al[i])
            pa[i, 1, 0] = h[i] + self.piv * np.sin(-al[i])
            pa[i, 0, 1:] = pa[i, 0, 0] + np.cos(-al[i]) * ax[1:] - np.sin(-al[i]) * ae[1:]
            pa[i, 1, 1:] = pa[i, 1, 0] + np.sin(-al[i]) * ax[1:] + np.cos(-al[i]) * ae[1:]
        self.path = {'airfoil': pa,
                     'airfoil_gamma_points': pa[:, :, :-1] + self.xgamma * (pa[:, :, 1:] - pa[:, :, :-1])}

    # LUDVM.py:572-595, expression order kept
    def airfoil_downwash(self, circulation, xw, zw, i):
        alpha, alpha_dot, h_dot = self.alpha[i], self.alpha_dot[i], self.h_dot[i]
        xp, zp = self.path['airfoil_gamma_points'][i, 0, :], self.path['airfoil_gamma_points'][i, 1, :]
        u1, w1 = self.induced_velocity(circulation, xw, zw, xp, zp)
        u = u1 * np.cos(alpha) - w1 * np.sin(alpha)
        w = u1 * np.sin(alpha) + w1 * np.cos(alpha)
        W = self.airfoil['detadx_panel'] * (self.Uinf * np.cos(alpha) + h_dot * np.sin(alpha) + u
                                            - alpha_dot * self.airfoil['eta_panel']) \
            - self.Uinf * np.sin(alpha) - alpha_dot * (self.airfoil['x_panel'] - self.piv) \
            + h_dot * np.cos(alpha) - w
        return W

    def tables(self):
        return tables_from(self)

    def time_loop(self, print_dt=50, BCcheck=False, tables=None):
        tb = tables if tables is not None else self.tables()
        self._tb = tb
        nt, P, Nc, nf, nv = tb['nt'], tb['P'], tb['Nc'], tb['nfree'], tb['nt'] - 1
        sin = _SimIn()
        for name, ctype in _SimIn._fields_:
            v = tb[name]
            setattr(sin, name, _p(v) if ctype is _dp else v)
        z = np.zeros
        out = dict(path_tev=z([nt, 2, nv]), path_lev=z([nt, 2, nv]), path_free=z([nt, 2, nf]),
                   g_tev=z(nv), g_lev=z(nv), g_bound=z(nv), g_airfoil=z([nv, P]), gamma_airfoil=z([nv, P]),
                   Gamma_airfoil=z([nv, P]), fourier=z([nt, 2, Nc]), lesp=z(nt), lesp_prev=z(nt),
                   lev_shed=z(nt), Fn=z(nt), Fs=z(nt), L=z(nt), D=z(nt), T=z(nt), M=z(nt))
        itev, ilev = C.c_long(0), C.c_long(0)
        sout = _SimOut()
        for name, ctype in _SimOut._fields_:
            if ctype is _dp:
                setattr(sout, name, _p(out[name]))
        sout.itev, sout.ilev = C.pointer(itev), C.pointer(ilev)
        rc = lib().oracle_sim_run(C.byref(sin), C.byref(sout))
        if rc != 0:
            raise RuntimeError("oracle_sim_run failed (%d)" % rc)
        self.path['TEV'], self.path['LEV'], self.path['FREE'] = out['path_tev'], out['path_lev'], out['path_free']
        self.circulation = {'TEV': out['g_tev'], 'LEV': out['g_lev'], 'FREE': self.circulation_freevort,
                            'bound': out['g_bound'], 'airfoil': out['g_airfoil'],
                            'gamma_airfoil': out['gamma_airfoil'], 'Gamma_airfoil': out['Gamma_airfoil'],
                            'IC': tb['ic']}
        self.fourier, self.LESP, self.LESP_prev, self.LEV_shed = out['fourier'], out['lesp'], out['lesp_prev'], out['lev_shed']
        self.Fn, self.Fs, self.L, self.D, self.T, self.M = (out[k] for k in ('Fn', 'Fs', 'L', 'D', 'T', 'M'))
        self.dp = np.zeros([nt, P])
        self.itev, self.ilev = itev.value, ilev.value
        self.BC = self._bc_check() if BCcheck else np.zeros([nv, self.Npoints])

    def _wake_before_convection(self, i, ilev):
        """(circulation, xw, zw) of the wake TEV[:itev+1] ++ LEV[:ilev+1] ++ FREE as it stood at step i BEFORE that step's
        convection (LUDVM.py:1095-1100): row i-1 of the stored paths plus the vortices placed in step i
        (LUDVM.py:672-681, :784-800).  Needs the full history."""
        itev = i - 1
        pa = self.path['airfoil']
        TEV, LEV, FREE = self.path['TEV'], self.path['LEV'], self.path['FREE']
        xT, zT = TEV[i - 1, 0, :itev + 1].copy(), TEV[i - 1, 1, :itev + 1].copy()
        if itev == 0:
            xT[0], zT[0] = pa[0, :, -1] + [0.5 * self.Uinf * self.dt, 0]
        else:
            xT[itev], zT[itev] = pa[i, :, -1] + 1 / 3 * (TEV[i - 1, :, itev - 1] - pa[i, :, -1])
        xL, zL = LEV[i - 1, 0, :ilev + 1].copy(), LEV[i - 1, 1, :ilev + 1].copy()
        cL = self.circulation['LEV'][:ilev + 1].copy()
        if self.LEV_shed[i] != -1:
            if ilev > 0 and self.LEV_shed[i - 1] != -1:
                xL[ilev], zL[ilev] = pa[i, :, 0] + 1 / 3 * (LEV[i - 1, :, ilev - 1] - pa[i, :, 0])
            else:
                xL[ilev], zL[ilev] = pa[i, :, 0]
        else:           # idle slot of row i: zero circulation at the origin
            xL[ilev], zL[ilev], cL[ilev] = 0.0, 0.0, 0.0
        ap = np.append
        return (ap(ap(self.circulation['TEV'][:itev + 1], cL), self.circulation['FREE']),
                ap(ap(xT, xL), FREE[i - 1, 0, :]), ap(ap(zT, zL), FREE[i - 1, 1, :]))

    def _bc_check(self):
        """Boundary-condition residual of LUDVM.py:1144-1161 (BCcheck=True), evaluated after the run from the stored
        history.  The reference raises there (it mixes `airfoil['x']`, length Npoints, with length-(Npoints-1) arrays
        at LUDVM.py:1153 and assigns Npoints-1 values to a row of Npoints at :1161); with `airfoil['x_panel']` in the
        first place, BC[itev, :Npoints-1] = BCnx + BCnz is the no-penetration residual, zero up to rounding; the last
        column keeps the reference's allocation (LUDVM.py:625) and stays zero."""
        nt, P = self.nt, self.Npoints - 1
        BC = np.zeros([nt - 1, self.Npoints])
        af, gpts = self.airfoil, self.path['airfoil_gamma_points']
        ilev = 0
        for i in range(1, nt):
            g, xw, zw = self._wake_before_convection(i, ilev)
            alpha, alpha_dot, h_dot = self.alpha[i], self.alpha_dot[i], self.h_dot[i]
            u1, w1 = self.induced_velocity(g, xw, zw, gpts[i, 0, :], gpts[i, 1, :])
            u = u1 * np.cos(alpha) - w1 * np.sin(alpha)
            w = u1 * np.sin(alpha) + w1 * np.cos(alpha)
            W = af['detadx_panel'] * (self.Uinf * np.cos(alpha) + h_dot * np.sin(alpha) + u
                                      - alpha_dot * af['eta_panel']) \
                - self.Uinf * np.sin(alpha) - alpha_dot * (af['x_panel'] - self.piv) \
                + h_dot * np.cos(alpha) - w
            BCnx = af['detadx_panel'] * (- u - self.Uinf * np.cos(alpha)
                                         - h_dot * np.sin(alpha) + alpha_dot * af['eta_panel'])
            BCnz = W + w + self.Uinf * np.sin(alpha) - h_dot * np.cos(alpha) \
                + alpha_dot * (af['x_panel'] - self.piv)
            BC[i - 1, :P] = BCnx + BCnz
            if self.LEV_shed[i] != -1:
                ilev += 1
        return BC

    def compute_coefficients(self):                                        # LUDVM.py:1173-1184
        q = 0.5 * self.rho * self.Uinf ** 2
        qc = q * self.chord
        self.Cp = self.dp / q
        self.Cn, self.Cs = self.Fn / qc, self.Fs / qc
        self.Cl, self.Cd, self.Ct = self.L / qc, self.D / qc, self.T / qc
        self.Cm = self.M / (qc * self.chord)

    def induced_velocity(self, circulation, xw, zw, xp, zp, viscous=True):
        return induced_velocity(circulation, xw, zw, xp, zp, self.v_core, viscous)

    def flowfield(self, xmin=-10, xmax=0, zmin=-4, zmax=4, dr=0.02, tsteps=(0, 1, 2)):
        """LUDVM.py:1186-1298, index quirks included."""
        x1, z1 = np.arange(xmin, xmax, dr), np.arange(zmin, zmax, dr)
        x, z = np.meshgrid(x1, z1, indexing='ij')
        xp, zp = np.ravel(x), np.ravel(z)
        u, w = np.zeros([len(tsteps), *x.shape]), np.zeros([len(tsteps), *x.shape])
        ap = np.append
        for ii, itev in enumerate(tsteps):
            if itev == 0:
                uu, ww = self.induced_velocity(self.circulation['FREE'], self.path['FREE'][0, 0], self.path['FREE'][0, 1], xp, zp)
            else:
                ilev = int(self.LEV_shed[itev])
                g = ap(ap(self.circulation['TEV'][:itev + 1], self.circulation['LEV'][:ilev + 1]), self.circulation['FREE'])
                xw = ap(ap(self.path['TEV'][itev - 1, 0, :itev + 1], self.path['LEV'][itev - 1, 0, :ilev + 1]), self.path['FREE'][itev, 0])
                zw = ap(ap(self.path['TEV'][itev - 1, 1, :itev + 1], self.path['LEV'][itev - 1, 1, :ilev + 1]), self.path['FREE'][itev, 1])
                gp = self.path['airfoil_gamma_points'][itev - 1]
                uw_, ww_ = self.induced_velocity(g, xw, zw, xp, zp)
                uf_, wf_ = self.induced_velocity(self.circulation['airfoil'][itev - 1], gp[0], gp[1], xp, zp)
                uu, ww = uw_ + uf_, ww_ + wf_
            u[ii], w[ii] = uu.reshape(x.shape), ww.reshape(x.shape)
        self.x_ff, self.z_ff, self.u_ff, self.w_ff = x, z, u, w
        self.ome_ff = vorticity(x, z, u, w)

"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference `/root/reference/LUDVM.py`.

The reference cannot travel to the GPU box (`/root/reference` does not exist there), so this module is
used only in the development container, by `tests/golden/make_golden.py` (which writes the committed
fixtures) and by the `not gpu` tests that pin `oracle/ludvm_oracle.c` against the real reference when it
is present.  Nothing in `ludvm_b200/` may import it.

Blockers to a plain `import LUDVM`:
  * module level `import matplotlib.pyplot`, `mpl.rc(...)`, `mpl.interactive(True)`  (LUDVM.py:3-4, 8-10)
  * `from airfoils import Airfoil` inside `airfoil_generation`                        (LUDVM.py:301-302)
Neither package is installed.  matplotlib is replaced by a MagicMock; `airfoils` by a stub that is only
valid for symmetric NACA 00xx sections, for which the camber line is identically zero and the package
contributes nothing numerically (LUDVM.py:328-340).
"""
import os
import sys
import types
import warnings
import contextlib
import io
from unittest.mock import MagicMock

import numpy as np

REFERENCE_DIR = os.environ.get("LUDVM_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "LUDVM.py"))


def _install_stubs():
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(m, MagicMock())
    if "airfoils" in sys.modules:
        return
    air = types.ModuleType("airfoils")
    fio = types.ModuleType("airfoils.fileio")

    class Airfoil:  # symmetric sections only
        def __init__(self, n):
            self._x_upper = self._x_lower = np.linspace(0, 1, n)
            self._y_upper = self._y_lower = np.zeros(n)
            self.all_points = None

        @classmethod
        def NACA4(cls, naca, n_points=200):
            if naca[:2] != "00":
                raise ValueError("stub airfoils package: cambered sections are parity-unpinned")
            return cls(n_points)

        def camber_line(self, x):
            return np.zeros_like(x)

        def camber_line_angle(self, x):
            return np.zeros_like(x)

    air.Airfoil = Airfoil
    fio.import_airfoil_data = None
    air.fileio = fio
    sys.modules["airfoils"] = air
    sys.modules["airfoils.fileio"] = fio


_REF = None


def load():
    """Return the reference module (imported once)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise FileNotFoundError("reference not present at %s" % REFERENCE_DIR)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    _install_stubs()
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import LUDVM as ref  # noqa: N811
    finally:
        sys.path.pop(0)
    _REF = ref
    return ref


def run(**kw):
    """Run the reference constructor (= the whole simulation, LUDVM.py:282-295) silently."""
    ref = load()
    kw.setdefault("verbose", False)
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.LUDVM(**kw)


def bare(v_core):
    """An un-constructed reference object that can only evaluate `induced_velocity` (LUDVM.py:549-570)."""
    ref = load()
    o = object.__new__(ref.LUDVM)
    o.v_core = v_core
    return o


README_KW = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30,
                 LESPcrit=0.2, Naca="0012")  # README.md:27-28

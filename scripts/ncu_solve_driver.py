"""Driver for `ncu -k regex:k_solve`: the README case stepped with individually launched kernels (no graph)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import LUDVM, _lib
L = _lib.load()
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
s = LUDVM(**README, verbose=False, run=False, mode=mode, store_history=False)
s.time_loop(nsteps=0)
ms = (C.c_double * 5)()
_lib.check(L.ludvm_sim_profile_steps(s._sim, int(sys.argv[2]) if len(sys.argv) > 2 else 320, ms))
print([v for v in ms])

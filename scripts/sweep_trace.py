"""Phase shares of the one-CTA-per-case driver (parameter sweeps) from the clock64 traces of a -DLUDVM_TRACE build:
cycles summed over all cases, per phase.  Usage: LUDVM_B200_LIB=scripts/_build/libludvm_trace.so python scripts/sweep_trace.py [mode] [ncases]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib, sweep
L = _lib.load()
L.ludvm_debug_trace.restype, L.ludvm_debug_trace.argtypes = C.c_int, [C.POINTER(C.c_longlong)]
mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 444
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
cases = sweep.lespcrit_k_grid(np.linspace(0.1, 0.4, 64), np.linspace(0.1, 1.0, 64), **README)[::max(1, 4096 // n)][:n]
sweep.run_sweep(cases[:4], mode=mode)
tr0 = (C.c_longlong * 64)(); L.ludvm_debug_trace(tr0)
t = time.perf_counter(); r = sweep.run_sweep(cases, mode=mode); dt = time.perf_counter() - t
tr = (C.c_longlong * 64)(); L.ludvm_debug_trace(tr)
names = ["wake-on-foil", "solve", "convection", "loads", "update + cumulative sums", "step begin (barrier)"]
acc = [tr[40 + q] - tr0[40 + q] for q in range(6)]
tot = float(sum(acc))
print("%s, %d cases: ludvm_sweep_run %.3f s; per case-step %.1f us of CTA time (CTAs share an SM)" %
      (mode, len(cases), r["timing"]["ludvm_sweep_run_s"], tot / len(cases) / 400 / 1965.0))
for nm, a in zip(names, acc):
    print("  %-28s %6.2f us per case-step  %5.1f %%" % (nm, a / len(cases) / 400 / 1965.0, 100 * a / tot))

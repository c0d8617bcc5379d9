"""clock64 trace of the solve kernel's phases (library built with -DLUDVM_TRACE into scripts/_build/libludvm_trace.so)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libludvm_trace.so")
from ludvm_b200 import LUDVM
L = _lib.load()
L.ludvm_debug_trace.restype = C.c_int
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
names = {0: "start", 1: "tables+placement", 2: "fold", 3: "sum TEV", 4: "sum LEV", 5: "T1,T2", 6: "I,J trapz", 7: "W", 8: "fourier",
         9: "lesp/shed branch", 10: "commit+bound", 11: "Gamma cumsum"}
for kw, mode, pts in ((README, "exact", [100, 300]), (README, "fast", [300]), (dict(README, dt=2e-3, tf=40), "fast", [4000, 16000])):
    s = LUDVM(**kw, verbose=False, run=False, mode=mode, store_history=False)
    s.time_loop(nsteps=0)
    done = 0
    ms = (C.c_double * 5)()
    for p in pts:
        _lib.check(L.ludvm_sim_run(s._sim, p - done)); done = p
        s.ctx.synchronize()
        acc = np.zeros(64)
        reps = 10
        for _ in range(reps):
            _lib.check(L.ludvm_sim_profile_steps(s._sim, 1, ms)); done += 1
            tr = (C.c_longlong * 64)()
            assert L.ludvm_debug_trace(tr) == 0
            t = np.array(tr[:], dtype=np.float64)
            t[3] = t[2]  # circulation sums are evaluated beside phase 1 on the graph path
            acc[:11] += np.concatenate([[0], np.diff(t[:11])])
            acc[20] += t[21] - t[20]; acc[21] += t[0] - t[21]; acc[22] += t[22] - t[10]; acc[23] += t[22] - t[20]
        acc /= reps
        print(mode, "step", p, " ".join("%s=%.0f" % (names[k], acc[k]) for k in range(1, 11)),
              "| step_begin=%.0f pre=%.0f exit=%.0f total_cycles=%.0f (%.1f us @1.965GHz)" % (acc[20], acc[21], acc[22], acc[23], acc[23] / 1965.0), flush=True)
    s.close()

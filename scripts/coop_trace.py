"""Phase breakdown of the persistent cooperative step (README case) from the clock64 traces of a -DLUDVM_TRACE build:
cycles of CTA 0 in each phase / barrier, averaged over the run.  Build the traced library first:
  nvcc ... -DLUDVM_TRACE -fmad=false -c ludvm_b200/csrc/sim.cu -o scripts/_build/sim_trace.o && link with the other objects
Usage: LUDVM_B200_LIB=scripts/_build/libludvm_trace.so python scripts/coop_trace.py [mode]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import LUDVM, _lib
L = _lib.load()
L.ludvm_debug_trace.restype, L.ludvm_debug_trace.argtypes = C.c_int, [C.POINTER(C.c_longlong)]
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
TF = float(os.environ.get("PROBE_TF", "20"))
NS = int(round(TF / 5e-2))
README = dict(t0=0, tf=TF, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
names = ["phase1 wake-on-foil", "barrier 1", "solve (CTA 0)", "barrier 2", "phase3 conv partials", "barrier 3",
         "phase4 loads (CTA 0)", "barrier 4"]
for rep in range(2):
    s = LUDVM(**README, verbose=False, run=False, mode=mode, store_history=False, steps_per_graph=400)
    s.time_loop(nsteps=0); s.ctx.synchronize()
    t = time.perf_counter(); _lib.check(L.ludvm_sim_run(s._sim, NS)); s.ctx.synchronize(); dt = time.perf_counter() - t
    tr = (C.c_longlong * 64)()
    assert L.ludvm_debug_trace(tr) == 0
    s.close()
acc = [tr[40 + q] / float(NS) for q in range(8)]
print("mode %s: %.2f us/step wall; CTA-0 cycles per step by phase (sum %.0f):" % (mode, dt / NS * 1e6, sum(acc)))
for n, a in zip(names, acc):
    print("  %-24s %8.0f cycles  %6.2f us @1.965 GHz" % (n, a, a / 1965.0))
sol = [tr[k] for k in range(11)]
print("  solve sub-phases (last step, cycles from entry):", [int(sol[k] - sol[0]) for k in range(1, 11)])
print("  inside the LAST block_fold / block_trapz call of CTA 0 (cycles between trace points 30..35 / 36..39):",
      [int(tr[k + 1] - tr[k]) for k in range(30, 35)], [int(tr[k + 1] - tr[k]) for k in range(36, 39)])

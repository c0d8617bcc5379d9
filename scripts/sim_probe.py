"""Time the README case (400 steps) on the device, exact and fast, kernel-only (graph replay)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from ludvm_b200 import LUDVM
kw = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
for mode in ("exact", "fast"):
    for K in (50, 400):
        s = LUDVM(**kw, verbose=False, run=False, mode=mode, steps_per_graph=K)
        best = 1e9
        for rep in range(3):
            t = time.perf_counter(); s.time_loop(); s.compute_coefficients(); dt = time.perf_counter() - t
            best = min(best, dt)
        print(json.dumps({"mode": mode, "K": K, "s": best, "steps_per_s": 400 / best, "Cl_end": float(s.Cl[-1]), "ilev": s.ilev}), flush=True)

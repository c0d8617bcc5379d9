"""Time the dt = 2e-3 high-resolution run (BASELINE.json configs[1]) or a prefix of it, fast or exact mode, for A/B
comparisons of the convection kernels (environment knobs are read by the library at graph-build time).
Usage: python scripts/hires_probe.py [--steps 20000] [--mode fast]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ludvm_b200 import LUDVM
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20000)
ap.add_argument("--mode", default="fast")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
kw = dict(t0=0, tf=40, dt=2e-3, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
s = LUDVM(**kw, verbose=False, run=False, mode=a.mode, store_history=False)
tb = s.step_tables()
for rep in range(a.reps + 1):
    t0 = time.perf_counter(); s.time_loop(tables=tb, nsteps=a.steps if rep else min(a.steps, 300)); dt = time.perf_counter() - t0
    if rep:
        env = {k: v for k, v in os.environ.items() if k.startswith("LUDVM_")}
        print(json.dumps({"mode": a.mode, "steps": a.steps, "seconds": dt, "steps_per_s": a.steps / dt, "env": env,
                          "L_last": float(s.L[a.steps]), "ilev": s.ilev, "range_proof_held": s.range_proof_held}), flush=True)

"""Time the batched sweep (BASELINE.json configs[3]) at a few sizes."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from ludvm_b200 import sweep
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for mode in ("exact", "fast"):
    cases = sweep.lespcrit_k_grid(np.linspace(0.1, 0.4, n), np.linspace(0.1, 1.0, n), **README)
    t = time.perf_counter(); res = sweep.run_sweep(cases, mode=mode); dt = time.perf_counter() - t
    t = time.perf_counter(); res = sweep.run_sweep(cases, mode=mode); dt2 = time.perf_counter() - t
    print(json.dumps({"cases": len(cases), "mode": mode, "s_first": dt, "s": dt2, "case_steps_per_s": len(cases) * 400 / dt2,
                      "levs_mean": float((res["LEV_shed"] != -1).sum(1).mean())}), flush=True)

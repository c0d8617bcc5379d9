"""Per-phase timing of the on-device time step (ludvm_sim_profile_steps: the graph path's kernels launched one by
one with CUDA events) at chosen wake sizes, next to the graph-replay time of the same steps.
Usage: python scripts/step_profile.py [--out gpurun_out/step_profile.json]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import LUDVM, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/step_profile.json")
ap.add_argument("--hires-points", default="500,1000,2000,3000,4000,6000,8000,12000,16000,19800")
ap.add_argument("--nprof", type=int, default=20)
args = ap.parse_args()
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
L = _lib.load()
rep = {}


def profile(kw, mode, points, nprof, K=50):
    s = LUDVM(**kw, verbose=False, run=False, mode=mode, store_history=False, steps_per_graph=K)
    s.time_loop(nsteps=0)
    sim, rows, done = s._sim, [], 0
    ms = (C.c_double * 5)()
    for p in points:
        if p > done:
            _lib.check(L.ludvm_sim_run(sim, p - done)); done = p
        s.ctx.synchronize()
        _lib.check(L.ludvm_sim_profile_steps(sim, nprof, ms)); done += nprof
        prof = [v / nprof * 1e3 for v in ms]
        s.ctx.synchronize()
        t = time.perf_counter(); _lib.check(L.ludvm_sim_run(sim, K)); s.ctx.synchronize(); g_us = (time.perf_counter() - t) / K * 1e6
        done += K
        cnt = np.empty(4, dtype=np.int64)
        _lib.check(L.ludvm_sim_fetch(sim, _lib.FIELDS["COUNTERS"], cnt.ctypes.data, cnt.nbytes))
        nw = int(cnt[1] + cnt[2] + 3)
        rows.append({"step": p, "wake": nw, "us_wake_on_foil": prof[0], "us_solve": prof[1], "us_conv": prof[2],
                     "us_finish": prof[3], "us_sum_events": prof[4], "us_per_step_graph": g_us,
                     "pairs_per_s_graph": (80.0 + nw) * nw / (g_us * 1e-6)})
        print(mode, rows[-1], flush=True)
    s.close()
    return rows


rep["readme_exact"] = profile(README, "exact", [50, 200, 300], args.nprof, K=50)
rep["readme_fast"] = profile(README, "fast", [50, 200, 300], args.nprof, K=50)
pts = [int(x) for x in args.hires_points.split(",")]
hi = dict(README, dt=2e-3, tf=40)
rep["hires_fast"] = profile(hi, "fast", pts, args.nprof)
rep["hires_exact"] = profile(hi, "exact", [p for p in pts if p <= 6000], args.nprof)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(rep, open(args.out, "w"), indent=1)

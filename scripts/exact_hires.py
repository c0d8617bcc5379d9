"""Exact-mode config-2-like run (dt=2e-3): time for N steps, and parity of a prefix against the tiled-off path."""
import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from ludvm_b200 import LUDVM
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
kw = dict(t0=0, tf=steps * 2e-3, dt=2e-3, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
for rep in range(2):
    t = time.perf_counter()
    s = LUDVM(**kw, verbose=False, mode="exact", store_history=False)
    dtm = time.perf_counter() - t
print("exact hires %d steps: %.3f s, itev %d ilev %d, L[-1] %r" % (steps, dtm, s.itev, s.ilev, s.L[-1]))

"""Driver for `ncu -k regex:k_conv_partials_sk`: a fast-mode run whose wake is large from step 0 (free vortices)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ludvm_b200 import LUDVM
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 22000
rng = np.random.default_rng(9)
xy = np.stack([rng.uniform(-20.0, -0.5, nf), rng.uniform(-2.0, 2.0, nf)])
gam = rng.standard_normal(nf) * 1e-3
kw = dict(t0=0, tf=0.5, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012",
          circulation_freevort=gam, xy_freevort=xy)
t = time.perf_counter()
f = LUDVM(**kw, verbose=False, mode="fast", store_history=False)
print("nf", nf, "steps", f.steps_done, "wall", time.perf_counter() - t, "L[-1]", f.L[-1])

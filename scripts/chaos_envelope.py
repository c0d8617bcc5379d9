"""The honest yardstick for any non-bit-exact arithmetic: how fast does the REFERENCE ALGORITHM
ITSELF (CPU oracle, bit-equal to the reference) diverge when one input is perturbed by 1 ulp?  CPU only."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ludvm_oracle as O

README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
eps = np.nextafter(1.0, 2.0)
for name, kw, n, marks in (("README dt=5e-2", README, 400, (50, 100, 200, 300, 400)),
                           ("hi-res dt=2e-3", dict(README, dt=2e-3, tf=40), 1500, (200, 500, 700, 800, 900, 1000, 1500))):
    a = O.OracleLUDVM(**kw, nsteps=n)
    b = O.OracleLUDVM(**dict(kw, h_max=eps), nsteps=n)
    sc = np.max(np.abs(a.L[:n + 1]))
    print(name, "| LEVs", int((a.LEV_shed != -1).sum()), "| first LEV at step", int(np.argmax(a.LEV_shed != -1)))
    for m in marks:
        print("   up to step %5d: max|dL|/max|L| = %.3e, LEV_shed pattern equal: %s"
              % (m, np.max(np.abs(a.L[:m + 1] - b.L[:m + 1])) / sc, np.array_equal(a.LEV_shed[:m + 1], b.LEV_shed[:m + 1])))

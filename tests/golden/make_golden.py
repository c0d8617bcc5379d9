"""Generate the committed golden fixtures from the UNMODIFIED reference (/root/reference/LUDVM.py).

Runs only in the development container (the reference does not travel to the GPU box).  Usage:

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Every simulation fixture holds
  * `kw`      -- the constructor kwargs (JSON) given to the reference,
  * `tb_*`    -- every host-evaluated input table the step consumes (so the device/oracle step can be
                 re-run from bit-identical inputs even if the box's libm rounds a cos() differently),
  * outputs   -- load/coefficient histories, Fourier coefficients, circulations, LESP, LEV_shed in full;
                 vortex path histories as a few full rows plus a sha256 of the raw bytes.
The reference has no tests or golden files of its own, so these are the pins.
"""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as R  # noqa: E402
from oracle.ludvm_oracle import tables_from  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PATH_ROWS = {"readme": [1, 2, 27, 100, 200, 400], "default": None}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sim_fixture(name, kw, rows=None, flowfield=None, level="full"):
    kwj = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    r = R.run(**kw)
    tb = tables_from(r)
    d = {"kw": np.array(json.dumps(kwj))}
    if level == "full":
        for k, v in tb.items():
            d["tb_" + k] = np.asarray(v)
    for k in ("Cl", "Cd", "Cm", "Cn", "Cs", "Ct", "Fn", "Fs", "L", "D", "T", "M", "LESP", "LESP_prev",
              "LEV_shed", "fourier", "alpha", "alpha_dot", "h_dot"):
        d[k] = np.asarray(getattr(r, k), dtype=np.float64)
    for k in ("TEV", "LEV", "bound"):
        d["circ_" + k] = np.asarray(r.circulation[k], dtype=np.float64)
    d["itev_ilev"] = np.array([r.itev, r.ilev])
    rows = rows if rows is not None else list(range(r.nt))
    crow = [x - 1 for x in rows if x >= 1]
    d["circ_rows"] = np.array(crow)
    for k in ("airfoil", "gamma_airfoil", "Gamma_airfoil"):
        d["circ_" + k + "_rows"] = np.ascontiguousarray(r.circulation[k][crow])
        d["circ_" + k + "_sha256"] = np.array(sha(r.circulation[k]))
    if level != "full":
        rows = []
    d["path_rows"] = np.array(rows)
    for k in ("TEV", "LEV", "FREE"):
        d["path_" + k + "_rows"] = np.ascontiguousarray(r.path[k][rows])
        d["path_" + k + "_sha256"] = np.array(sha(r.path[k]))
    if flowfield is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            r.flowfield(**flowfield)
        d["ff_kw"] = np.array(json.dumps(flowfield))
        for k in ("x_ff", "z_ff", "u_ff", "w_ff", "ome_ff"):
            d[k] = getattr(r, k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-18s nt=%d ilev=%d Cl[-1]=%r  %d KB" % (name, r.nt, r.ilev, float(r.Cl[-1]),
                                                 os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024))
    return r


def percall_fixture():
    """Per-call `induced_velocity` vectors (LUDVM.py:549-570) on seeded clouds around numpy's
    pairwise-summation block boundaries (8, 128, the n//2 - n//2%8 split)."""
    rng = np.random.default_rng(20260101)
    o = R.bare(0.065)
    d, cases = {}, []
    sizes = [(5, 1), (16, 7), (16, 8), (16, 9), (3, 127), (16, 128), (16, 129), (9, 136), (16, 255), (16, 256),
             (16, 257), (33, 602), (12, 1000), (12, 1025), (80, 2048), (7, 5000), (300, 300), (1, 1)]
    for c, (npnt, nw) in enumerate(sizes):
        xw, zw = rng.uniform(-20, 0, nw), rng.uniform(-4, 4, nw)
        g = rng.standard_normal(nw) * 1e-2
        if npnt == nw:   # self-interaction (diagonal terms are exactly zero with the viscous core)
            xp, zp = xw.copy(), zw.copy()
        else:
            xp, zp = rng.uniform(-20, 0, npnt), rng.uniform(-4, 4, npnt)
        u, w = o.induced_velocity(g, xw, zw, xp, zp)
        for k, v in (("g", g), ("xw", xw), ("zw", zw), ("xp", xp), ("zp", zp), ("u", u), ("w", w)):
            d["c%d_%s" % (c, k)] = v
        cases.append(c)
    # unit-strength broadcast (LUDVM.py:751) and viscous=False
    xp, zp = rng.uniform(-1, 1, 80), rng.uniform(-1, 1, 80)
    u, w = o.induced_velocity(np.array([1]), np.array([0.3]), np.array([-0.2]), xp, zp)
    d.update(b_xp=xp, b_zp=zp, b_u=u, b_w=w)
    xw, zw, g = rng.uniform(-1, 1, 50), rng.uniform(-1, 1, 50), rng.standard_normal(50)
    u, w = o.induced_velocity(g, xw, zw, xp, zp, viscous=False)
    d.update(i_xw=xw, i_zw=zw, i_g=g, i_u=u, i_w=w)
    d["ncases"] = np.array(len(cases))
    d["v_core"] = np.array(0.065)
    np.savez_compressed(os.path.join(OUT, "percall.npz"), **d)
    print("percall            %d cases  %d KB" % (len(cases), os.path.getsize(os.path.join(OUT, "percall.npz")) // 1024))


def main():
    ref = R.load()
    percall_fixture()
    sim_fixture("readme", dict(R.README_KW), rows=PATH_ROWS["readme"])
    sim_fixture("ramesh_tf2", dict(R.README_KW, tf=2, method="Ramesh", LESPcrit=0.1))
    xy, g = ref.generate_free_single_vortex()
    sim_fixture("freevort_tf3", dict(R.README_KW, tf=3, circulation_freevort=g, xy_freevort=xy.T),
                flowfield=dict(xmin=-3.0, xmax=0.5, zmin=-1.5, zmax=1.0, dr=0.05, tsteps=[0, 10, 40]))
    sim_fixture("hires_200", dict(R.README_KW, dt=2e-3, tf=0.4), rows=[1, 100, 200])
    # sweep samples (BASELINE.json config 4 corner/interior cases); only histories are kept
    for a, (lc, k) in enumerate([(0.1, 0.1), (0.1, 1.0), (0.4, 0.1), (0.4, 1.0), (0.25, 0.55)]):
        sim_fixture("sweep_%d" % a, dict(R.README_KW, LESPcrit=lc, k=k), rows=[400], level="hist")


if __name__ == "__main__":
    main()

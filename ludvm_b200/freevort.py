"""Free-vortex cloud builders (the reference's module-level helpers, LUDVM.py:18-130): inputs for
`LUDVM(circulation_freevort=..., xy_freevort=...)`.  Host-side numpy, not on the step path.

Return conventions follow the reference: `xyvorts` is [n, 2] (pass `xyvorts.T` as `xy_freevort`), `gammavorts` [n].
Quirk kept for parity: the circulation of group n is divided by the CUMULATIVE number of point
vortices generated so far, not by the size of the group (LUDVM.py:46)."""
import numpy as np


def _ring_points(radii, counts):
    """Concentric rings: counts[i] points equally spaced on the circle of radius radii[i] (LUDVM.py:27-35)."""
    xs, ys = [], []
    for r, n in zip(radii, counts):
        th = np.linspace(0, 2 * np.pi, n, endpoint=False)
        xs.append(r * np.cos(th))
        ys.append(r * np.sin(th))
    return np.concatenate(xs), np.concatenate(ys)


def generate_free_vortices(nvorts, cvorts, vortradius, layerspervort, npervortlayer, gammapervort):
    """`nvorts` clouds of point vortices centred at `cvorts[n]`, each made of `layerspervort` rings (LUDVM.py:18-51)."""
    radii = vortradius * np.linspace(0, 1, layerspervort)
    rx, ry = _ring_points(radii, npervortlayer)
    xv, yv, gv = np.empty(0), np.empty(0), np.empty(0)
    for n in range(nvorts):
        xv = np.append(xv, rx + cvorts[n, 0])
        yv = np.append(yv, ry + cvorts[n, 1])
        gv = np.append(gv, gammapervort[n] / len(xv) * np.ones(len(rx)))
    return np.stack([xv, yv], axis=1), gv


def generate_free_single_vortex():
    """One cloud of 61 point vortices, circulation 10, centred at (-2.5, -0.5) (LUDVM.py:53-71)."""
    cv = np.array([[-2.5, -0.5]])
    return generate_free_vortices(1, cv, 0.5, 5, 1 * np.array([1, 5, 10, 15, 30]), 10 * np.array([1]))


def generate_flowfield_vortices(vortex_radius=0.2, gamma=0.5, xmin=-5, xmax=0, ymin=-3, ymax=2.5, layerspervort=2,
                                npervortlayer=np.array([1, 5]), centers_separation_factor=1):
    """Taylor-Green-like lattice of counter-rotating clouds (LUDVM.py:73-96)."""
    d = centers_separation_factor * 2 * vortex_radius
    cx = np.arange(xmin + vortex_radius, xmax - vortex_radius + d, d)
    cy = np.arange(ymin + vortex_radius, ymax - vortex_radius + d, d)
    cxv, cyv = np.meshgrid(cx, cy, indexing='ij')
    cv = np.stack([np.ravel(cxv), np.ravel(cyv)], axis=1)
    sign_i = np.where(np.arange(len(cx)) % 2 == 0, 1.0, -1.0)[:, None]
    sign_j = np.where(np.arange(len(cy)) % 2 == 0, 1.0, -1.0)[None, :]
    gammas = gamma * sign_i * sign_j
    return generate_free_vortices(cv.shape[0], cv, vortex_radius, layerspervort, npervortlayer, np.ravel(gammas))


def generate_flowfield_turbulence(vortex_radius=0.2, vortex_density=0.8, gamma=0.5, xmin=-5, xmax=0, ymin=-3, ymax=2.5,
                                  layerspervort=2, npervortlayer=np.array([1, 5]), overlap=False, rng=None):
    """Random field of non-overlapping clouds (LUDVM.py:99-130).  The reference draws from the unseeded global
    `np.random`; pass `rng` (a `numpy.random.Generator`) for reproducible fields."""
    rs = rng if rng is not None else np.random.default_rng()
    area = (xmax - xmin) * (ymax - ymin)
    nv = int(vortex_density * area / (np.pi * vortex_radius ** 2))
    gpv = gamma * rs.choice([-1, 1], nv)
    dmin = 2 * vortex_radius
    cv = np.stack([rs.uniform(xmin, xmax, nv), rs.uniform(ymin, ymax, nv)], axis=1)
    for n in range(1, nv):
        tries = 0
        if overlap:
            cv[n] = rs.uniform(xmin, xmax), rs.uniform(ymin, ymax)
            continue
        while np.any(np.hypot(cv[:n, 0] - cv[n, 0], cv[:n, 1] - cv[n, 1]) < dmin) and tries < 20000:
            cv[n] = rs.uniform(xmin, xmax), rs.uniform(ymin, ymax)
            tries += 1
        if tries == 20000:
            print('VortexError: Cannot locate more vortices with the actual radius and separation')
    return generate_free_vortices(nv, cv, vortex_radius, layerspervort, npervortlayer, gpv)

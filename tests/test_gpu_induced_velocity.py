"""GPU parity tests of the all-pairs Biot-Savart kernels (C ABI `ludvm_induced_velocity`, replacing
LUDVM.induced_velocity, LUDVM.py:549-570) against the golden vectors and the CPU oracle."""
import numpy as np
import pytest

from conftest import biteq, load_golden

pytestmark = pytest.mark.gpu

VC = 0.065


@pytest.fixture(scope="module")
def ops():
    from ludvm_b200 import ops as o
    return o


def cloud(rng, n):
    return rng.standard_normal(n) * 1e-2, rng.uniform(-20, 0, n), rng.uniform(-4, 4, n)


def cond_bound(g, xw, zw, xp, zp, vc):
    """sum_j |Gamma_j K_ij| per target for both components: the condition-aware error scale."""
    dx, dz = xp[:, None] - xw[None, :], zp[:, None] - zw[None, :]
    k = np.abs(g)[None, :] / (2 * np.pi * np.sqrt((dx * dx + dz * dz) ** 2 + vc ** 4))
    return (k * np.abs(dz)).sum(1), (k * np.abs(dx)).sum(1)


def test_exact_golden_percall(ops):
    g = load_golden("percall")
    vc = float(g["v_core"])
    for c in range(int(g["ncases"])):
        u, w = ops.induced_velocity(g["c%d_g" % c], g["c%d_xw" % c], g["c%d_zw" % c], g["c%d_xp" % c],
                                    g["c%d_zp" % c], vc, mode="exact")
        assert biteq(u, g["c%d_u" % c]) and biteq(w, g["c%d_w" % c]), "case %d" % c
    u, w = ops.induced_velocity(np.array([1]), np.array([0.3]), np.array([-0.2]), g["b_xp"], g["b_zp"], vc)
    assert biteq(u, g["b_u"]) and biteq(w, g["b_w"])
    u, w = ops.induced_velocity(g["i_g"], g["i_xw"], g["i_zw"], g["b_xp"], g["b_zp"], vc, viscous=False)
    assert biteq(u, g["i_u"]) and biteq(w, g["i_w"])


@pytest.mark.parametrize("npnt,nw", [(1, 1), (3, 5), (80, 7), (80, 8), (80, 9), (5, 127), (80, 128), (80, 129),
                                     (80, 137), (81, 1000), (80, 4097), (1, 20001), (80, 40003), (602, 602),
                                     (2049, 2049), (5000, 3333), (9000, 700), (8192, 2000),
                                     # one-thread-per-row tiled exact kernel (>= 4096 rows, >= 1024 sources): self
                                     # terms (library fallback), ragged leaves, several staged tiles, cut tree
                                     (4096, 4096), (4100, 1025), (6000, 9999), (4097, 1031), (20000, 2500)])
def test_exact_vs_oracle_bitwise(ops, oracle, npnt, nw):
    rng = np.random.default_rng(npnt * 100003 + nw)
    g, xw, zw = cloud(rng, nw)
    if npnt == nw:
        xp, zp = xw, zw
    else:
        xp, zp = rng.uniform(-20, 0, npnt), rng.uniform(-4, 4, npnt)
    uo, wo = oracle.induced_velocity(g, xw, zw, xp, zp, VC)
    u, w = ops.induced_velocity(g, xw, zw, xp, zp, VC, mode="exact")
    assert biteq(u, uo) and biteq(w, wo)


@pytest.mark.parametrize("npnt,nw", [(80, 1), (80, 600), (80, 40000), (602, 602), (4096, 4096), (20000, 5000),
                                     (9001, 12345), (300, 100000)])
def test_fast_f64_within_1e12(ops, oracle, npnt, nw):
    """BASELINE.json: per-call induced velocities within 1e-12 relative in fp64 (relative to sum |terms|)."""
    rng = np.random.default_rng(npnt * 7 + nw)
    g, xw, zw = cloud(rng, nw)
    xp, zp = (xw, zw) if npnt == nw else (rng.uniform(-20, 0, npnt), rng.uniform(-4, 4, npnt))
    uo, wo = oracle.induced_velocity(g, xw, zw, xp, zp, VC)
    u, w = ops.induced_velocity(g, xw, zw, xp, zp, VC, mode="fast")
    sel = slice(0, min(npnt, 256))
    bu, bw = cond_bound(g, xw, zw, xp[sel], zp[sel], VC)
    assert np.all(np.abs(u - uo)[sel] <= 1e-12 * bu + 1e-300)
    assert np.all(np.abs(w - wo)[sel] <= 1e-12 * bw + 1e-300)
    assert np.max(np.abs(u - uo)) <= 1e-12 * np.max(np.abs(uo))
    assert np.max(np.abs(w - wo)) <= 1e-12 * np.max(np.abs(wo))


def test_fast_f32_reported_accuracy(ops, oracle):
    rng = np.random.default_rng(9)
    g, xw, zw = cloud(rng, 20000)
    uo, wo = oracle.induced_velocity(g, xw, zw, xw, zw, VC)
    u, w = ops.induced_velocity(g, xw, zw, xw, zw, VC, mode="fp32")
    err = max(np.max(np.abs(u - uo)) / np.max(np.abs(uo)), np.max(np.abs(w - wo)) / np.max(np.abs(wo)))
    assert err < 5e-4   # fp32 pair arithmetic on O(20) coordinates: reported, not a parity claim


def test_per_source_core_and_empty(ops, oracle):
    rng = np.random.default_rng(4)
    g, xw, zw = cloud(rng, 300)
    xp, zp = rng.uniform(-20, 0, 17), rng.uniform(-4, 4, 17)
    vcs = np.full(300, VC ** 4)
    u1, w1 = ops.induced_velocity(g, xw, zw, xp, zp, VC, mode="exact", vc4_per_source=vcs)
    u2, w2 = oracle.induced_velocity(g, xw, zw, xp, zp, VC)
    assert biteq(u1, u2) and biteq(w1, w2)
    g, xw, zw = cloud(rng, 1500)   # tiled exact kernel with a per-source core
    xp2, zp2 = rng.uniform(-20, 0, 4500), rng.uniform(-4, 4, 4500)
    u1, w1 = ops.induced_velocity(g, xw, zw, xp2, zp2, VC, mode="exact", vc4_per_source=np.full(1500, VC ** 4))
    u2, w2 = oracle.induced_velocity(g, xw, zw, xp2, zp2, VC)
    assert biteq(u1, u2) and biteq(w1, w2)
    u, w = ops.induced_velocity(np.zeros(0), np.zeros(0), np.zeros(0), xp, zp, VC)
    assert np.all(u == 0) and np.all(w == 0)
    u, w = ops.induced_velocity(g, xw, zw, np.zeros(0), np.zeros(0), VC)
    assert u.size == 0
    with pytest.raises(ValueError):
        ops.induced_velocity(g[:5], xw, zw, xp, zp, VC)


def test_linearity_and_self_term(ops):
    """Size-independent properties at a larger size: linear in Gamma; the viscous self-term is exactly zero."""
    rng = np.random.default_rng(12)
    g, xw, zw = cloud(rng, 30000)
    u1, w1 = ops.induced_velocity(g, xw, zw, xw[:512], zw[:512], VC, mode="fast")
    u2, w2 = ops.induced_velocity(2.0 * g, xw, zw, xw[:512], zw[:512], VC, mode="fast")
    assert np.array_equal(u2, 2.0 * u1) and np.array_equal(w2, 2.0 * w1)   # scaling by 2 is exact in binary fp
    u, w = ops.induced_velocity(np.array([3.0]), np.array([1.5]), np.array([-2.5]), np.array([1.5]), np.array([-2.5]), VC)
    assert u[0] == 0.0 and w[0] == 0.0


def _plan():
    from ludvm_b200 import _lib
    return _lib.default_context().last_plan()


@pytest.mark.parametrize("npnt,nw", [(80, 700), (300, 4100), (4500, 1500), (8200, 5000)])
def test_exact_range_proof_and_its_fallback(ops, oracle, npnt, nw, monkeypatch):
    """The exact kernels drop the per-pair range words when a scan of the coordinate arrays proves them redundant
    (common.cuh: coord_in_safe_window).  Operands the proof does not cover must still reach the flagged
    instantiation and, through it, the library's __ddiv_rn / __dsqrt_rn: results stay bit-equal to the oracle for
      * ordinary clouds (flag-free), with many exactly coincident coordinates (zero numerators) and signed zeros;
      * coordinates of 1e-300, subnormals, 1e+200 (out of the window: flagged instantiation, library fallback);
      * viscous=False (vc^4 = 0: no proof attempted), coincident points included (0/0 = nan like numpy)."""
    rng = np.random.default_rng(npnt + nw)
    g, xw, zw = cloud(rng, nw)
    xp, zp = rng.uniform(-20, 0, npnt), rng.uniform(-4, 4, npnt)
    # coincidences and signed zeros
    xp[::3], zp[::5] = xw[(np.arange(0, npnt, 3) * 7) % nw], zw[(np.arange(0, npnt, 5) * 11) % nw]
    xp[1], zp[1], xw[2], zw[2], xw[3], zw[4] = 0.0, -0.0, -0.0, 0.0, 0.0, -0.0
    kinds = ("exact_rows", "exact_tiled")
    u, w = ops.induced_velocity(g, xw, zw, xp, zp, VC, mode="exact")
    p = _plan()
    assert p["kernel"] in kinds and p["variant"] == 1 and p["range_bad"] == 0, p
    uo, wo = oracle.induced_velocity(g, xw, zw, xp, zp, VC)
    assert biteq(u, uo) and biteq(w, wo)
    monkeypatch.setenv("LUDVM_EXACT_FLAGS", "1")          # the flagged instantiation on the same operands
    u, w = ops.induced_velocity(g, xw, zw, xp, zp, VC, mode="exact")
    assert _plan()["variant"] == 0 and biteq(u, uo) and biteq(w, wo)
    monkeypatch.delenv("LUDVM_EXACT_FLAGS")
    # out-of-window operands, one family at a time and together
    for fam in range(4):
        xw2, zw2, xp2, zp2 = xw.copy(), zw.copy(), xp.copy(), zp.copy()
        if fam in (0, 3):
            xw2[5::97], zp2[7::89] = 1e-300, -3e-301
        if fam in (1, 3):
            xw2[11::101], zw2[13::103], xp2[17::107] = 5e-324, -2.5e-310, 1e-320
        if fam in (2, 3):
            xw2[19::109], zp2[23::113] = 1e200, -4e199
        u, w = ops.induced_velocity(g, xw2, zw2, xp2, zp2, VC, mode="exact")
        p = _plan()
        assert p["variant"] == 1 and p["range_bad"] == 1, (fam, p)
        with np.errstate(all="ignore"):
            uo2, wo2 = oracle.induced_velocity(g, xw2, zw2, xp2, zp2, VC)
        assert biteq(u, uo2) and biteq(w, wo2), fam
    with np.errstate(all="ignore"):
        uo3, wo3 = oracle.induced_velocity(g, xw, zw, xp, zp, VC, viscous=False)
    u, w = ops.induced_velocity(g, xw, zw, xp, zp, VC, viscous=False, mode="exact")
    assert _plan()["variant"] == 0
    assert np.isnan(uo3).any() or npnt < 3
    assert np.array_equal(np.isnan(u), np.isnan(uo3)) and np.array_equal(np.isnan(w), np.isnan(wo3))
    ok = ~np.isnan(uo3) & ~np.isnan(wo3)
    assert biteq(u[ok], uo3[ok]) and biteq(w[ok], wo3[ok])

"""Sampler-level parity of the PRODUCTION (wavefront) kernels' re-implemented samplers against the oracle.

The benchmarked kernels do not call the lock-step device functions for the hot samplers: they use the mixed-precision,
block-draw versions of csrc/omc_physics_f32.cuh (drange_m, eloss_m, msdist_b incl. mscat / spinRejection, sscat_b, compton_b,
moller_b) and the Woodcock photon flight of omc_wavefront.cu.  omc_gpu_test_samplers() runs exactly those device functions on
explicit inputs; here each is compared with the oracle's function for the reference code it replaces
(src/ompmc.c:3979-4014 computeDrange, :4016-4108 computeEloss, :3787-3976 msdist, :3606-3785 mscat, :3097-3199 spinRejection /
sscat, :1670-1783 compton, :4359-4435 moller, :1951-2019 photon transport loop).  The oracle's functions are pinned bit for bit
to the unmodified reference's by tests/test_oracle_samplers.py.

Deterministic functions: relative tolerance written at the assert.  Samplers: the production code consumes its Philox stream
differently (whole blocks), so distributions are compared -- two-sample Kolmogorov-Smirnov on every observable of a step,
N = 10^5 per side and regime, plus means within 4.5 standard errors.
"""
import numpy as np
import pytest
from scipy import stats

from tests import sampler_cases as S

pytestmark = pytest.mark.gpu

N = 100_000
KS_P = 1e-4          # per-observable rejection level (seeds are fixed: the outcome is deterministic)


@pytest.fixture(scope="module")
def pair(gpu, oracle_lib):
    prob, ph = S.problem_tissue4()
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    return gpu, oracle_lib, prob, ph


def ks_distance_tol(a, b, rel=2e-4, abs_=2e-7):
    """Two-sample Kolmogorov-Smirnov distance that forgives a shift of the abscissa by rel * |x| + abs_.

    The step observables are NOT continuous: mscat() takes the tabulated u values without interpolation, u = 0 among
    them, and the no-scattering amplitude puts 1-30 % of the mass exactly on the axis, so the distributions carry atoms of a few
    per cent.  The production samplers place those atoms where fp32 arithmetic puts them (relative 1e-6 from the oracle's fp64
    positions); a plain KS distance then reports the MASS of the atom, whatever the sample size.  Here F_a is compared with F_b
    evaluated a tolerance to the right (and vice versa), which is the plain statistic for continuous parts and ignores shifts
    far below any physical scale (angles, path lengths: 2e-4 relative)."""
    a, b = np.sort(a), np.sort(b)
    x = np.concatenate([a, b])
    tol = rel * np.abs(x) + abs_
    d1 = np.searchsorted(a, x - tol, side="right") / len(a) - np.searchsorted(b, x + tol, side="right") / len(b)
    d2 = np.searchsorted(b, x - tol, side="right") / len(b) - np.searchsorted(a, x + tol, side="right") / len(a)
    return max(float(d1.max()), float(d2.max()), 0.0)


def ks_same(a, b, what):
    d = ks_distance_tol(a, b)
    n = len(a) * len(b) / (len(a) + len(b))
    pvalue = float(stats.kstwobign.sf(d * np.sqrt(n)))
    assert pvalue > KS_P, f"{what}: KS D = {d:.5f}, p = {pvalue:.2e}"
    se = np.sqrt(a.var() / len(a) + b.var() / len(b))
    if se > 0:
        assert abs(a.mean() - b.mean()) < 4.5 * se + 1e-7 * abs(b.mean()), f"{what}: means {a.mean():.8g} vs {b.mean():.8g} (se {se:.3g})"


def test_drange_eloss_vs_oracle(pair):
    gpu, orc, prob, ph = pair
    inp = S.drange_inputs(prob)
    g, o = gpu.test_samplers(S.DRANGE, inp)[:, 0], orc.test_samplers(S.DRANGE, inp)[:, 0]
    rel = np.abs(g - o) / o
    assert rel.max() < 1e-5, f"computeDrange: max rel error {rel.max():.3g}"          # fp32 series, fp64 energies
    inp = S.eloss_inputs(prob)
    g, o = gpu.test_samplers(S.ELOSS, inp), orc.test_samplers(S.ELOSS, inp)
    rel_r = np.abs(g[:, 0] - o[:, 0]) / o[:, 0]
    rel_e = np.abs(g[:, 1] - o[:, 1]) / o[:, 1]
    assert rel_r.max() < 1e-5, f"range: max rel error {rel_r.max():.3g}"
    assert rel_e.max() < 1e-5, f"computeEloss: max rel error {rel_e.max():.3g}"
    assert (g[:, 1] > 0).all() and (g[:, 1] <= inp[:, 3]).all()


@pytest.mark.parametrize("gi", range(len(S.MSDIST_GROUPS)), ids=[f"med{g[0]}_q{g[1]}_E{g[3]}_f{g[4]}" for g in S.MSDIST_GROUPS])
def test_msdist_distributions_vs_oracle(pair, gi):
    gpu, orc, prob, ph = pair
    grp = S.MSDIST_GROUPS[gi]
    for tilted in (False, True):
        inp = S.msdist_inputs(grp, N, tilted)
        d0 = inp[0, 5:8]
        g = gpu.test_samplers(S.MSDIST, inp, first_history=10_000_000 * (gi + 1))
        o = orc.test_samplers(S.MSDIST, inp, first=500_000_000 + 10_000_000 * gi)
        assert np.isfinite(g).all()
        # energy loss of the step is deterministic
        assert np.abs(g[:, 7] - o[:, 7]).max() <= 1e-5 * o[0, 7]
        tustep = np.median(o[:, 0])
        og, oo = S.msdist_observables(g, tustep, d0), S.msdist_observables(o, tustep, d0)
        assert np.abs(og["norm"] - 1.0).max() < 1e-6
        for k in ("omc", "z", "rperp", "ustep", "corr"):
            ks_same(og[k], oo[k], f"msdist {grp} tilted={tilted} {k}")
        # azimuthal symmetry around the initial direction
        sel = og["omc"] > 1e-6
        if sel.sum() > 1000:
            r = stats.kstest((og["phi"][sel] + np.pi) / (2 * np.pi), "uniform")
            assert r.pvalue > KS_P, f"msdist {grp}: azimuth not uniform, p = {r.pvalue:.2e}"


def test_sscat_vs_oracle(pair):
    gpu, orc, prob, ph = pair

    def xi_of(out):
        c, s = out[:, 0], out[:, 1]
        s2 = s * s
        return np.where(c > 0.5, s2 / (1.0 + np.sqrt(np.maximum(1.0 - s2, 0.0))), 1.0 - c)      # 1 - cos without cancellation

    for gi, grp in enumerate(S.sscat_groups(prob)):
        inp = np.tile(np.asarray(grp, dtype=np.float64), (N, 1))
        g = gpu.test_samplers(S.SSCAT, inp, first_history=3_000_000 * (gi + 1))
        o = orc.test_samplers(S.SSCAT, inp, first=700_000_000 + 3_000_000 * gi)
        ks_same(np.log(np.maximum(xi_of(g), 1e-30)), np.log(np.maximum(xi_of(o), 1e-30)), f"sscat {grp} log(1-cos)")
        assert np.abs(g[:, 0] ** 2 + g[:, 1] ** 2 - 1.0).max() < 1e-5
        assert np.abs(g[:, 2] ** 2 + g[:, 3] ** 2 - 1.0).max() < 1e-5
        phi = np.arctan2(g[:, 3], g[:, 2])
        r = stats.kstest((phi + np.pi) / (2 * np.pi), "uniform")
        assert r.pvalue > KS_P, f"sscat azimuth p = {r.pvalue:.2e}"


def test_compton_moller_vs_oracle(pair):
    gpu, orc, prob, ph = pair
    d = np.array([0.6, 0.0, 0.8])
    for k, e in enumerate(S.COMPTON_ENERGIES):
        inp = np.tile(np.array([e, *d]), (N, 1))
        g = gpu.test_samplers(S.COMPTON, inp, first_history=40_000_000 + 1_000_000 * k)
        o = orc.test_samplers(S.COMPTON, inp, first=800_000_000 + 1_000_000 * k)
        assert np.allclose(g[:, 0] + g[:, 4] - S.RM, e, rtol=1e-12)                        # energy conservation (fp64 sampler)
        ks_same(g[:, 0], o[:, 0], f"compton {e} MeV photon energy")
        # Compton kinematics: 1/br - 1 = k0 (1 - cos theta), photon and electron directions
        cg = g[:, 1:4] @ d
        assert np.abs((e / g[:, 0] - 1.0) - (e / S.RM) * (1.0 - cg)).max() < 1e-9 * max(1.0, e / S.RM)
        ks_same(g[:, 5:8] @ d, o[:, 5:8] @ d, f"compton {e} MeV electron polar cosine")
        assert np.abs(np.linalg.norm(g[:, 1:4], axis=1) - 1).max() < 1e-12 and np.abs(np.linalg.norm(g[:, 5:8], axis=1) - 1).max() < 1e-12
    for k, (imed, ek) in enumerate(S.MOLLER_ENERGIES):
        inp = np.tile(np.array([imed, ek + S.RM, *d]), (N, 1))
        g = gpu.test_samplers(S.MOLLER, inp, first_history=60_000_000 + 1_000_000 * k)
        o = orc.test_samplers(S.MOLLER, inp, first=900_000_000 + 1_000_000 * k)
        assert (g[:, 4] > 0).all() and np.allclose(g[:, 0] + g[:, 4], ek + 2 * S.RM, rtol=1e-12)
        ks_same(g[:, 4], o[:, 4], f"moller {ek} MeV secondary energy")
        ks_same(g[:, 1:4] @ d, o[:, 1:4] @ d, f"moller {ek} MeV primary cosine")
        ks_same(g[:, 5:8] @ d, o[:, 5:8] @ d, f"moller {ek} MeV secondary cosine")


@pytest.mark.parametrize("energy", [0.06, 1.0, 6.0])
def test_woodcock_flight_vs_reference_optical_depth(pair, energy):
    """The interaction depth of the Woodcock flight along a ray through air, tissue, lung and bone must be exponential in the
    optical depth that the reference's voxel-to-voxel march (photon() src/ompmc.c:1951-2019, howfar, gmfp x Rayleigh
    correction, density scaling) accumulates; the site's region index must be the voxel that holds the site (bit-exact)."""
    gpu, orc, prob, ph = pair
    n = 200_000
    zmid = 0.5 * (ph.zbounds[0] + ph.zbounds[-1])
    start = np.array([ph.xbounds[0] + 0.07, 0.13, zmid + 0.21])
    d = np.array([0.995, 0.05, 0.08]); d /= np.linalg.norm(d)
    inp = np.tile(np.array([energy, *start, *d]), (n, 1))
    g = gpu.test_samplers(S.WOODCOCK, inp, first_history=int(energy * 1000) * 1_000_000)
    hit = g[:, 0] > 0.5
    pos = g[hit, 1:4]
    # bit-exact index bookkeeping: the region of the site
    ix = np.searchsorted(ph.xbounds, pos[:, 0], side="right") - 1
    iy = np.searchsorted(ph.ybounds, pos[:, 1], side="right") - 1
    iz = np.searchsorted(ph.zbounds, pos[:, 2], side="right") - 1
    assert np.array_equal(g[hit, 4].astype(np.int64), 1 + ix + iy * ph.isize + iz * ph.isize * ph.jsize)
    # the sites lie on the ray
    s = (pos - start) @ d
    assert np.abs(pos - (start + np.outer(s, d))).max() < 1e-9
    q = np.tile(np.array([*start, *d]), (int(hit.sum()), 1))
    tau, _ = orc.test_photon_tau(energy, q, s)
    tau_tot, ir_end = orc.test_photon_tau(energy, q[:1], np.array([1000.0]))
    assert ir_end[0] == 0
    p_int = 1.0 - np.exp(-tau_tot[0])
    # escape fraction: binomial
    sd = np.sqrt(p_int * (1 - p_int) / n)
    assert abs(hit.mean() - p_int) < 4.5 * sd, f"interaction probability {hit.mean():.5f} vs {p_int:.5f} +- {sd:.5f}"
    u = (1.0 - np.exp(-tau)) / p_int
    r = stats.kstest(u, "uniform")
    assert r.pvalue > KS_P, f"Woodcock depth distribution at {energy} MeV: KS D = {r.statistic:.5f}, p = {r.pvalue:.2e}"
    assert g[:, 5].mean() >= 1.0


def test_step_size_phase_consistency(pair):
    """estep_size(): tperp is hownear(), the range is the one the energy-loss test pinned, and the class follows
    electron() src/ompmc.c:4960-4967 (condensed history iff the step fits the voxel and exceeds the skin depth)."""
    gpu, orc, prob, ph = pair
    rng = np.random.default_rng(17)
    n = 20000
    ix = rng.integers(2, ph.isize - 2, n); iy = rng.integers(2, ph.jsize - 2, n); iz = rng.integers(2, ph.ksize - 2, n)
    f = rng.random((n, 3)) * 0.98 + 0.01
    pos = np.column_stack([ph.xbounds[ix] + f[:, 0] * np.diff(ph.xbounds)[ix], ph.ybounds[iy] + f[:, 1] * np.diff(ph.ybounds)[iy],
                           ph.zbounds[iz] + f[:, 2] * np.diff(ph.zbounds)[iz]])
    ir = (1 + ix + iy * ph.isize + iz * ph.isize * ph.jsize).astype(np.int32)
    iq = rng.choice([-1, 1], n)
    e = np.exp(rng.uniform(np.log(0.25), np.log(15.0), n)) + S.RM
    out = gpu.test_samplers(S.ESTEP, np.column_stack([iq, e, pos, ir]), first_history=123456)
    cls, tustep, tperp, rng_, total_tstep, demfp, blccl, ssmfp = out.T
    _, _, _, tp = gpu.test_geometry(np.column_stack([pos, np.tile([0.0, 0.0, 1.0], (n, 1))]), ir, np.full(n, 1e10))
    assert np.array_equal(tperp, tp)
    live = cls > 0
    assert live.mean() > 0.9
    med = prob["region_med"][ir]
    rhof = prob["region_rhof"][ir]
    o = orc.test_samplers(S.ELOSS, np.column_stack([med, iq, rhof, e - S.RM, np.full(n, 0.5)]))
    m = live & (med >= 0)
    assert (np.abs(rng_[m] - o[m, 0]) <= 1e-5 * o[m, 0]).all()
    ch = (tustep <= tperp) & (tustep > 3 * ssmfp)
    assert np.array_equal(cls[m] == 1, ch[m])
    assert (tustep[m] <= rng_[m] * (1 + 1e-12)).all() and (demfp[m] >= 1e-5).all()

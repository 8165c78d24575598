"""Explicit inputs for the sampler-level tests (shared by the CPU test oracle-vs-reference and the GPU test
production-kernels-vs-oracle).  Sampler ids = OMC_SAMPLER_* of include/ompmc_b200.h."""
import numpy as np

DRANGE, ELOSS, MSDIST, SSCAT, COMPTON, MOLLER, WOODCOCK, ESTEP = range(8)
RM = 0.5109989461


def problem_tissue4(n=(40, 40, 40), voxel=(0.5, 0.5, 0.5)):
    """4-media TG119-like phantom (lung + bone inserts) with the committed 700icru tables."""
    from ompmc_b200 import problem as P
    media = P.load_blob(P.golden("media_700_tissue4.blob"))
    ph = P.tissue_phantom(n, voxel, "tg119")
    prob = P.build_problem(media, ph, ecut=0.700, pcut=0.010, collimator=(-5, 5, -5, 5), ssd=90.0,
                           cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]), nsplit=1)
    return prob, ph


def bin_edges(prob, imed, eke):
    """[lo, hi) kinetic-energy edges of the PWL bin that holds eke (pwlfInterval, src/ompmc.c:203)."""
    k = np.floor(prob["eke1"][imed] * np.log(eke) + prob["eke0"][imed])
    lo = np.exp((k - prob["eke0"][imed]) / prob["eke1"][imed])
    hi = np.exp((k + 1 - prob["eke0"][imed]) / prob["eke1"][imed])
    return lo, hi


def drange_inputs(prob, n=4000, seed=5):
    rng = np.random.default_rng(seed)
    nmed = int(prob["nmed"][0])
    imed = rng.integers(0, nmed, n)
    iq = rng.choice([-1, 1], n)
    ekei = np.exp(rng.uniform(np.log(0.25), np.log(20.0), n))
    lo, hi = bin_edges(prob, imed, ekei)
    # keep both energies 0.1 % inside the bin: the production code finds the bin with an fp32 log
    ekei = np.clip(ekei, lo * 1.002, hi * 0.998)
    ekef = lo * 1.001 + rng.random(n) * (ekei - lo * 1.001)
    return np.column_stack([imed, iq, ekei, ekef])


def eloss_inputs(prob, n=6000, seed=6):
    rng = np.random.default_rng(seed)
    nmed = int(prob["nmed"][0])
    imed = rng.integers(0, nmed, n)
    iq = rng.choice([-1, 1], n)
    rhof = rng.choice([0.3, 1.0, 1.037, 1.85], n)
    eke = np.exp(rng.uniform(np.log(0.25), np.log(20.0), n))
    lo, hi = bin_edges(prob, imed, eke)
    eke = np.clip(eke, lo * 1.002, hi * 0.998)
    # path lengths from 1e-4 of the range up to the whole range (second branch of computeEloss, and tuss <= 0)
    frac = np.where(rng.random(n) < 0.8, np.exp(rng.uniform(np.log(1e-4), 0.0, n)), rng.uniform(0.9, 1.0, n))
    frac[:50] = 1.0
    return np.column_stack([imed, iq, rhof, eke, frac])


# (imed, iq, rhof, eke [MeV], tustep / range): chosen to cover the lambda regimes of mscat() src/ompmc.c:3640-3785
# (plural scattering lambda <= 1, the no-scattering / single-scattering amplitudes below 13.8, the tabulated q-surfaces)
MSDIST_GROUPS = [
    (2, -1, 1.0, 0.30, 0.002), (2, -1, 1.0, 0.30, 0.05), (2, -1, 1.0, 1.0, 0.0005), (2, -1, 1.0, 1.0, 0.004),
    (2, -1, 1.0, 1.0, 0.05), (2, -1, 1.03, 1.0, 0.3), (2, -1, 1.0, 6.0, 0.001), (2, -1, 1.0, 6.0, 0.02),
    (2, 1, 1.0, 2.0, 0.01), (2, 1, 0.97, 0.5, 0.1), (3, -1, 1.85, 1.5, 0.003), (3, -1, 1.85, 1.5, 0.08),
    (1, -1, 0.26, 1.0, 0.01), (0, -1, 0.0012, 2.0, 0.001), (3, 1, 1.9, 10.0, 0.02), (2, -1, 1.0, 15.0, 0.1),
]


def msdist_inputs(group, n, tilted):
    imed, iq, rhof, eke, frac = group
    d = np.array([0.0, 0.0, 1.0]) if not tilted else np.array([0.48, -0.6, 0.64])
    return np.tile(np.array([imed, iq, rhof, eke, frac, d[0], d[1], d[2]]), (n, 1))


def msdist_observables(out, tustep_ref=None, d0=(0.0, 0.0, 1.0)):
    """Rotation-invariant observables of a condensed-history step relative to the initial direction d0."""
    d0 = np.asarray(d0, dtype=np.float64)
    ustep, disp, dirf = out[:, 0], out[:, 1:4], out[:, 4:7]
    z = disp @ d0
    perp = disp - np.outer(z, d0)
    rperp = np.linalg.norm(perp, axis=1)
    cost = dirf @ d0
    dperp = dirf - np.outer(cost, d0)
    # correlation between the lateral displacement and the lateral direction (PRESTA-II couples them)
    # (undefined for a step without deflection -- 1 to 30 % of the steps sit exactly on the axis: there both vectors are rounding
    # noise, 1e-17 in fp64 and 1e-8 in fp32, and their "angle" is garbage; such steps get corr = 0 on both sides)
    dn = np.linalg.norm(dperp, axis=1)
    scale_len = np.median(ustep) if tustep_ref is None else tustep_ref
    ok = (rperp > 1e-6 * scale_len) & (dn > 1e-6)
    corr = np.where(ok, np.einsum("ij,ij->i", perp, dperp) / np.where(ok, rperp * dn, 1.0), 0.0)
    # azimuth of the final direction around d0
    e1 = np.cross(d0, [1.0, 0.0, 0.0]) if abs(d0[0]) < 0.9 else np.cross(d0, [0.0, 1.0, 0.0])
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(d0, e1)
    phi = np.arctan2(dperp @ e2, dperp @ e1)
    scale = np.median(ustep) if tustep_ref is None else tustep_ref
    return {"omc": 1.0 - cost, "z": z / scale, "rperp": rperp / scale, "ustep": ustep / scale, "corr": corr, "phi": phi,
            "norm": np.linalg.norm(dirf, axis=1)}


def sscat_groups(prob):
    """(imed, qel, chia2, elke, beta2) as the boundary-crossing step computes them (electron() src/ompmc.c:5180-5207)."""
    groups = []
    for imed, qel, ekems in [(2, 0, 0.25), (2, 0, 1.0), (2, 0, 6.0), (3, 0, 1.0), (3, 1, 2.0), (1, 0, 0.5), (2, 1, 0.4)]:
        p2 = ekems * (ekems + 2.0 * RM)
        beta2 = p2 / (p2 + RM * RM)
        chia2 = prob["xcc"][imed] / (4.0 * prob["blcc"][imed] * p2)          # (x eta' ~ 1: the screening correction)
        groups.append((imed, qel, chia2, float(np.log(ekems)), beta2))
    return groups


COMPTON_ENERGIES = [0.03, 0.3, 0.9, 1.25, 6.0, 20.0]
MOLLER_ENERGIES = [(2, 1.0), (2, 3.0), (2, 12.0), (3, 2.0), (1, 5.0)]        # (imed, kinetic energy)

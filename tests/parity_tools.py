"""Statistics shared by the parity tests: batch-method moments, z-scores, gamma index at native resolution."""
import numpy as np


def batch_stats(accum, accum2, nbatch):
    """Per-voxel mean energy per batch and the variance of that mean (accumulateResults(), omc_dosxyz.c:719-799)."""
    mean = accum[1:] / nbatch
    var = np.maximum(accum2[1:] / nbatch - mean * mean, 0.0) / (nbatch - 1)
    return mean, var


def dose_grid(mean, ph):
    """Energy per batch -> dose-like quantity energy / mass on the (z, y, x) grid; voxels below 0.044 g/cm3 are zero as in
    accumulateResults() (omc_dosxyz.c:760-766)."""
    vol = np.multiply.outer(np.diff(ph.zbounds), np.multiply.outer(np.diff(ph.ybounds), np.diff(ph.xbounds)))
    dens = ph.med_densities.reshape(ph.ksize, ph.jsize, ph.isize)
    e = mean.reshape(ph.ksize, ph.jsize, ph.isize)
    return np.where(dens >= 0.044, e / (vol * np.maximum(dens, 1e-30)), 0.0)


def gamma_pass(ref, ev, spacing_mm, dd=0.01, dta_mm=1.0, cut=0.1, step_mm=0.25):
    """Global gamma index of `ev` against `ref` (both on the same grid, NATIVE resolution): dose criterion dd of the reference
    maximum, distance criterion dta, `ev` interpolated trilinearly at offsets up to 1.5 dta.  Returns (pass rate, number of
    voxels above the low-dose cut, largest gamma)."""
    from scipy.ndimage import map_coordinates
    dmax = ref.max()
    sel = ref > cut * dmax
    idx = np.argwhere(sel).astype(np.float64)
    best = np.full(len(idx), np.inf)
    r = np.arange(-1.5 * dta_mm, 1.5 * dta_mm + 1e-9, step_mm)
    rv = ref[sel]
    for dz in r:
        for dy in r:
            for dx in r:
                d2 = dx * dx + dy * dy + dz * dz
                if d2 > (1.5 * dta_mm) ** 2:
                    continue
                coords = (idx + np.array([dz / spacing_mm[2], dy / spacing_mm[1], dx / spacing_mm[0]])).T
                v = map_coordinates(ev, coords, order=1, mode="nearest")
                best = np.minimum(best, d2 / dta_mm ** 2 + ((v - rv) / (dd * dmax)) ** 2)
    gam = np.sqrt(best)
    return float((gam <= 1.0).mean()), int(sel.sum()), float(gam.max())

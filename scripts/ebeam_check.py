"""High-statistics check of the electron transport of the wavefront kernels against the lock-step kernel (which is in
lock-step with the reference): 6 MeV electron and 3 MeV positron pencil-ish beams on water, depth dose and lateral
profile.  Electron beams are the sensitive probe of the multiple-scattering samplers (msdist_b, sscat_b)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import json
import numpy as np
from test_gpu_wavefront import CASES, make_problem
from ompmc_b200.api import GpuTransport

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4000000
g = GpuTransport(0)
out = {}
for idx in (3, 4):
    name, cfg, nb, per = CASES[idx]
    prob, ph = make_problem(cfg)
    g.load_problem(prob)
    res = {}
    for kernel in (0, 1):
        g.set_option('kernel', kernel)
        g.reset_tallies()
        nb = 10
        for ib in range(nb):
            g.run_batch(ib * (n // nb), n // nb)
        a, a2, _ = g.get_tallies()
        m = a[1:] / nb
        v = np.maximum(a2[1:] / nb - m * m, 0) / (nb - 1)
        res[kernel] = (m.reshape(ph.ksize, ph.jsize, ph.isize), v.reshape(ph.ksize, ph.jsize, ph.isize))
    (m0, v0), (m1, v1) = res[0], res[1]
    dd0, dd1 = m0.sum(axis=(1, 2)), m1.sum(axis=(1, 2))                 # depth dose
    sd = np.sqrt(v0.sum(axis=(1, 2)) + v1.sum(axis=(1, 2)))
    lat0, lat1 = m0.sum(axis=(0, 1)), m1.sum(axis=(0, 1))               # lateral profile (x)
    sl = np.sqrt(v0.sum(axis=(0, 1)) + v1.sum(axis=(0, 1)))
    zd = (dd1 - dd0)[dd0 > 0.02 * dd0.max()] / sd[dd0 > 0.02 * dd0.max()]
    zl = (lat1 - lat0)[lat0 > 0.02 * lat0.max()] / sl[lat0 > 0.02 * lat0.max()]
    sel = (m0 > 0.1 * m0.max()) & (v0 + v1 > 0)
    z = (m1[sel] - m0[sel]) / np.sqrt(v0[sel] + v1[sel])
    out[name] = {"histories": n, "depth_dose_rel_diff_max": float(np.abs(dd1 / dd0 - 1)[dd0 > 0.05 * dd0.max()].max()),
                 "depth_dose_z": [round(float(x), 2) for x in zd], "lateral_z": [round(float(x), 2) for x in zl],
                 "voxel_z_mean": float(z.mean()), "voxel_z_std": float(z.std()), "voxels": int(sel.sum()),
                 "frac_within_2sigma": float((np.abs(z) < 2).mean()), "total_ratio": float(m1.sum() / m0.sum())}
    print(name, json.dumps(out[name]), flush=True)
json.dump(out, open('gpurun_out/ebeam_check.json', 'w'), indent=1)

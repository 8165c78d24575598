#!/usr/bin/env python
"""BASELINE config 5: PROSTATE 6 MV with the voxel grid up-sampled (3 mm -> 1.5 mm -> 1 mm): throughput and
statistical uncertainty per history count on one B200.  usage: python scripts/scaling_voxels.py [histories]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ompmc_b200 import problem as P  # noqa: E402
from ompmc_b200.api import GpuTransport  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
nb = 4
out = []
w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"]))
base = w["phantom"]()
g = GpuTransport(0)
for f in (1, 2, 3):
    t0 = time.time()
    ph = base if f == 1 else P.resample_phantom(base, (f, f, f))
    prob = P.build_problem(media, ph, ecut=w["ecut"], pcut=0.010, collimator=w["coll"], ssd=w["ssd"],
                           cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
    t1 = time.time()
    g.load_problem(prob)
    g.set_option("kernel", 1)
    t2 = time.time()
    g.run_batch(0, 1 << 20); g.synchronize()          # warm-up (allocations)
    g.reset_tallies()
    t3 = time.time()
    for ib in range(nb):
        g.run_batch((ib + 1) * n, n)
    g.synchronize()
    t4 = time.time()
    a, a2, _ = g.get_tallies()
    dose, unc = P.accumulate_results(ph, a, a2, n, nb)
    sel = dose > 0.5 * dose.max()
    c = g.counters()
    row = {"voxel_mm": 3.0 / f, "nvox": ph.nvox, "hist_per_s": nb * n / (t4 - t3), "histories": nb * n,
           "sigma_rel_above_half_dmax": float(unc[sel].mean()), "host_build_s": t1 - t0, "upload_s": t2 - t1,
           "photon_steps_per_hist": c["photon_steps"] / (nb * n), "electron_steps_per_hist": c["electron_steps"] / (nb * n)}
    row["histories_for_1pct"] = row["histories"] * (row["sigma_rel_above_half_dmax"] / 0.01) ** 2
    row["time_to_1pct_s"] = row["histories_for_1pct"] / row["hist_per_s"]
    print(json.dumps(row), flush=True)
    out.append(row)
    del prob, ph
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scaling_voxels.json"), "w"), indent=1)

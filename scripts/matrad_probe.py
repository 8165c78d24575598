"""BASELINE config 4 probe: beamlet dose-influence columns on the PROSTATE-like phantom through the matRad source,
throughput of the beamlet loop (histories/s, beamlets/s) on one GPU."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import bench
from ompmc_b200 import problem as P, matrad
from ompmc_b200.api import GpuTransport
nhist = sys.argv[1] if len(sys.argv) > 1 else "1000000"
w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"]))
ph = w["phantom"]()
bl = P.matrad_beamlets(ph, gantry_deg=(0.0, 72.0, 144.0, 216.0, 288.0), nbix=(3, 2), bixel_cm=0.5)
prob = P.build_problem_matrad(media, ph, bl, ecut=0.7, pcut=0.01, cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
nb = int(bl["mr_nbeamlets"][0])
g = GpuTransport(0)
g.load_problem(prob)
g.set_option("kernel", 1)
matrad.dose_influence_matrix(g, ph, 2, "100000", "10", 1e-3)          # warm-up
t0 = time.time()
jc, ir, val = matrad.dose_influence_matrix_device(g, ph, nb, nhist, "10", 1e-3, group=64)
dt = time.time() - t0
print(json.dumps({"driver": "multi-beamlet pass + device CSC", "beamlets": nb, "histories_per_beamlet": int(nhist), "seconds": dt,
                  "beamlets_per_s": nb / dt, "hist_per_s": nb * int(nhist) / dt, "nnz": int(jc[-1]), "nvox": ph.nvox}), flush=True)
for nbatch in ("10",):
    t0 = time.time()
    jc, ir, val = matrad.dose_influence_matrix(g, ph, nb, nhist, nbatch, 1e-3)
    dt = time.time() - t0
    print(json.dumps({"driver": "beamlet loop", "beamlets": nb, "histories_per_beamlet": int(nhist), "nbatch": int(nbatch), "seconds": dt, "beamlets_per_s": nb / dt,
                      "hist_per_s": nb * int(nhist) / dt, "nnz": int(jc[-1]), "nvox": ph.nvox}), flush=True)

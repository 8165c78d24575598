#!/usr/bin/env python
"""BASELINE config 4 at treatment-plan scale through the plain-C matRad driver (ompmc_b200/host/omc_matrad_b200.c): PROSTATE-like
phantom, 5 gantry angles x (nx x ny) bixels of 5 mm, nHistories per beamlet, relDoseThreshold 1e-3; one process per GPU
(-r rank -w world -d device), beamlets dealt round-robin; the per-rank CSC files are merged in beamlet order and compared with
the single-GPU matrix.  VERIFICATION / MEASUREMENT SCRIPT (not product code; like tests/ it may use the test infrastructure under oracle/).
usage: python scripts/run_config4.py [nx=8] [ny=8] [histories_per_beamlet=1000000]"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ompmc_b200 import build, matrad, problem as P  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nhist = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1000000


def read_csc(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"OMCCSC1\0"
        nrows, ncols, nnz = np.fromfile(f, dtype="<i8", count=3)
        jc = np.fromfile(f, dtype="<i8", count=ncols + 1)
        ir = np.fromfile(f, dtype="<i8", count=nnz)
        pr = np.fromfile(f, dtype="<f8", count=nnz)
    return int(nrows), int(ncols), jc, ir, pr


w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"]))
ph = w["phantom"]()
bl = P.matrad_beamlets(ph, gantry_deg=(0.0, 72.0, 144.0, 216.0, 288.0), nbix=(nx, ny), bixel_cm=0.5)
prob = P.build_problem_matrad(media, ph, bl, ecut=0.7, pcut=0.01, cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
nb = int(bl["mr_nbeamlets"][0])
os.makedirs("/tmp/c4", exist_ok=True)
P.save_blob("/tmp/c4/m.blob", prob)
build.build()
try:
    import torch
    ngpu = torch.cuda.device_count()
except Exception:
    ngpu = 1
rows = []
single = None
for world, group in [(n, 64) for n in (1, 2, 4, 8) if n <= ngpu] + [(1, 320), (1, 1)][:int(os.environ.get('C4_SWEEP', '0')) * 2]:
    t0 = time.time()
    procs = [subprocess.Popen([build.MATRAD_EXE, "-p", "/tmp/c4/m.blob", "-n", str(nhist), "-b", "10", "-t", "0.001", "-o", f"/tmp/c4/w{world}r{r}",
                               "-g", str(group), "-d", str(r), "-r", str(r), "-w", str(world)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate()[0] for p in procs]
    dt = time.time() - t0
    assert all(p.returncode == 0 for p in procs), outs[0][-1500:]
    parts = [read_csc(f"/tmp/c4/w{world}r{r}.csc") for r in range(world)]
    owner = {b0 + k: r for b0, n, r in matrad.beamlet_groups(nb, world, group if group else 1 << 16) for k in range(n)}
    cols = [parts[owner[b]] for b in range(nb)]
    nnz = sum(int(c[2][b + 1] - c[2][b]) for b, c in enumerate(cols))
    row = {"config": "config4_prostate_matrad", "gpus": world, "beamlets_per_pass": group, "beamlets": nb, "histories_per_beamlet": nhist, "wall_s": round(dt, 3),
           "beamlets_per_s": nb / dt, "hist_per_s": nb * nhist / dt, "nnz": nnz, "nvox": ph.nvox}
    row["per_rank"] = [[ln.strip() for ln in o.splitlines() if ln.startswith(("Beamlets computed", "Beamlets per pass", "Total execution time", "Execution time up"))] for o in outs]
    if single is None:
        single = cols
    else:                                     # same history ids per beamlet whatever the rank count: columns agree to fp32 summation order
        worst = 0.0
        for b in range(0, nb, max(1, nb // 40)):
            d0 = np.zeros(ph.nvox); d1 = np.zeros(ph.nvox)
            c0, c1 = single[b], cols[b]
            d0[c0[3][c0[2][b]:c0[2][b + 1]]] = c0[4][c0[2][b]:c0[2][b + 1]]
            d1[c1[3][c1[2][b]:c1[2][b + 1]]] = c1[4][c1[2][b]:c1[2][b + 1]]
            worst = max(worst, float(np.abs(d1 - d0).max() / d0.max()))
        row["max_column_diff_vs_1gpu_rel_to_column_max"] = worst
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"run_config4_{ngpu}gpu.json"), "w"), indent=1)

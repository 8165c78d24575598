#!/usr/bin/env python
"""BASELINE.json configs 1-3 end to end FROM THE REFERENCE'S INPUT-FILE FORMAT (.inp + .egsphant + PEGS4 / XCOM / msnew / spinms
data files) to a .3ddose file on one B200, twice per config:
  own     ompmc_b200/host/omc_dosxyz_b200 -i   (this repository's plain-C host: restated table builders, device-side writer)
  dropin  oracle/_ref/omc_dosxyz_dropin -i     (the reference's own user code + init code, batch loop on libompmc_b200.so)
Both hand the device the same problem and the same history ids, so the two dose files must agree to the summation order of
the fp32 dose atomics.  Phantoms are the synthetic stand-ins of bench.py (the reference checkout lacks its .egsphant files),
spinms.data the synthetic one of oracle/gen_fixtures.py.  VERIFICATION / MEASUREMENT SCRIPT (not product code; like tests/ it may use the test infrastructure under oracle/).
usage: python scripts/run_configs.py [scale=1.0]"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import gen_fixtures as G  # noqa: E402
from ompmc_b200 import build, problem as P  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
DROPIN = os.path.join(ROOT, "oracle", "_ref", "omc_dosxyz_dropin")
CONFIGS = [
    # BASELINE.json configs[0..2]
    dict(name="config1_water_250kV_521icru", phantom=lambda: P.water_phantom("H2O521ICRU", (61, 61, 60), (0.5, 0.5, 0.5)), pegs="521icru.pegs4dat",
         spectrum="250.spectrum", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.521, ncase=1000000),
    dict(name="config2_water_mohan6_700icru", phantom=bench.WORKLOADS["water6mv"]["phantom"], pegs="700icru.pegs4dat",
         spectrum="mohan6.spectrum", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.700, ncase=100000000),
    dict(name="config3_tg119_var6MV_700icru", phantom=bench.WORKLOADS["tg119_6mv"]["phantom"], pegs="700icru.pegs4dat",
         spectrum="var_6MV.spectrum", coll=(-5, 5, -5, 5), ssd=90.0, ecut=0.700, ncase=1000000000),
]
work = G.prepare_workdir()
build.build()
rows = []
for c in CONFIGS:
    ph = c["phantom"]()
    ppath = os.path.join(work, c["name"] + ".egsphant")
    P.write_egsphant(ppath, ph)
    stem = os.path.join(work, c["name"])
    ncase = max(10, int(c["ncase"] * scale))
    G.write_inp(stem, phantom=ppath, pegs=c["pegs"], spectrum=c["spectrum"], mono=0.0, charge=0, coll=c["coll"], ssd=c["ssd"], ecut=c["ecut"],
                pcut=0.01, nsplit=1, ncase=ncase, nbatch=10)
    row = {"config": c["name"], "nvox": ph.nvox, "ncase": ncase}
    files = {}
    for key, cmd, out in (("own", [build.HOST_EXE, "-i", stem, "-o", stem + "_own"], stem + "_own.3ddose"),
                          ("dropin", [DROPIN, "-i", stem, "-o", c["name"] + "_dropin"], stem + "_dropin.3ddose")):
        if not os.path.exists(cmd[0]):
            continue
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.time() - t0
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
        m = re.search(r"(?:Histories per second \(batch loop\): |histories/s)", r.stdout)
        rate = None
        for ln in r.stdout.splitlines():
            if ln.startswith("Histories per second (batch loop):"):
                rate = float(ln.split(":")[1])
            elif ln.startswith("Batch loop:"):
                rate = float(ln.split(",")[1].split()[0])
        row[key] = {"wall_s": round(wall, 3), "batch_loop_hist_per_s": rate}
        files[key] = out
    if "own" in files:
        dims, _, dose, unc = P.read_3ddose(files["own"])
        sel = dose > 0.5 * dose.max()
        row["sigma_rel_above_half_dmax"] = float(unc[sel].mean())
        if "dropin" in files:
            _, _, dose2, unc2 = P.read_3ddose(files["dropin"])
            row["dropin_vs_own_max_rel_diff_above_20pct_dmax"] = float(np.abs(dose2 / np.maximum(dose, 1e-300) - 1.0)[dose > 0.2 * dose.max()].max())
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "run_configs.json"), "w"), indent=1)

#!/usr/bin/env python
"""Output phase of BASELINE config 5 (up-sampled PROSTATE grid): outputResults() with the statistics and the .3ddose text on the
device (omc_gpu_write_3ddose) against the reference's per-value fprintf loop on the host (plain-C driver, OMC_HOST_RESULTS=3:
both writers on the same tallies); the two files must be byte-identical.  VERIFICATION / MEASUREMENT SCRIPT (not product code; like tests/ it may use the test infrastructure under oracle/).
usage: python scripts/writer_probe.py [factor=2] [histories=2e6]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ompmc_b200 import build, problem as P  # noqa: E402

f = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2000000
w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"]))
ph = w["phantom"]()
if f > 1:
    ph = P.resample_phantom(ph, (f, f, f))
prob = P.build_problem(media, ph, ecut=w["ecut"], pcut=0.010, collimator=w["coll"], ssd=w["ssd"],
                       cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
os.makedirs("/tmp/wp", exist_ok=True)
P.save_blob("/tmp/wp/p.blob", prob)
build.build()
row = {"nvox": ph.nvox, "voxel_mm": 3.0 / f, "histories": n}
r = subprocess.run([build.HOST_EXE, "-p", "/tmp/wp/p.blob", "-n", str(n), "-b", "4", "-o", "/tmp/wp/out"], capture_output=True, text=True,
                   env=dict(os.environ, OMC_HOST_RESULTS="3"))
assert r.returncode == 0, r.stdout[-800:] + r.stderr[-400:]
for ln in r.stdout.splitlines():
    if ln.startswith("Device writer:"):
        row["device_text_s"] = float(ln.split()[2])
    if ln.startswith("Host writer:"):
        row["host_fprintf_s"] = float(ln.split()[2])
row["file_bytes"] = os.path.getsize("/tmp/wp/out.3ddose")
row["identical"] = open("/tmp/wp/out.3ddose", "rb").read() == open("/tmp/wp/out_host.3ddose", "rb").read()
row["speedup"] = row["host_fprintf_s"] / row["device_text_s"]
print(json.dumps(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(row, open(os.path.join(ROOT, "gpurun_out", f"writer_probe_f{f}.json"), "w"))

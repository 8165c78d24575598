#!/usr/bin/env python
"""Join an ncu SASS source page (per-instruction counts) with nvdisasm line info -> per-source-line table.

usage: ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> [topN]
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, kre, cubin = sys.argv[1:4]
# "ncu-name-regex|||mangled-symbol-regex" when the two differ (templates)
kre, sre = kre.split("|||") if "|||" in kre else (kre, kre)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
end = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
# which launch of the matching kernels (env NCU_LAUNCH, default the first one)
import os
which = int(os.environ.get("NCU_LAUNCH", "0"))
print("kernel:", rows[end[which]][1])
H = rows[hdr_idx[which]]
stop = end[which + 1] if len(end) > which + 1 else len(rows)
body = [r for r in rows[hdr_idx[which] + 1:stop] if len(r) == len(H)]
col = {h: i for i, h in enumerate(H)}
base = int(body[0][col["Address"]], 16)
# nvdisasm with line info
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
fn_re = re.compile(r"^\s*\.text\.(\S+):")
infunc = False
cur = ("?", 0)
lines = {}
for ln in dis:
    m = fn_re.match(ln)
    if m:
        infunc = re.search(sre, m.group(1)) is not None
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0, 0, 0])
tot_inst = tot_thr = tot_samp = 0
for r in body:
    off = int(r[col["Address"]], 16) - base
    key = lines.get(off, ("?", 0))
    ie = int(r[col["Instructions Executed"]] or 0)
    te = int(r[col["Thread Instructions Executed"]] or 0)
    sm = int(r[col["# Samples"]] or 0)
    a = agg[key]
    a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot_inst += ie; tot_thr += te; tot_samp += sm
print(f"total warp-instr {tot_inst}  thread-instr {tot_thr}  avg active threads {tot_thr / max(tot_inst, 1):.2f}  samples {tot_samp}")
print(f"{'file:line':34s} {'warp-inst%':>10s} {'samples%':>9s} {'avg thr':>8s} {'#sass':>6s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{key[0] + ':' + str(key[1]):34s} {100 * a[0] / tot_inst:10.2f} {100 * a[2] / max(tot_samp, 1):9.2f} {a[1] / max(a[0], 1):8.2f} {a[3]:6d}")
# per file summary
pf = defaultdict(lambda: [0, 0, 0])
for key, a in agg.items():
    pf[key[0]][0] += a[0]; pf[key[0]][1] += a[1]; pf[key[0]][2] += a[2]
for f, a in pf.items():
    print("FILE", f, f"warp-inst {100 * a[0] / tot_inst:.1f}% samples {100 * a[2] / max(tot_samp, 1):.1f}% avg thr {a[1] / max(a[0], 1):.2f}")

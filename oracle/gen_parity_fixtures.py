#!/usr/bin/env python
"""Reference batch statistics for the statistical parity tests.  TEST INFRASTRUCTURE (oracle/).

Runs the UNMODIFIED reference (oracle/_ref/libompmc_ref_omp.so: src/ompmc.c + omc_random.c compiled where they lie,
its own RANMAR generator, its own `#pragma omp parallel for schedule(dynamic)` history loop, oracle/ref_harness.c:333)
on the scaled-down BASELINE configurations of oracle/parity_cases.py and writes tests/golden/parity_<case>.npz:

    mean, var   per-voxel mean energy deposit per batch and the variance OF THAT MEAN (batch method,
                accumulateResults() omc_dosxyz.c:719-799), float32, indexed like the dose grid without region 0
    totals      deposited energy of every batch (float64) -> total-energy test
    ensrc       score.ensrc
    info        json: histories, batches, threads, seconds, reference rate

The GPU box has no /root/reference, so these files are committed; tests/test_gpu_parity.py compares the production
CUDA kernels with them.  usage: python oracle/gen_parity_fixtures.py [case ...] [--scale F]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cpudrv  # noqa: E402
from oracle.parity_cases import CASES, build_case, fixture_path  # noqa: E402


def run_case(name: str, scale: float = 1.0) -> None:
    prob, ph, c = build_case(name)
    nb = c["nbatch"]
    per = int(c["nhist"] * scale) // nb
    ref = cpudrv.RefTransport(omp=True)
    ncores = len(os.sched_getaffinity(0))
    ref.set_num_threads(ncores)
    ref.load_problem(prob)
    ref.set_rng("ranmar")
    ref.set_nsplit(c["nsplit"])
    ref.reset_score()
    totals = np.zeros(nb)
    t0 = time.time()
    for ib in range(nb):
        ref.run_histories(ib * per, per)
        totals[ib] = ref.get_endep()[1:].sum()
        ref.accum_endep()
        if ib == 0:
            print(f"{name}: first batch of {per} histories {time.time() - t0:.1f} s -> ETA {(time.time() - t0) * nb:.0f} s", flush=True)
    dt = time.time() - t0
    a, a2, ensrc = ref.get_accum()
    mean = a[1:] / nb
    var = np.maximum(a2[1:] / nb - mean * mean, 0.0) / (nb - 1)
    info = {"case": name, "histories": per * nb, "nbatch": nb, "hist_per_batch": per, "threads": ref.num_threads(), "seconds": dt,
            "hist_per_s": per * nb / dt, "rng": "RANMAR (src/omc_random.c), one sequence per OpenMP thread", "nsplit": c["nsplit"],
            "spinms": "synthetic (McKinley-Feshbach), oracle/gen_fixtures.py", "grid": [ph.isize, ph.jsize, ph.ksize]}
    sel = mean > 0.5 * mean.max()
    info["rel_sigma_above_half_dmax"] = float((np.sqrt(var[sel]) / mean[sel]).mean())
    np.savez_compressed(fixture_path(name), mean=mean.astype(np.float32), var=var.astype(np.float32), totals=totals, ensrc=ensrc,
                        info=json.dumps(info))
    print(name, info, flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    for i, a in enumerate(sys.argv):
        if a == "--scale":
            scale = float(sys.argv[i + 1])
            args = [x for x in args if x != sys.argv[i + 1]]
    for name in (args or list(CASES)):
        run_case(name, scale)

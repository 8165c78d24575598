#!/usr/bin/env python
"""BASELINE config 5: PROSTATE 6 MV to a MEASURED 1 % sigma above half Dmax, voxel grid resampled to 2 mm / 1 mm, on N GPUs.
MEASUREMENT SCRIPT (not product code).

  python scripts/run_config5.py --voxel-mm 2                      (one GPU)
  torchrun --nproc-per-node N ... scripts/run_config5.py --voxel-mm 1

The 3 mm PROSTATE-like phantom of bench.py is resampled with ompmc_b200.problem.resample_phantom_to (any ratio; the C host
does the same with `omc_dosxyz_b200 -v`).  Batches of --hist-per-batch histories PER GPU are run through omc_gpu_run_batch()
(NCCL inside the library for N > 1) until the batch-method relative uncertainty, averaged over the voxels above half Dmax
(accumulateResults on the device), is <= --target; the statistics are checked after 4 batches and then where the last check
predicts the target (sigma ~ 1/sqrt(batches)), so the pipeline is interrupted a handful of times only.  Prints one JSON line:
wall time from the first batch to the batch that met the target (checks included), histories, rate."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ompmc_b200 import problem as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--voxel-mm", type=float, default=3.0)
ap.add_argument("--target", type=float, default=0.01)
ap.add_argument("--hist-per-batch", type=int, default=1 << 25, help="per GPU")
ap.add_argument("--max-batches", type=int, default=4000)
args = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ompmc_b200.api import GpuTransport  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"]))
ph = w["phantom"]()
if abs(args.voxel_mm - 3.0) > 1e-9:
    v = args.voxel_mm / 10.0
    ph = P.resample_phantom_to(ph, (v, v, v))
prob = P.build_problem(media, ph, ecut=w["ecut"], pcut=0.010, collimator=w["coll"], ssd=w["ssd"],
                       cdfinv=(media["cdfinv1_" + w["spectrum"]], media["cdfinv2_" + w["spectrum"]]), nsplit=1)
tr = GpuTransport(local)
tr.load_problem(prob)
tr.set_option("kernel", 1)
if world > 1:
    box = [tr.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tr.comm_init(rank, world, box[0])
H = args.hist_per_batch
tr.run_batch(10 ** 12, 2_000_000 * world); tr.synchronize(); tr.reset_tallies()      # warm-up (module load, queues, graph)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
nb, next_check, checks, sigma, t_check = 0, 4, [], 1.0, 0.0
while nb < args.max_batches:
    tr.run_batch(nb * H * world, H * world)
    nb += 1
    if nb < next_check:
        continue
    tc = time.perf_counter()
    dose, unc = tr.accumulate_results(ph.med_densities, H * world, nb)               # completes the batch in flight first
    sel = dose > 0.5 * dose.max()
    sigma = float(unc[sel].mean())
    t_check += time.perf_counter() - tc
    checks.append((nb, sigma))
    if sigma <= args.target:
        break
    next_check = max(nb + 1, int(np.ceil(nb * (sigma / args.target) ** 2 * 1.02)))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
dt = float(t.item())
if rank == 0:
    cnt = tr.counters()
    row = {"config": "config5_prostate6mv", "voxel_mm": args.voxel_mm, "grid": [ph.isize, ph.jsize, ph.ksize], "voxels": ph.nvox, "gpus": world,
           "target_sigma": args.target, "sigma_rel_above_half_dmax": sigma, "reached": sigma <= args.target, "batches": nb,
           "hist_per_batch_per_gpu": H, "histories": nb * H * world, "wall_s": dt, "of_which_statistics_checks_s": t_check,
           "hist_per_s": nb * H * world / dt, "checks": checks,
           "electron_steps_per_history": cnt["electron_steps"] / max(cnt["histories"], 1), "photon_steps_per_history": cnt["photon_steps"] / max(cnt["histories"], 1)}
    print(json.dumps(row), flush=True)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"config5_{args.voxel_mm:g}mm_n{world}.json"), "w") as f:
        json.dump(row, f, indent=1)
if world > 1:
    dist.destroy_process_group()

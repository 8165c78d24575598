#!/usr/bin/env python
"""bench.py -- headline benchmark of the ompMC shower() hot path on B200.

Metric (BASELINE.json): histories/s (and time to 1 % sigma at Dmax) of a 6 MV dose calculation on the
PROSTATE phantom, at 1/2/4/8 B200 beside the reference's OpenMP CPU path.

Workload (config.workload = "prostate6mv"): synthetic PROSTATE-like phantom (the real .egsphant is
missing from the reference checkout), 183x183x90 voxels of 3 mm, 4 media of 700icru.pegs4dat,
var_6MV.spectrum point source at SSD 90 cm, 10x10 cm2 field, ECUT 0.700 / PCUT 0.010, nsplit 1.
A "step" is one statistical batch = one pass {initHistory(); shower();} x H + accumEndep()
(omc_dosxyz.c:1237-1263) with H = --hist-per-step histories PER GPU (weak scaling); with N > 1 ranks
the batch grid is summed over ranks by NCCL inside the library before accumEndep()
(omc_gpu_comm_init, include/ompmc_b200.h).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU)

Prints ONE JSON line on rank 0.  `--impl reference` times the unmodified reference (oracle/_ref, all
host threads, its own RANMAR generator) on bounded samples of the same workload.

Other workloads of BASELINE.json (not the default line): `--workload matrad_prostate` = config 4 (a step = one pass of 64
beamlets x --hist-per-beamlet histories per GPU through omc_gpu_run_beamlets(); e2e adds the fetch of every pass's columns
and the in-library gather of all ranks' columns), `--voxel-mm 2|1` = config 5's resampled grids, `--nsplit 20` = the
splitting of the reference's shipped input file, `--workload tg119_6mv | water6mv` = configs 3 and 2.  OMC_BENCH_OPTIONS
("check_every=16,pool_size=...") passes tuning options to the library for A/B runs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ompmc_b200 import problem as P  # noqa: E402

WORKLOADS = {
    # name: (media blob, phantom builder, spectrum key, collimator, ssd, ecut)
    "prostate6mv": dict(media="media_700_tissue4.blob", phantom=lambda: P.tissue_phantom((183, 183, 90), (0.3, 0.3, 0.3), "prostate"),
                        spectrum="var_6MV", coll=(-5, 5, -5, 5), ssd=90.0, ecut=0.700, desc="PROSTATE-like 183x183x90 @3mm, 4 media 700icru, var_6MV, 10x10 cm2, SSD 90, nsplit 1"),
    "tg119_6mv": dict(media="media_700_tissue4.blob", phantom=lambda: P.tissue_phantom((167, 167, 129), (0.3, 0.3, 0.25), "tg119"),
                      spectrum="var_6MV", coll=(-5, 5, -5, 5), ssd=90.0, ecut=0.700, desc="TG119-like 167x167x129 @3x3x2.5mm, 4 media 700icru, var_6MV, 10x10 cm2, SSD 90, nsplit 1"),
    "water6mv": dict(media="media_700_water.blob", phantom=lambda: P.water_phantom("H2O700ICRU", (61, 61, 60), (0.5, 0.5, 0.5)),
                     spectrum="mohan6", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.700, desc="WATER 61x61x60 @5mm, H2O700ICRU, mohan6, 10x10 cm2, SSD 100, nsplit 1"),
}


def build_workload(name: str, nsplit: int = 1, voxel_mm: float = 0.0):
    w = WORKLOADS[name]
    media = P.load_blob(P.golden(w["media"]))
    ph = w["phantom"]()
    if voxel_mm > 0.0:                       # BASELINE config 5: the same phantom resampled to 2 mm / 1 mm voxels (any ratio)
        ph = P.resample_phantom_to(ph, (voxel_mm / 10.0,) * 3)
    prob = P.build_problem(media, ph, ecut=w["ecut"], pcut=0.010, collimator=w["coll"], ssd=w["ssd"],
                           cdfinv=(media["cdfinv1_" + w["spectrum"]], media["cdfinv2_" + w["spectrum"]]), nsplit=nsplit)
    return prob, ph, w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class quiet_stdout:
    """The reference prints from C (printf: 'RNG seeds ...', 'Warning!, negative ustep'); keep it off our stdout,
    where exactly one JSON line is expected."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved); os.close(self.null)
        return False


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when the first communicator comes up (NCCL_DEBUG=VERSION on the GPU boxes);
    stdout must carry exactly one JSON line, so the communicator set-up runs with fd 1 pointing at stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def alg_bytes_per_history(work: dict) -> float:
    """B_alg = 8 N_ausgab + 5 (N_howfar + N_hownear) + 8 N_pwlfEval + 14 N_mscat + 8 N_spinRejection."""
    return (8.0 * work["ausgab"] + 5.0 * (work["howfar"] + work["hownear"]) + 8.0 * work["pwlf"] + 14.0 * work["mscat"]
            + 8.0 * work["spin"])


# ---------------------------------------------------------------------------------------------
def cpu_reference_transport(prob):
    """The unmodified reference on all host threads when oracle/_ref was built, else the oracle port."""
    from oracle import cpudrv
    # The reference forces omp_get_num_procs() threads (omc_dosxyz.c:1184-1191); torchrun exports OMP_NUM_THREADS=1 to its
    # children, which must not throttle the baseline: ask for every core this process may run on.
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if cpudrv.have_ref(omp=True):
        tr, kind = cpudrv.RefTransport(omp=True), "reference"
        tr.set_num_threads(ncores)
        tr.load_problem(prob)
        tr.set_rng("ranmar")         # the reference's own generator
    else:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        tr, kind = cpudrv.OracleTransport(), "port"
        tr.set_num_threads(ncores)
        tr.load_problem(prob)
        tr.set_rng("ranmar")
    return tr, kind


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, ph, w = build_workload(args.workload, args.nsplit, args.voxel_mm)
    with quiet_stdout():
        tr, kind = cpu_reference_transport(prob)
        cores = tr.num_threads()
        # calibrate the per-step sample so the whole run ends within a few minutes
        t = tr.time_batches(0, 4000, 1)
        rate0 = 4000 / max(t, 1e-6)
        budget = 90.0 / max(args.steps + args.warmup, 1)
        nper = int(min(max(rate0 * min(budget, 6.0), 2000), 2_000_000))
        first = 10_000
        for i in range(args.warmup):
            tr.time_batches(first, nper, 1); first += nper
        t0 = time.perf_counter()
        for i in range(args.steps):
            tr.time_batches(first, nper, 1); first += nper
        dt = time.perf_counter() - t0
    value = args.steps * nper / dt
    sample = f"{args.steps} batches x {nper} histories of the workload, RANMAR, schedule(dynamic)"
    line = {"impl": "reference", "metric": "histories/s", "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"], "hist_per_step": nper, "nsplit": args.nsplit},
            "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: omc_matrad dose-influence matrix, beamlets sharded over the GPUs
MATRAD_DESC = ("omc_matrad: PROSTATE-like 183x183x90 @3mm, 4 media 700icru, var_6MV, 5 gantry angles x 25x25 bixels of 5 mm (3125 beamlets), "
               "relDoseThreshold 1e-3, 64 beamlets per pass")
MATRAD_PASS = 64


def build_matrad_workload():
    w = WORKLOADS["prostate6mv"]
    media = P.load_blob(P.golden(w["media"]))
    ph = w["phantom"]()
    bl = P.matrad_beamlets(ph, gantry_deg=(0.0, 72.0, 144.0, 216.0, 288.0), nbix=(25, 25), bixel_cm=0.5)
    prob = P.build_problem_matrad(media, ph, bl, ecut=0.7, pcut=0.01, cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
    return prob, ph, int(bl["mr_nbeamlets"][0])


def run_matrad(args) -> None:
    """One step = one pass of omc_gpu_run_beamlets() per GPU: 64 beamlets x --hist-per-beamlet histories through the wavefront
    kernels, accumulateResults + threshold + CSC columns on the device (omc_matrad.c:1389-1477).  Step i gives rank r the
    beamlets [64 (i world + r), +64) (weak scaling: 64 beamlets per GPU per step); beamlet b always owns history ids
    [b nhist, (b+1) nhist).  `value` times the passes with the columns left on the device; `e2e` adds what a user of the plugin
    pays: fetching every pass's columns to the host and the in-library gather of all ranks' columns into the complete matrix."""
    import torch
    import torch.distributed as dist
    from ompmc_b200 import build as builder
    from ompmc_b200.api import GpuTransport

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
    if rank == 0:
        builder.build()
    if world > 1:
        dist.barrier()
    prob, ph, nbeam = build_matrad_workload()
    nh, nbatch, rel = args.hist_per_beamlet, 10, 1.0e-3
    need = MATRAD_PASS * args.steps * world
    if need > nbeam:
        raise SystemExit(f"bench.py: {args.steps} steps x {world} GPUs x {MATRAD_PASS} beamlets exceed the plan's {nbeam} beamlets")
    tr = GpuTransport(local)
    tr.load_problem(prob)
    tr.set_option("kernel", 1)
    if world > 1:
        with stdout_to_stderr():
            box = [tr.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            tr.comm_init(rank, world, box[0])
            tr.comm_sum([1.0])
    stream = torch.cuda.ExternalStream(tr.stream_ptr(), device=f"cuda:{local}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(i: int, fetch: bool):
        b0 = MATRAD_PASS * (i * world + rank)
        with torch.cuda.stream(stream):
            flush.zero_()
        return b0, tr.run_beamlets(b0 * nh, nh, nbatch, b0, MATRAD_PASS, rel, ph.med_densities, fetch=fetch)

    for i in range(args.warmup):
        one_pass(i % max(args.steps, 1), False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = tr.counters()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        one_pass(i, False)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = tr.counters()["kernel_launches"] - l0
    # end to end: every pass's columns to the host, then the complete matrix on every rank
    barrier()
    t0 = time.perf_counter()
    mine = {}
    d2h = 0
    for i in range(args.steps):
        b0, (jc, ir, val) = one_pass(i, True)
        d2h += ir.nbytes + val.nbytes + jc.nbytes
        for k in range(MATRAD_PASS):
            mine[b0 + k] = (ir[jc[k]:jc[k + 1]], val[jc[k]:jc[k + 1]])
    nb_total = MATRAD_PASS * args.steps * world
    jc_all, ir_all, val_all = tr.comm_gather_columns(nb_total, mine)
    barrier()
    e2e_s = time.perf_counter() - t0
    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = (float(x) for x in times.tolist())
    if rank == 0:
        total_hist = args.steps * MATRAD_PASS * nh * world
        peak, peak_kind = measured_peaks()
        line = {"metric": "histories/s", "value": total_hist / (ms * 1e-3), "unit": "histories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "mixed f64/f32", "data": "synthetic",
                "config": {"workload": "matrad_prostate", "desc": MATRAD_DESC, "beamlets_per_step_per_gpu": MATRAD_PASS, "hist_per_beamlet": nh,
                           "beamlets_total": nb_total, "kernel": "wavefront", "l2": "256 MiB buffer written between steps (L2 flush)"},
                "clocks": clocks, "gpu_launches": int(launches), "beamlets_per_s": args.steps * MATRAD_PASS * world / (ms * 1e-3),
                "e2e": {"value": total_hist / (e2e_ms * 1e-3), "unit": "histories/s", "beamlets_per_s": nb_total / (e2e_ms * 1e-3),
                        "h2d_bytes_per_step": int(ph.nvox * 8), "d2h_bytes_per_step": int(d2h / max(args.steps, 1)),
                        "steps": args.steps, "gather": "omc_gpu_comm_gather_columns (NCCL inside the library)" if world > 1 else "single rank"},
                "matrix": {"rows": int(ph.nvox), "cols": int(nb_total), "nnz": int(jc_all[-1]),
                           "checksum": float(val_all.sum()), "row_checksum": int(ir_all.sum())}}
        work = json.load(open(os.path.join(ROOT, "profiles", "work_counts.json"))).get("prostate6mv")
        if work:
            balg = alg_bytes_per_history(work)
            achieved = balg * MATRAD_PASS * nh / (ms / args.steps * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                                "peak_kind": peak_kind, "alg_bytes_per_history": balg,
                                "note": "per-history work counts of the prostate6mv dosxyz workload (same phantom, spectrum and cut-offs)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_matrad_reference_arm(args) -> None:
    """The unmodified reference on the host cores for the same beamlets: initHistory(ibeamlet) + shower() of omc_matrad.c through
    the reference harness, a bounded number of histories per beamlet."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    prob, ph, nbeam = build_matrad_workload()
    from oracle import cpudrv
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with quiet_stdout():
        if not cpudrv.have_ref(omp=True, matrad=True):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libompmc_ref_matrad_omp.so not built (needs /root/reference)"}))
            return
        tr, kind = cpudrv.RefTransport(omp=True, matrad=True), "reference"
        tr.set_num_threads(ncores)
        tr.load_problem(prob)
        tr.set_rng("ranmar")
        cores = tr.num_threads()
        nper = 20000
        t0 = time.perf_counter()
        done = 0
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                t0 = time.perf_counter(); done = 0
            for b in range(8):                                   # eight beamlets of the step's pass, nper histories each
                bb = (MATRAD_PASS * i + 8 * b) % nbeam
                tr.set_beamlet(bb)
                tr.run_histories(bb * args.hist_per_beamlet, nper)
                done += nper
        dt = time.perf_counter() - t0
    value = done / dt
    line = {"impl": "reference", "metric": "histories/s", "value": value, "unit": "histories/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "matrad_prostate", "desc": MATRAD_DESC, "hist_per_step": 8 * nper},
            "cpu_baseline": {"value": value, "unit": "histories/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} steps x 8 beamlets x {nper} histories, RANMAR, schedule(dynamic)"},
            "e2e": {"value": value, "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="prostate6mv", choices=list(WORKLOADS) + ["matrad_prostate"])
    ap.add_argument("--hist-per-beamlet", type=int, default=1_000_000, help="matrad_prostate: nHistories of every beamlet")
    ap.add_argument("--voxel-mm", type=float, default=0.0, help="BASELINE config 5: resample the phantom to this voxel size (2 or 1)")
    # one step = one statistical batch; 1 % sigma above half Dmax needs ~6e8 histories on this workload = 10 batches of ~6e7
    ap.add_argument("--hist-per-step", type=int, default=1 << 26, help="histories per step PER GPU")
    ap.add_argument("--kernel", type=int, default=-1, help="-1 production default, 0 lock-step, 1 wavefront")
    ap.add_argument("--nsplit", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=-1)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.workload == "matrad_prostate":
            run_matrad_reference_arm(args)
        else:
            run_reference_arm(args)
        return
    if args.workload == "matrad_prostate":
        run_matrad(args)
        return

    import torch
    import torch.distributed as dist
    from ompmc_b200 import build as builder
    from ompmc_b200.api import GpuTransport, DEFAULT_KERNEL

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
    if rank == 0:
        builder.build()
    if world > 1:
        dist.barrier()

    prob, ph, w = build_workload(args.workload, args.nsplit, args.voxel_mm)
    tr = GpuTransport(local)
    tr.load_problem(prob)
    kernel = DEFAULT_KERNEL if args.kernel < 0 else args.kernel
    tr.set_option("kernel", kernel)
    for kv in filter(None, os.environ.get("OMC_BENCH_OPTIONS", "").split(",")):      # tuning experiments: "check_every=16,pool_size=..."
        k, v = kv.split("=")
        tr.set_option(k.strip(), int(v))
    tr.reset_tallies()
    H = args.hist_per_step
    stream = torch.cuda.ExternalStream(tr.stream_ptr(), device=f"cuda:{local}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")      # > 126 MB L2
    if world > 1:
        # NCCL inside the library (include/ompmc_b200.h, multi-GPU): rank 0 makes the unique id, torch.distributed only carries
        # its 128 bytes to the other ranks.  From here on run_batch() shards every batch over the ranks and sums the completed
        # batch grids on a side stream before accumEndep(); torch takes no part in the data path.
        with stdout_to_stderr():
            box = [tr.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            tr.comm_init(rank, world, box[0])
            tr.comm_sum([1.0])                                # first collective of the library's communicator

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i: int):
        # batch i = history ids [i*H*world, (i+1)*H*world); the library gives rank r its contiguous slice.  Batches are
        # pipelined: batch i is started while the tail of batch i-1 is still in flight; i-1 is then summed over the ranks
        # and accumulated on the side stream.
        with torch.cuda.stream(stream):
            flush.zero_()                                     # L2 flush between timed iterations
        tr.run_batch(i * H * world, H * world)

    def finish():
        tr.synchronize()

    for i in range(args.warmup):
        step(i)
    finish()
    barrier()
    tr.reset_tallies()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    finish()                                                  # the tail of the last batch belongs to the timed region
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    cnt = tr.counters()
    # kernel-only time of the dominant (transport) kernels: same steps, events around run_histories only
    barrier()
    kms = 0.0
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nk = max(2, min(args.steps, 4))
    ka.record(stream)
    for i in range(nk):                                        # same pipelined batches, no L2 flush: the launch group of the roofline
        tr.run_batch((args.warmup + args.steps + i) * H * world, H * world)
    finish()
    kb.record(stream)
    torch.cuda.synchronize()
    kms = ka.elapsed_time(kb)
    kernel_ms = kms / nk
    launches_per_step = cnt["kernel_launches"] / max(args.steps, 1)

    # sigma at Dmax: mean relative uncertainty over voxels with D > 0.5 Dmax, all batches so far
    a, a2, ensrc = tr.get_tallies()
    nb = args.steps + nk
    dose, unc = P.accumulate_results(ph, a, a2, H * world, nb)
    sel = dose > 0.5 * dose.max()
    sigma = float(unc[sel].mean())

    # end-to-end through the C-ABI with host buffers: upload problem, run a batch, read the tallies back
    e2e_steps = args.e2e_steps if args.e2e_steps >= 0 else max(2, args.steps // 2)
    h2d = sum(v.nbytes for k, v in prob.items() if not k.startswith(("med_", "pegs_rho", "pegs_ae", "cdfinv")))
    d2h = 2 * ph.nreg * 8 + 8
    # the host arrays of the problem in page-locked memory (same contents; numpy views of pinned torch tensors)
    e2e_prob, pinned = prob, False
    try:
        keep = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in prob.items()}
        e2e_prob = {k: t.numpy() for k, t in keep.items()}
        pinned = all(t.is_pinned() for t in keep.values())
    except Exception:
        e2e_prob, pinned = prob, False
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        tr.load_problem(e2e_prob)                              # H2D of tables + geometry + source (host arrays)
        tr.set_option("kernel", kernel)
        tr.run_batch((1000 + i) * H * world, H * world)
        tr.get_tallies()                                       # D2H of accum_endep / accum_endep2 (completes the batch first)
    barrier()
    e2e_s = time.perf_counter() - t0

    times = torch.tensor([ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    per_rank = None
    if world > 1:
        parts = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(parts, times)
        per_rank = [round(float(p[0]) / args.steps, 2) for p in parts]      # each rank's own ms per step (skew diagnostic)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = (float(x) for x in times.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_hist = args.steps * H * world
    value = total_hist / (ms * 1e-3)
    e2e_value = e2e_steps * H * world / (e2e_ms * 1e-3)
    peak, peak_kind = measured_peaks()
    line = {"metric": "histories/s", "value": value, "unit": "histories/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "mixed f64/f32" if kernel == 1 else "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": w["desc"] + (f", resampled to {args.voxel_mm:g} mm voxels ({ph.isize}x{ph.jsize}x{ph.ksize})" if args.voxel_mm > 0 else ""),
                       "hist_per_step_per_gpu": H, "nsplit": args.nsplit,
                       "kernel": {0: "lockstep", 1: "wavefront"}[kernel], "rng": "philox4x32-10 per history",
                       "l2": "256 MiB buffer written between steps (L2 flush)", "spinms": "synthetic (McKinley-Feshbach)"},
            "clocks": clocks, "gpu_launches": int(cnt["kernel_launches"]),
            "e2e": {"value": e2e_value, "unit": "histories/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "host_buffers": "pinned" if pinned else "pageable"},
            "ms_per_step_per_rank": per_rank,
            "sigma_rel_above_half_dmax": sigma, "histories_scored": int(nb * H * world),
            "time_to_1pct_sigma_s": (nb * H * world / value) * (sigma / 0.01) ** 2}

    # CPU legs (rank 0, N = 1 only): per-history work counts from the oracle -> algorithmic bytes; reference timing
    work = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpudrv
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        orc = cpudrv.OracleTransport()
        orc.load_problem(prob)
        orc.set_rng("philox")
        orc.reset_score()
        orc.run_histories(0, 40000)
        work = orc.work_per_history()
        with quiet_stdout():
            ref, kind = cpu_reference_transport(prob)
            t = ref.time_batches(0, 4000, 1)
            n = int(min(max(4000 / max(t, 1e-6) * 15.0, 4000), 4_000_000))
            t = ref.time_batches(100000, n, 1)
        line["cpu_baseline"] = {"value": n / t, "unit": "histories/s", "cores": ref.num_threads(), "kind": kind,
                                "sample": f"1 batch of {n} histories of the same workload ({t:.1f} s), RANMAR, all host threads"}
    if work is None:
        work = json.load(open(os.path.join(ROOT, "profiles", "work_counts.json")))[args.workload] if os.path.exists(
            os.path.join(ROOT, "profiles", "work_counts.json")) else None
    if work is not None:
        balg = alg_bytes_per_history(work)
        achieved = balg * H / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):       # measured DRAM bytes per history (ncu --set full) x histories of one launch group
            rec = json.load(open(tp)).get(f"{args.workload}:{line['config']['kernel']}")
            if rec:
                traffic = rec["dram_bytes_per_history"] * H
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "peak_kind": peak_kind, "alg_bytes_per_history": balg, "work_per_history": work,
                            "kernel_ms_per_launch_group": kernel_ms, "histories_per_launch_group": H,
                            "launches_per_step": launches_per_step}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

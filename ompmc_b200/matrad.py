"""Dose-influence matrix driver == the beamlet loop of ucodes/omc_matrad/omc_matrad.c:1389-1493 without MATLAB.

For every beamlet: nbatch batches of nperbatch histories, accumulateResults(iout=1, nhist, nbatch) (note:
normalised by the TOTAL history count, unlike omc_dosxyz -- SURVEY Q11), threshold at relDoseThreshold * max,
append one CSC column.  Between beamlets only accum_endep is zeroed: the sigma grid is never
exported, so this affects nothing that leaves the loop.

Beamlets are independent, so with several ranks they are dealt round-robin (per-beamlet loop: beamlet b -> rank
b % world; multi-beamlet pass: whole groups of consecutive beamlets, beamlet_groups()), each rank builds its columns, and the columns are gathered in beamlet order (no collective inside the transport).
"""
from __future__ import annotations

import numpy as np

from . import problem as P


def beamlet_column(ph: P.Phantom, accum: np.ndarray, accum2: np.ndarray, nhist: int, nbatch: int, rel_threshold: float):
    """accumulateResults + threshold + sparse column of omc_matrad.c:1416-1477: returns (rows, values)."""
    dose, _unc = P.accumulate_results(ph, accum, accum2, nhist, nbatch, iout=1)
    dmax = dose.max() if dose.size else 0.0
    thresh = dmax * rel_threshold
    rows = np.nonzero(dose > thresh)[0]
    return rows.astype(np.int64), dose[rows]


def dose_influence_matrix(tr, ph: P.Phantom, nbeamlets: int, ncase, nbatch, rel_threshold: float, rank: int = 0, world: int = 1,
                          gather=None, first_history: int = 0):
    """Run the beamlet loop on transport `tr` (GpuTransport or a CPU checker with the same methods).

    Returns (jc, ir, values) = CSC arrays of the nvox x nbeamlets matrix (like mxCreateSparse's Jc/Ir/Pr).  Every
    beamlet b uses history ids [first_history + b*nhist, first_history + (b+1)*nhist), so the result does not
    depend on how beamlets are distributed over ranks."""
    nhist, nb, nper = P.batch_plan(ncase, nbatch)
    mine = {}
    tr.reset_tallies()
    for b in range(rank, nbeamlets, world):
        for ib in range(nb):
            tr.run_batch(first_history + b * nhist + ib * nper, nper, b)
        accum, accum2, _ = tr.get_tallies()
        mine[b] = beamlet_column(ph, accum, accum2, nhist, nb, rel_threshold)
        tr.reset_tallies(1)                       # memset(score.accum_endep) only, omc_matrad.c:1482
    cols = mine if gather is None or world == 1 else gather(mine)
    jc = np.zeros(nbeamlets + 1, dtype=np.int64)
    irs, vals = [], []
    for b in range(nbeamlets):
        r, v = cols[b]
        irs.append(r); vals.append(v)
        jc[b + 1] = jc[b] + len(r)
    return jc, (np.concatenate(irs) if irs else np.zeros(0, np.int64)), (np.concatenate(vals) if vals else np.zeros(0))


def gather_columns_torch(mine: dict, group=None) -> dict:
    """all_gather_object of the per-rank column dicts (host-side gather)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    out = {}
    for p in parts:
        out.update(p)
    return out


def beamlet_groups(nbeamlets: int, world: int = 1, group: int = 64):
    """Sharding plan of the multi-beamlet pass: contiguous groups of at most `group` beamlets (one pass of the wavefront
    kernels each -- the pass pays the tail of its longest particle lineages once per GROUP, so a rank must own whole groups,
    not single beamlets), as many groups as a multiple of `world`, dealt round-robin.  Returns [(first beamlet, count, rank)].
    ompmc_b200/host/omc_matrad_b200.c computes the same plan."""
    if nbeamlets <= 0:
        return []
    world, group = max(1, int(world)), max(1, int(group))
    ngroups = world * -(-nbeamlets // (world * group))
    size = -(-nbeamlets // ngroups)
    return [(b0, min(size, nbeamlets - b0), (b0 // size) % world) for b0 in range(0, nbeamlets, size)]


def dose_influence_matrix_device(tr, ph: P.Phantom, nbeamlets: int, ncase, nbatch, rel_threshold: float, rank: int = 0, world: int = 1,
                                 gather=None, first_history: int = 0, group: int | None = None):
    """Same matrix as dose_influence_matrix() built the B200 way (north_star: "each GPU builds its slice of the sparse
    dose-influence matrix"): the beamlets of a rank are run `group` at a time in ONE pass of the wavefront kernels
    (omc_gpu_run_beamlets: history id -> beamlet -> its own dose grid; no per-batch tail per beamlet), and
    accumulateResults + threshold + column assembly run on the device.  Beamlet b keeps the history ids
    [first_history + b*nhist, +nhist), so up to scheduling-independent statistics this is the per-beamlet loop's result."""
    nhist, nb, nper = P.batch_plan(ncase, nbatch)
    mine = {}
    if group is None:               # 64 per pass unless the grids would not fit (api.beamlet_capacity)
        group = tr.beamlet_capacity()
    for b0, n, owner in beamlet_groups(nbeamlets, world, group):
        if owner != rank:
            continue
        jc, ir, val = tr.run_beamlets(first_history + b0 * nhist, nhist, nb, b0, n, rel_threshold, ph.med_densities)
        for k in range(n):
            mine[b0 + k] = (ir[jc[k]:jc[k + 1]].copy(), val[jc[k]:jc[k + 1]].copy())
    cols = mine if gather is None or world == 1 else gather(mine)
    jc = np.zeros(nbeamlets + 1, dtype=np.int64)
    irs, vals = [], []
    for b in range(nbeamlets):
        r, v = cols[b]
        irs.append(r); vals.append(v)
        jc[b + 1] = jc[b] + len(r)
    return jc, (np.concatenate(irs) if irs else np.zeros(0, np.int64)), (np.concatenate(vals) if vals else np.zeros(0))

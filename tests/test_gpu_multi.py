"""Multi-GPU inside the C library (include/ompmc_b200.h, "multi-GPU"): NCCL communicator per handle, history ids of a batch
sharded over the ranks by omc_gpu_run_batch(), completed batch grids summed on a side stream BEFORE accumEndep() squares
them (the statistics must be those of a single-GPU run of the same batches), one handle over several devices for
single-process C user codes (omc_gpu_multi_*), the reference's own omc_dosxyz user code on it (-g N).

The driver's GPU box has one device: the world-size-2 cases skip there and are run with `gpurun --gpus 2`
(scripts/gpu_multi.sh; result committed under profiles/)."""
import os
import subprocess

import numpy as np
import pytest

from ompmc_b200 import build, problem as P
from tests.test_gpu_wavefront import CASES, make_problem

pytestmark = pytest.mark.gpu


def ndev():
    import torch
    return torch.cuda.device_count()


def run_batches(tr, nb, per):
    tr.set_option("kernel", 1)
    # no drain kernel: it finishes the last few particles of a run on CONTINUED random streams, so which particles it takes
    # depends on how the batches are cut; without it a history is the same whichever device runs it
    tr.set_option("drain_threshold", 0)
    try:
        tr.reset_tallies()
        for ib in range(nb):
            tr.run_batch(ib * per, per)
        a, a2, e = tr.get_tallies()
        return a[1:], a2[1:], e, tr.counters()
    finally:
        tr.set_option("drain_threshold", 8192)


def same_statistics(ref, got, nb):
    (a0, b0, e0, c0), (a1, b1, e1, c1) = ref, got
    assert c1["histories"] == c0["histories"] and c1["errors"] == 0
    assert abs(e1 - e0) <= 1e-9 * e0                                    # the same source particles, summed in another order
    assert c1["deposits"] == c0["deposits"] and c1["electron_steps"] == c0["electron_steps"]     # (no drain: the same histories)
    # history id -> RNG stream: every history is the same whichever device ran it; only the fp32 atomics order differs
    np.testing.assert_allclose(a1, a0, rtol=3e-4, atol=1e-4 * a0.max())
    np.testing.assert_allclose(b1, b0, rtol=6e-4, atol=1e-4 * b0.max())


def test_one_device_behind_the_multi_handle(gpu):
    from ompmc_b200.api import MultiGpuTransport
    prob, ph = make_problem(CASES[1][1])
    gpu.load_problem(prob)
    ref = run_batches(gpu, 4, 50000)
    m = MultiGpuTransport(ndev=1)
    try:
        assert m.ndev == 1
        m.load_problem(prob)
        got = run_batches(m, 4, 50000)
    finally:
        m.close()
    same_statistics(ref, got, 4)


def test_communicator_of_one_is_a_no_op(gpu):
    prob, ph = make_problem(CASES[0][1])
    gpu.load_problem(prob)
    ref = run_batches(gpu, 3, 20000)
    gpu.comm_init(0, 1, gpu.comm_unique_id())
    got = run_batches(gpu, 3, 20000)
    same_statistics(ref, got, 3)
    assert np.array_equal(gpu.comm_sum([1.5, 2.0]), [1.5, 2.0])


@pytest.mark.parametrize("n", [2, 4, 8])
def test_batches_sharded_over_devices_equal_one_device(gpu, n):
    if ndev() < n:
        pytest.skip(f"needs {n} GPUs")
    from ompmc_b200.api import MultiGpuTransport
    prob, ph = make_problem(CASES[1][1])
    gpu.load_problem(prob)
    nb, per = 6, 200001                                                 # (not divisible by the device count: ragged slices)
    ref = run_batches(gpu, nb, per)
    m = MultiGpuTransport(ndev=n)
    try:
        m.load_problem(prob)
        got = run_batches(m, nb, per)
        # the statistics live on every device: device 1's copy equals device 0's bit for bit (same all-reduced grids)
        dose, unc = m.accumulate_results(ph.med_densities, per, nb)
    finally:
        m.close()
    same_statistics(ref, got, nb)
    d0, u0 = gpu.accumulate_results(ph.med_densities, per, nb)
    np.testing.assert_allclose(dose, d0, rtol=3e-4, atol=1e-4 * d0.max())
    sel = d0 > 0.2 * d0.max()
    # batch-method uncertainty: squares were taken AFTER the sum over devices, so it matches the single-device one (a per-device
    # accumEndep() would give sqrt(n) times less)
    np.testing.assert_allclose(unc[sel], u0[sel], rtol=0.05)


def test_reference_user_code_on_two_gpus(gpu):
    """The reference's own omc_dosxyz.c with its batch loop on the library, -g 2 against -g 1 (omc_dosxyz.c:1237-1263 replaced)."""
    from tests.test_gpu_dropin import DROPIN, write_case
    from oracle import gen_fixtures as G
    if ndev() < 2:
        pytest.skip("needs 2 GPUs")
    if not (os.path.exists(DROPIN) and G.have_data()):
        pytest.skip("oracle/_ref/omc_dosxyz_dropin not built")
    work, stem, ph = write_case("golden_tissue4_6MV", 800008, 8)
    build.build()
    out = {}
    for g in (1, 2):
        r = subprocess.run([DROPIN, "-i", stem, "-o", f"dropin_g{g}", "-g", str(g)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
        assert f"GPUs: {g}" in r.stdout
        out[g] = P.read_3ddose(os.path.join(work, f"dropin_g{g}.3ddose"))
    (dims1, _, dose1, unc1), (dims2, _, dose2, unc2) = out[1], out[2]
    assert dims1 == dims2
    # same histories on either device count except the few that the drain kernel finishes on continued random streams (the
    # program runs with the default drain): the dose agrees voxel for voxel up to those, and in total
    close = np.isclose(dose2, dose1, rtol=5e-4, atol=5e-4 * dose1.max())
    assert close.mean() > 0.99, f"{(~close).sum()} of {close.size} voxels differ"
    np.testing.assert_allclose(dose2, dose1, rtol=0.1, atol=0.02 * dose1.max())
    assert abs(dose2.sum() - dose1.sum()) < 5e-4 * dose1.sum()
    sel = dose1 > 0.2 * dose1.max()
    np.testing.assert_allclose(unc2[sel], unc1[sel], rtol=0.1, atol=3e-3)


@pytest.mark.parametrize("n", [1, 2, 4])
def test_beamlet_passes_over_devices_gathered_in_beamlet_order(gpu, n):
    """omc_gpu_multi_run_beamlets(): passes of consecutive beamlets dealt to the devices, column slices gathered in beamlet order
    (omc_matrad.c:1416-1477) -- against one device running the same passes; beamlet b owns the same history ids either way, so
    the columns agree to the summation order of the fp32 dose atomics (entries at the threshold may fall on either side)."""
    if ndev() < n:
        pytest.skip(f"needs {n} GPUs")
    from ompmc_b200.api import MultiGpuTransport
    from tests.test_matrad import matrad_problem
    prob, ph, nb = matrad_problem(nbix=(2, 2), angles=(0.0, 120.0, 250.0))          # 12 beamlets
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    nh = 100000
    ref = [gpu.run_beamlets(b0 * nh, nh, 4, b0, 4, 0.05, ph.med_densities) for b0 in (0, 4, 8)]
    m = MultiGpuTransport(ndev=n)
    try:
        m.load_problem(prob)
        m.set_option("kernel", 1)
        jc, ir, val = m.run_beamlets(0, nh, 4, 0, nb, 0.05, ph.med_densities, per_pass=4)
    finally:
        m.close()
    assert len(jc) == nb + 1 and jc[0] == 0 and jc[-1] == len(ir) == len(val)
    for b in range(nb):
        rjc, rir, rval = ref[b // 4]
        k = b % 4
        d0 = np.zeros(ph.nvox); d0[rir[rjc[k]:rjc[k + 1]]] = rval[rjc[k]:rjc[k + 1]]
        r1 = ir[jc[b]:jc[b + 1]]
        assert (np.diff(r1) > 0).all()
        d1 = np.zeros(ph.nvox); d1[r1] = val[jc[b]:jc[b + 1]]
        np.testing.assert_allclose(d1, d0, rtol=2e-3, atol=0.051 * d0.max())
        assert abs(d1.sum() - d0.sum()) < 2e-3 * d0.sum()


def test_column_gather_of_one_rank_is_the_identity(gpu):
    mine = {0: (np.array([1, 5, 9], dtype=np.int64), np.array([0.5, 0.25, 1.0])), 1: (np.zeros(0, dtype=np.int64), np.zeros(0)),
            2: (np.array([7], dtype=np.int64), np.array([2.0]))}
    jc, ir, val = gpu.comm_gather_columns(3, mine)
    assert jc.tolist() == [0, 3, 3, 4] and ir.tolist() == [1, 5, 9, 7] and val.tolist() == [0.5, 0.25, 1.0, 2.0]

"""matRad path (BASELINE config 4): beamlet source initHistory(ibeamlet) + dose-influence column assembly.

CPU part: the oracle restatement against the reference's own omc_matrad.c, compiled here with a mex.h stand-in
(oracle/mexshim) -- bit-exact with the per-history Philox stream; the CSC assembly against a direct numpy
restatement of omc_matrad.c:1416-1477; beamlet sharding over 2 gloo ranks.
GPU part: lock-step kernel vs oracle (lock-step), wavefront kernels statistically.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from ompmc_b200 import matrad, problem as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def matrad_problem(nbix=(2, 2), angles=(0.0, 70.0, 200.0)):
    media = P.load_blob(P.golden("media_700_tissue4.blob"))
    ph = P.tissue_phantom((24, 10, 24), (0.8, 0.8, 0.8))
    bl = P.matrad_beamlets(ph, gantry_deg=angles, nbix=nbix, bixel_cm=0.5)
    prob = P.build_problem_matrad(media, ph, bl, ecut=0.7, pcut=0.01, cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
    return prob, ph, int(bl["mr_nbeamlets"][0])


def test_oracle_matches_reference_matrad_source(oracle_lib):
    from oracle import cpudrv
    if not cpudrv.have_ref(matrad=True):
        pytest.skip("oracle/_ref/libompmc_ref_matrad.so not built (no /root/reference here)")
    prob, ph, nb = matrad_problem()
    ref = cpudrv.RefTransport(matrad=True)
    ref.load_problem(prob); ref.set_rng("philox")
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob); oracle_lib.set_rng("philox")
    for b in (0, 3, 6, nb - 1):
        ref.reset_score(); oracle_lib.reset_score()
        ref.set_beamlet(b); oracle_lib.set_beamlet(b)
        r1 = ref.run_histories(5000 * b, 700, records=True)
        r2 = oracle_lib.run_histories(5000 * b, 700, records=True)
        assert np.array_equal(r1, r2)
        assert np.array_equal(ref.get_endep(), oracle_lib.get_endep())


class _OracleAsTransport:
    """Adapter giving the CPU oracle the run_batch/get_tallies/reset_tallies surface of GpuTransport."""

    def __init__(self, orc):
        self.o = orc

    def run_batch(self, first, n, ibeamlet=-1):
        self.o.set_beamlet(max(ibeamlet, 0))
        self.o.run_histories(first, n)
        self.o.accum_endep()

    def get_tallies(self):
        return self.o.get_accum()

    def beamlet_capacity(self):
        return 64

    def run_beamlets(self, first, nhist, nbatch, ib0, nb, rel_threshold, med_densities, ph=None):
        """CPU stand-in of omc_gpu_run_beamlets(): beamlet ib0+k owns history ids [first + k*nhist, +nhist); (jc, ir, val)."""
        ph = ph or self.ph
        nper = nhist // nbatch
        jc, irs, vals = [0], [], []
        for k in range(nb):
            self.reset_tallies(0)
            for ib in range(nbatch):
                self.run_batch(first + k * nhist + ib * nper, nper, ib0 + k)
            a, a2, _ = self.get_tallies()
            rows, v = matrad.beamlet_column(ph, a, a2, nhist, nbatch, rel_threshold)
            irs.append(rows); vals.append(v); jc.append(jc[-1] + len(rows))
        return np.asarray(jc, dtype=np.int64), np.concatenate(irs), np.concatenate(vals)

    def reset_tallies(self, which=0):
        if which == 0:
            self.o.reset_score()
        else:                                    # accum_endep only (omc_matrad.c:1482)
            a, a2, _ = self.o.get_accum()
            lib = self.o.lib
            import ctypes as C
            lib.orc_zero_accum.argtypes = []
            lib.orc_zero_accum()


def test_dose_influence_matrix_assembly(oracle_lib):
    prob, ph, nb = matrad_problem(nbix=(2, 1), angles=(0.0, 90.0))
    oracle_lib.set_num_threads(2)
    oracle_lib.load_problem(prob); oracle_lib.set_rng("philox")
    tr = _OracleAsTransport(oracle_lib)
    jc, ir, val = matrad.dose_influence_matrix(tr, ph, nb, "2000", "4", 0.01)
    assert jc[0] == 0 and jc[-1] == len(ir) == len(val) and len(jc) == nb + 1
    assert (np.diff(jc) > 0).all()
    # direct restatement of the column loop for beamlet 1
    oracle_lib.reset_score()
    oracle_lib.set_beamlet(1)
    for ib in range(4):
        oracle_lib.run_histories(1 * 2000 + ib * 500, 500)
        oracle_lib.accum_endep()
    a, a2, _ = oracle_lib.get_accum()
    dose, _ = P.accumulate_results(ph, a, a2, 2000, 4)           # nhist, not nperbatch (Q11)
    keep = np.nonzero(dose > dose.max() * 0.01)[0]
    np.testing.assert_array_equal(ir[jc[1]:jc[2]], keep)
    np.testing.assert_allclose(val[jc[1]:jc[2]], dose[keep], rtol=1e-12)
    assert (np.diff(ir[jc[1]:jc[2]]) > 0).all()                   # rows ascending, like mxCreateSparse wants
    oracle_lib.set_num_threads(1)


WORKER = r"""
import sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle.cpudrv import OracleTransport
from ompmc_b200 import matrad
from tests.test_matrad import matrad_problem, _OracleAsTransport
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prob, ph, nb = matrad_problem(nbix=(2, 1), angles=(0.0, 90.0, 180.0))
orc = OracleTransport(); orc.set_num_threads(1); orc.load_problem(prob); orc.set_rng("philox")
tr = _OracleAsTransport(orc); tr.ph = ph
jc, ir, val = matrad.dose_influence_matrix(tr, ph, nb, "600", "3", 0.02, rank, world, matrad.gather_columns_torch)
# the multi-beamlet flow: whole groups of consecutive beamlets per rank (6 beamlets in passes of 2 -> groups 0,2 on rank 0, 1 on rank 1)
jd, id_, vd = matrad.dose_influence_matrix_device(tr, ph, nb, "600", "3", 0.02, rank, world, matrad.gather_columns_torch, group=2)
assert [g for g in matrad.beamlet_groups(nb, world, 2) if g[2] == rank]
np.savez(%(out)r + f".{rank}.npz", jc=jc, ir=ir, val=val, jd=jd, id=id_, vd=vd)
dist.destroy_process_group()
"""


def test_beamlets_sharded_over_two_ranks(tmp_path, oracle_lib):
    out = str(tmp_path / "dij")
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29633", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    prob, ph, nb = matrad_problem(nbix=(2, 1), angles=(0.0, 90.0, 180.0))
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob); oracle_lib.set_rng("philox")
    jc, ir, val = matrad.dose_influence_matrix(_OracleAsTransport(oracle_lib), ph, nb, "600", "3", 0.02)
    for rank in (0, 1):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_array_equal(z["jc"], jc)
        np.testing.assert_array_equal(z["ir"], ir)
        np.testing.assert_allclose(z["val"], val, rtol=1e-12)
        # grouped sharding of the multi-beamlet flow: same history ids per beamlet -> the same matrix
        np.testing.assert_array_equal(z["jd"], jc)
        np.testing.assert_array_equal(z["id"], ir)
        np.testing.assert_allclose(z["vd"], val, rtol=1e-12)


@pytest.mark.gpu
def test_gpu_matrad_lockstep_vs_oracle(gpu, oracle_lib):
    prob, ph, nb = matrad_problem()
    gpu.load_problem(prob)
    gpu.set_option("kernel", 0)
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob); oracle_lib.set_rng("philox")
    for b in (0, 5, nb - 1):
        gpu.reset_tallies(); oracle_lib.reset_score()
        oracle_lib.set_beamlet(b)
        ro = oracle_lib.run_histories(9000 * b, 1500, records=True)
        rg = gpu.run_histories(9000 * b, 1500, records=True, ibeamlet=b)
        assert np.array_equal(rg["ir_start"], ro["ir_start"])
        same = (rg["ndraws"] == ro["ndraws"]) & (rg["ndeposit"] == ro["ndeposit"])
        assert same.mean() >= 0.998
        rel = np.abs(rg["edep"][same] - ro["edep"][same]) / np.maximum(ro["edep"][same], 1e-30)
        assert rel.max() < 1e-9
    from ompmc_b200.api import OmcGpuError
    with pytest.raises(OmcGpuError):
        gpu.run_histories(0, 10, ibeamlet=nb)          # beamlet index out of range


@pytest.mark.gpu
def test_gpu_matrad_dij_wavefront_vs_lockstep(gpu):
    prob, ph, nb = matrad_problem(nbix=(2, 1), angles=(0.0, 120.0))
    gpu.load_problem(prob)
    cols = {}
    for kernel in (0, 1):
        gpu.set_option("kernel", kernel)
        cols[kernel] = matrad.dose_influence_matrix(gpu, ph, nb, "200000", "10", 0.05)
    gpu.set_option("kernel", 1)
    jc0, ir0, v0 = cols[0]
    jc1, ir1, v1 = cols[1]
    assert len(jc0) == len(jc1) == nb + 1
    for b in range(nb):
        d0 = np.zeros(ph.nvox); d1 = np.zeros(ph.nvox)
        d0[ir0[jc0[b]:jc0[b + 1]]] = v0[jc0[b]:jc0[b + 1]]
        d1[ir1[jc1[b]:jc1[b + 1]]] = v1[jc1[b]:jc1[b + 1]]
        assert abs(d0.sum() - d1.sum()) < 0.02 * d0.sum()
        hot = d0 > 0.5 * d0.max()
        assert hot.sum() >= 3
        assert np.abs(d1[hot] / d0[hot] - 1.0).mean() < 0.05


@pytest.mark.gpu
def test_gpu_matrad_multi_beamlet_pass_vs_beamlet_loop(gpu):
    """omc_gpu_run_beamlets: all beamlets in one pass of the wavefront kernels + accumulateResults / threshold / CSC on the
    device, against the per-beamlet loop (same history ids per beamlet; the loop's drain makes the draws of the last few
    particles differ, so the comparison is statistical on the dense columns and exact on the structure)."""
    prob, ph, nb = matrad_problem(nbix=(2, 2), angles=(0.0, 120.0, 250.0))
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    jc0, ir0, v0 = matrad.dose_influence_matrix(gpu, ph, nb, "100000", "4", 0.05)
    jc1, ir1, v1 = matrad.dose_influence_matrix_device(gpu, ph, nb, "100000", "4", 0.05, group=5)      # 12 beamlets in passes of at most 5: 4, 4, 4
    assert len(jc0) == len(jc1) == nb + 1 and jc1[0] == 0 and jc1[-1] == len(ir1) == len(v1)
    for b in range(nb):
        r1 = ir1[jc1[b]:jc1[b + 1]]
        assert (np.diff(r1) > 0).all() and r1.min() >= 0 and r1.max() < ph.nvox       # rows ascending, like the reference's loop
        d0 = np.zeros(ph.nvox); d1 = np.zeros(ph.nvox)
        d0[ir0[jc0[b]:jc0[b + 1]]] = v0[jc0[b]:jc0[b + 1]]
        d1[r1] = v1[jc1[b]:jc1[b + 1]]
        assert (v1[jc1[b]:jc1[b + 1]] > 0.05 * d1.max() * (1 - 1e-12)).all()           # threshold test of omc_matrad.c:1420-1432
        assert abs(d0.sum() - d1.sum()) < 0.03 * d0.sum()
        hot = d0 > 0.5 * d0.max()
        assert hot.sum() >= 3 and np.abs(d1[hot] / d0[hot] - 1.0).mean() < 0.06
        assert abs(len(r1) - (jc0[b + 1] - jc0[b])) <= 0.15 * (jc0[b + 1] - jc0[b]) + 3
    # a single-beamlet group reproduces the same column as the same beamlet inside a larger group (history ids are per beamlet;
    # no drain in either): equal to fp32 summation order
    jcA, irA, vA = gpu.run_beamlets(3 * 100000, 100000, 4, 3, 1, 0.05, ph.med_densities)
    dA = np.zeros(ph.nvox); dA[irA] = vA
    dB = np.zeros(ph.nvox); dB[ir1[jc1[3]:jc1[4]]] = v1[jc1[3]:jc1[4]]
    np.testing.assert_allclose(dA, dB, rtol=2e-3, atol=0.051 * dB.max())
    assert abs(dA.sum() - dB.sum()) < 2e-3 * dB.sum()


def test_beamlet_group_plan_covers_every_beamlet_once_and_balances_ranks():
    """Sharding plan of the multi-beamlet pass (matrad.beamlet_groups; omc_matrad_b200.c computes the same): whole groups of
    consecutive beamlets per rank, every beamlet exactly once, groups no larger than asked, ranks balanced to one group size,
    ragged and degenerate cases (fewer beamlets than ranks, none at all)."""
    for n, world, group in ((320, 2, 64), (320, 1, 64), (6, 2, 64), (6, 2, 4), (1, 2, 64), (3, 8, 64), (0, 2, 64), (129, 2, 64), (320, 8, 64), (7, 3, 1)):
        plan = matrad.beamlet_groups(n, world, group)
        assert [b for b0, c, r in plan for b in range(b0, b0 + c)] == list(range(n))
        assert all(0 < c <= group and 0 <= r < world for _, c, r in plan)
        per = [sum(c for _, c, r in plan if r == k) for k in range(world)]
        if n:
            assert max(per) - min(per) <= max(c for _, c, _ in plan)
        assert [r for _, _, r in plan] == [j % world for j in range(len(plan))]


class _RecordingTransport:
    """Stands in for GpuTransport.run_beamlets(): records the passes and returns one non-zero per beamlet (row = beamlet)."""

    def __init__(self, capacity):
        self.capacity, self.calls = capacity, []

    def beamlet_capacity(self):
        return self.capacity

    def run_beamlets(self, first, nhist, nbatch, ib0, nb, rel, dens):
        self.calls.append((first, nhist, ib0, nb))
        return np.arange(nb + 1, dtype=np.int64), np.arange(ib0, ib0 + nb, dtype=np.int64), np.full(nb, 1.0)


def test_device_matrix_default_group_and_call_pattern():
    """dose_influence_matrix_device() without `group` takes the pass size from the transport (a stale second
    definition evaluated my[i:i+None]); with two ranks each owns whole groups of consecutive beamlets, one pass per group."""
    ph = P.tissue_phantom((6, 6, 6), (1.0, 1.0, 1.0))
    tr = _RecordingTransport(capacity=4)
    jc, ir, val = matrad.dose_influence_matrix_device(tr, ph, 10, "1000", "2", 0.01)
    assert [c[2:] for c in tr.calls] == [(0, 4), (4, 4), (8, 2)]
    assert [c[0] for c in tr.calls] == [0, 4 * 1000, 8 * 1000] and all(c[1] == 1000 for c in tr.calls)
    assert np.array_equal(jc, np.arange(11)) and np.array_equal(ir, np.arange(10))
    plan = matrad.beamlet_groups(10, 2, 4)
    for rank in (0, 1):
        tr = _RecordingTransport(capacity=4)
        merged = {}
        matrad.dose_influence_matrix_device(tr, ph, 10, "1000", "2", 0.01, rank, 2, gather=lambda mine: merged.update(mine) or
                                            {b: (np.zeros(0, np.int64), np.zeros(0)) for b in range(10)})
        assert [(c[2], c[3]) for c in tr.calls] == [(b0, n) for b0, n, owner in plan if owner == rank]
        assert len(tr.calls) == 2                                # whole groups: 10 beamlets = 4 passes over 2 ranks, not 5 + 5
        assert sorted(merged) == [b for b0, n, owner in plan if owner == rank for b in range(b0, b0 + n)]

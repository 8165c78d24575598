"""world_size-2 gloo test of the N > 1 host logic (ompmc_b200/dist.py) with the CPU oracle standing in
for the GPU transport: sharded batches + all-reduce before accumEndep() == single-process batches."""
import os
import subprocess
import sys

import numpy as np

from ompmc_b200 import dist as odist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle.cpudrv import OracleTransport
from oracle.gen_fixtures import golden_problem
from ompmc_b200 import dist as odist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prob, ph, cfg = golden_problem("golden_water700_6MV")
tr = OracleTransport(); tr.set_num_threads(1); tr.load_problem(prob); tr.set_rng("philox")
tr.accum_batch = tr.accum_endep
from ompmc_b200.api import shard_range_c
for ib in range(3):
    # (the slice the C library would take on this rank: same rule)
    assert shard_range_c(100 + ib * 401, 401, rank, world) == odist.shard_range(100 + ib * 401, 401, rank, world)
    odist.run_batch_sharded(tr, 100 + ib * 401, 401, rank, world, odist.allreduce_cpu_grid)
a, a2, _ = tr.get_accum()
np.savez(%(out)r + f".{rank}.npz", a=a, a2=a2)
dist.destroy_process_group()
"""


def test_shard_range_partitions():
    """The Python rule and the rule INSIDE the library (omc_gpu_shard_range: what omc_gpu_run_batch() applies on every rank of an
    NCCL communicator, include/ompmc_b200.h) are the same partition."""
    from ompmc_b200.api import shard_range_c
    for n in (0, 1, 7, 401, 1000, 67108864 * 8 + 5):
        for world in (1, 2, 3, 8):
            parts = [odist.shard_range(50, n, r, world) for r in range(world)]
            assert parts == [shard_range_c(50, n, r, world) for r in range(world)]
            assert sum(p[1] for p in parts) == n
            pos = 50
            for lo, cnt in parts:
                assert lo == pos
                pos += cnt
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


def test_two_rank_batches_equal_single_rank(tmp_path, oracle_lib):
    out = str(tmp_path / "res")
    code = WORKER % dict(root=ROOT, out=out)
    script = tmp_path / "w.py"
    script.write_text(code)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    from oracle.gen_fixtures import golden_problem
    prob, ph, cfg = golden_problem("golden_water700_6MV")
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    for ib in range(3):
        oracle_lib.run_histories(100 + ib * 401, 401)
        oracle_lib.accum_endep()
    a, a2, _ = oracle_lib.get_accum()
    for rank in (0, 1):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_allclose(z["a"], a, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(z["a2"], a2, rtol=1e-12, atol=1e-300)


PIPE_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle.cpudrv import OracleTransport
from oracle.gen_fixtures import golden_problem
from ompmc_b200 import dist as odist


class PipelinedChecker:
    # CPU stand-in with the pipelined surface of GpuTransport (start_batch / finish_batches / completed_batches /
    # accum_batch, include/ompmc_b200.h): a started batch becomes "completed" only when the NEXT one is started (or
    # finish_batches is called), exactly the contract the multi-rank flow relies on; get_endep/set_endep address the
    # oldest completed batch grid, which is what the all-reduce acts on.
    def __init__(self, orc):
        self.o, self.pending, self.done = orc, None, []
        self.a = self.a2 = None
    def start_batch(self, first, n, ibeamlet=-1):
        self.finish_batches()
        if n > 0:
            self.o.run_histories(first, n)
        self.pending = self.o.get_endep().copy()
        self.o.set_endep(np.zeros_like(self.pending))
    def finish_batches(self):
        if self.pending is not None:
            self.done.append(self.pending); self.pending = None
    def completed_batches(self):
        return len(self.done)
    def get_endep(self):
        return self.done[0]
    def set_endep(self, g):
        self.done[0] = np.array(g, dtype=np.float64)
    def accum_batch(self):
        g = self.done.pop(0)
        if self.a is None:
            self.a = np.zeros_like(g); self.a2 = np.zeros_like(g)
        self.a += g; self.a2 += g * g


dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prob, ph, cfg = golden_problem("golden_water700_6MV")
orc = OracleTransport(); orc.set_num_threads(1); orc.load_problem(prob); orc.set_rng("philox")
tr = PipelinedChecker(orc)
first = 100
for n in (401, 1, 0, 777):                   # incl. a batch smaller than the rank count and an empty one
    odist.start_batch_sharded(tr, first, n, rank, world, odist.allreduce_cpu_grid)
    first += n
odist.finish_batches_sharded(tr, rank, world, odist.allreduce_cpu_grid)
assert tr.completed_batches() == 0
np.savez(%(out)r + f".{rank}.npz", a=tr.a, a2=tr.a2)
dist.destroy_process_group()
"""


def test_two_rank_pipelined_batches_equal_single_rank(tmp_path, oracle_lib):
    """start_batch_sharded / finish_batches_sharded (the flow bench.py uses): every rank completes batch k-1 inside its start
    of batch k, so the collectives line up, and the accumulated statistics equal the single-process batch loop."""
    out = str(tmp_path / "pres")
    script = tmp_path / "pw.py"
    script.write_text(PIPE_WORKER % dict(root=ROOT, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29612", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    from oracle.gen_fixtures import golden_problem
    prob, ph, cfg = golden_problem("golden_water700_6MV")
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    oracle_lib.reset_score()
    first = 100
    for n in (401, 1, 0, 777):
        if n:
            oracle_lib.run_histories(first, n)
        oracle_lib.accum_endep()
        first += n
    a, a2, _ = oracle_lib.get_accum()
    for rank in (0, 1):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_allclose(z["a"], a, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(z["a2"], a2, rtol=1e-12, atol=1e-300)

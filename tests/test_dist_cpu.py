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
for ib in range(3):
    odist.run_batch_sharded(tr, 100 + ib * 401, 401, rank, world, odist.allreduce_cpu_grid)
a, a2, _ = tr.get_accum()
np.savez(%(out)r + f".{rank}.npz", a=a, a2=a2)
dist.destroy_process_group()
"""


def test_shard_range_partitions():
    for n in (0, 1, 7, 401, 1000):
        for world in (1, 2, 3, 8):
            parts = [odist.shard_range(50, n, r, world) for r in range(world)]
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

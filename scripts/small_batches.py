"""Throughput of omc_gpu_run_batch() as a function of the batch size (the ring of dose grids hides the ~20 ms tail of a batch
behind the next batches).  MEASUREMENT SCRIPT.  usage (GPU box): python scripts/small_batches.py"""
import sys, time
sys.path.insert(0, '.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
g.run_batch(10**12, 4000000); g.synchronize(); g.reset_tallies()
for per, nb in ((1 << 26, 3), (1 << 24, 12), (1 << 22, 24), (1 << 20, 40), (100000, 10)):
    g.reset_tallies(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ib in range(nb):
        g.run_batch(ib * per, per)
    g.synchronize()
    dt = time.perf_counter() - t0
    print('batches of', per, 'x', nb, ': %.4g hist/s' % (per * nb / dt), '%.1f ms per batch' % (1e3 * dt / nb), flush=True)

# round 2: ncu capture of one steady-state wave on the 1 mm grid (8.1e7 voxels), and the small-batch rates
set -x
cat > /tmp/cap1mm.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv', 1, 1.0)
g = GpuTransport(0); g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('overlap', 0); g.set_option('pool_size', 1<<23)
g.run_histories(0, 12000000); g.synchronize()
print(g.counters())
PY
timeout 900 ncu --set full --clock-control none -k regex:"esize_kernel|edo_kernel|misc_kernel" -s 300 -c 4 -o gpurun_out/prof_r02_1mm python /tmp/cap1mm.py > gpurun_out/ncu_r02_1mm.log 2>&1
tail -2 gpurun_out/ncu_r02_1mm.log
cat > /tmp/small.py <<'PY'
import sys, time
sys.path.insert(0,'.')
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
PY
python /tmp/small.py

set -x
python -m pytest tests/test_gpu_wavefront.py -x -q 2>&1 | tail -15
cat > /tmp/sweep.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
tag = sys.argv[1]
for wl in ('prostate6mv',):
    prob, ph, w = bench.build_workload(wl)
    g = GpuTransport(0)
    g.load_problem(prob)
    n = 4000000
    g.set_option('kernel', 1)
    g.run_histories(0, 200000); g.synchronize()
    for pool in (1<<20, 1<<22, 1<<23):
        for iters in (1,2,4,8):
            for cross in (8, 16, 32):
                g.set_option('pool_size', pool); g.set_option('electron_iters', iters); g.set_option('max_cross', cross)
                g.reset_tallies()
                t=time.time(); g.run_histories(0, n); g.synchronize(); dt=time.time()-t
                c=g.counters()
                print(tag, wl, 'pool',pool,'iters',iters,'cross',cross, '%.3g hist/s'%(n/dt), 'launches',c['kernel_launches'], {k:round(v/n,2) for k,v in c.items() if k.endswith('steps') or k=='deposits'}, flush=True)
    g.close()
PY
python /tmp/sweep.py mb3
OMC_NVCC_FLAGS="-DOMC_WAVE_MINBLOCKS=2" python ompmc_b200/build.py --force > /dev/null
python /tmp/sweep.py mb2
OMC_NVCC_FLAGS="-DOMC_WAVE_MINBLOCKS=4" python ompmc_b200/build.py --force > /dev/null
python /tmp/sweep.py mb4

set -x
python -m pytest tests/test_gpu_wavefront.py -x -q 2>&1 | tail -30
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
for wl in ('prostate6mv','water6mv'):
    prob, ph, w = bench.build_workload(wl)
    g = GpuTransport(0)
    g.load_problem(prob)
    n = 2000000
    g.set_option('kernel', 0); g.reset_tallies()
    t=time.time(); g.run_histories(0, n); g.synchronize(); dt=time.time()-t
    print(wl, 'lockstep', n/dt, 'hist/s')
    g.set_option('kernel', 1)
    g.run_histories(0, 100000); g.synchronize()
    for pool in (1<<18, 1<<20, 1<<22):
        for iters in (1,2,4,8):
            for cross in (16, 64):
                g.set_option('pool_size', pool); g.set_option('electron_iters', iters); g.set_option('max_cross', cross)
                g.reset_tallies()
                t=time.time(); g.run_histories(0, n); g.synchronize(); dt=time.time()-t
                c=g.counters()
                print(wl, 'wave pool',pool,'iters',iters,'cross',cross, '%.3g hist/s'%(n/dt), 'launches',c['kernel_launches'], {k:round(v/n,2) for k,v in c.items() if k.endswith('steps') or k=='deposits'}, flush=True)
    g.close()
PY

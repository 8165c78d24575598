set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_r1b.err | grep '^{' | tee gpurun_out/bench_r1b.json
cat > /tmp/steady.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
stream = torch.cuda.ExternalStream(g.stream_ptr())
for pool in (1<<20, 3<<19, 1<<21, 3<<20, 1<<22):
    g.set_option('pool_size', pool)
    g.run_histories(0, 4000000); g.synchronize()
    for n in (20000000,):
        g.reset_tallies()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize()
        c=g.counters()
        ms=e0.elapsed_time(e1)
        print('pool',pool,'n',n,'%.3g hist/s'%(n/ms*1e3), 'gpu ms %.1f'%ms, 'launches', c['kernel_launches'], flush=True)
PY
python /tmp/steady.py

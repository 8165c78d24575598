set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
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
g.run_histories(0, 100000); g.synchronize()
for n in (2000000, 8000000, 20000000):
  for pool, cross, every, drain in [(1<<21,16,16,32768),(1<<21,16,16,0),(1<<21,16,32,100000),(1<<22,16,16,32768),(1<<21,32,16,32768),(1<<20,16,16,32768)]:
    g.set_option('pool_size', pool); g.set_option('max_cross', cross); g.set_option('check_every', every); g.set_option('drain_threshold', drain)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print('n',n,'pool',pool,'cross',cross,'every',every,'drain',drain, '%.3g hist/s wall'%(n/dt), 'gpu ms %.1f'%e0.elapsed_time(e1), 'launches', c['kernel_launches'], flush=True)
PY
python /tmp/steady.py

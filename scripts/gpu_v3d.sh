set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
cat > /tmp/steady.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('pool_size', 1<<21)
stream = torch.cuda.ExternalStream(g.stream_ptr())
g.run_histories(0, 4000000); g.synchronize()
for n in (8000000, 20000000):
  for overlap, every, cross in [(1,16,16),(0,16,16),(1,32,16),(1,16,32),(1,16,64)]:
    g.set_option('overlap', overlap); g.set_option('check_every', every); g.set_option('max_cross', cross)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print('n',n,'overlap',overlap,'every',every,'cross',cross,'%.3g hist/s wall'%(n/dt), 'gpu ms %.1f'%e0.elapsed_time(e1), 'launches', c['kernel_launches'], flush=True)
PY
python /tmp/steady.py
cat > /tmp/steady2.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('overlap', 0); g.set_option('pool_size', 1<<21)
g.run_histories(0, 20000000); g.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 60 --csv --log-file gpurun_out/launches_r01_wave3d.csv python /tmp/steady2.py > /dev/null 2>&1

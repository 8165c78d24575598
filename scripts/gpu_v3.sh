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
n = int(sys.argv[1])
g.run_histories(0, 100000); g.synchronize()
for pool, cross, every in [(1<<21,16,16),(1<<22,16,16),(1<<21,8,16),(1<<21,32,16),(1<<20,16,16)]:
    g.set_option('pool_size', pool); g.set_option('max_cross', cross); g.set_option('check_every', every)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print('n',n,'pool',pool,'cross',cross,'every',every, '%.3g hist/s wall'%(n/dt), 'gpu ms', e0.elapsed_time(e1), 'launches', c['kernel_launches'], flush=True)
    if len(sys.argv) > 2: break
PY
python /tmp/steady.py 20000000
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 50 --csv --log-file gpurun_out/launches_r01_wave3.csv python /tmp/steady.py 20000000 one > /dev/null 2>&1

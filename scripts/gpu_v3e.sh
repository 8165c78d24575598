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
for n in (8000000, 20000000, 20000000):
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print(sys.argv[1], 'n',n,'%.3g hist/s wall'%(n/dt), 'gpu ms %.1f'%e0.elapsed_time(e1), 'launches', c['kernel_launches'], flush=True)
t=time.time()
for i in range(3): g.load_problem(prob)
print('load_problem s', (time.time()-t)/3)
t=time.time()
for i in range(3): g.get_tallies()
print('get_tallies s', (time.time()-t)/3)
PY
python /tmp/steady.py mixed
OMC_NVCC_FLAGS="-DOMC_WAVE_F32=0" python ompmc_b200/build.py --force > /dev/null
python /tmp/steady.py f64
python ompmc_b200/build.py --force > /dev/null

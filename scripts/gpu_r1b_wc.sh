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
def run(tag, n=20000000):
    g.run_histories(0, 4000000); g.synchronize()
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize()
    c=g.counters(); ms=e0.elapsed_time(e1)
    print(tag,'n',n,'%.3g hist/s'%(n/ms*1e3), 'gpu ms %.1f'%ms, 'launches', c['kernel_launches'], 'psteps/h %.1f esteps/h %.1f'%(c['photon_steps']/n, c['electron_steps']/n), flush=True)
g.set_option('photon_tracking', 0); run('march')
g.set_option('photon_tracking', 1)
for mv in (4, 8, 16):
    g.set_option('max_virtual', mv); run('woodcock mv %d'%mv)
g.set_option('max_virtual', 8)
for pool in (1<<21, 3<<20):
    g.set_option('pool_size', pool); run('woodcock pool %d'%pool)
PY
python /tmp/steady.py

"""Fixed cost of one omc_gpu_run_histories() call (ramp-up + tail + drain) vs options."""
import sys
sys.path.insert(0, '.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
stream = torch.cuda.ExternalStream(g.stream_ptr())
g.run_histories(0, 4000000); g.synchronize()
def run(n, **opt):
    for k, v in opt.items(): g.set_option(k, v)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize()
    c = g.counters(); ms = e0.elapsed_time(e1)
    print('n', n, opt, 'ms %.1f' % ms, '%.3g hist/s' % (n / ms * 1e3), 'launches', c['kernel_launches'], flush=True)
for n in (250000, 1000000, 4000000, 16000000, 32000000):
    run(n)
for dt in (0, 8192, 131072, 524288, 2000000):
    run(16000000, drain_threshold=dt)
g.set_option('drain_threshold', 32768)
for ce in (4, 8, 32):
    run(16000000, check_every=ce)
g.set_option('check_every', 16)
for pool in (1 << 21, 3 << 21, 1 << 23):
    run(16000000, pool_size=pool)

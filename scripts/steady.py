"""Steady-state throughput of the wavefront kernels on the bench workload (CUDA events around run_histories).
usage: steady.py TAG [n] [opt=value ...]"""
import sys
sys.path.insert(0, '.')
import bench, torch
from ompmc_b200.api import GpuTransport
tag = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 20000000
opts = dict(a.split('=') for a in sys.argv[2:] if '=' in a)
prob, ph, w = bench.build_workload(opts.pop('workload', 'prostate6mv'), int(opts.pop('nsplit', 1)))
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
for k, v in opts.items():
    g.set_option(k, int(v))
stream = torch.cuda.ExternalStream(g.stream_ptr())
g.run_histories(0, 4000000); g.synchronize()
for rep in range(2):
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize()
    c = g.counters(); ms = e0.elapsed_time(e1)
    print(tag, opts, 'n', n, '%.4g hist/s' % (n / ms * 1e3), 'gpu ms %.1f' % ms, 'launches', c['kernel_launches'],
          'psteps/h %.1f esteps/h %.1f edep/h %.5f' % (c['photon_steps'] / n, c['electron_steps'] / n, g.get_endep()[1:].sum() / n), flush=True)

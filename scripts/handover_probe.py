"""Pipelined batches (omc_gpu_run_batch) with and without the straggler hand-over, over batch sizes.
usage: handover_probe.py [workload]"""
import sys
sys.path.insert(0, '.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else 'prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
stream = torch.cuda.ExternalStream(g.stream_ptr())
g.run_histories(0, 4000000); g.synchronize()
for per, nb in ((1000000, 20), (4000000, 10), (16777216, 6), (67108864, 4)):
    for ho in (0, 1):
        g.set_option('handover', ho)
        g.reset_tallies()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ib in range(nb):
            g.run_batch(ib * per, per)
        g.finish_batches()
        e1.record(stream); g.synchronize()
        ms = e0.elapsed_time(e1)
        c = g.counters()
        a, a2, ensrc = g.get_tallies()
        n = nb * per
        print('per', per, 'nb', nb, 'handover', ho, 'ms/batch %.1f' % (ms / nb), '%.4g hist/s' % (n / ms * 1e3), 'launches', c['kernel_launches'],
              'handovers', c['handovers'], 'handed_over', c['handed_over'], 'edep/h %.6f' % (a[1:].sum() / n), 'errors', c['errors'], flush=True)
g.set_option('handover', 0)

"""How long are the longest particle lineages?  (waves needed after the last history was started, drain off)"""
import sys, struct
sys.path.insert(0, '.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
g.set_option('trace', 1)
g.set_option('check_every', 64)
g.set_option('drain_threshold', 0)
g.run_histories(0, 2000000); g.synchronize()
g.set_option('trace', 0)
g.set_option('drain_threshold', 8192)
g.reset_tallies()
g.run_histories(0, 16000000); g.synchronize()
c = g.counters()
print('drain stats', c.get('reserved'))
if c.get('reserved'):
    r = c['reserved']
    print('longest drained chain (electron steps):', r[0], ' chains > 20000 steps:', r[1], ' energy', struct.unpack('d', struct.pack('Q', r[2]))[0], 'ir/iq', r[3] & 0xffffffff, (r[3] >> 32) - 2)

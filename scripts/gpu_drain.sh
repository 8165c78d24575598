cat > /tmp/dr.py <<'PY'
import sys, time, struct
sys.path.insert(0,'.')
import bench, numpy as np
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
g.set_option('pool_size', 1<<21); g.set_option('check_every', 16)
for thr in (32768, 100000, 32768, 8192):
    g.set_option('drain_threshold', thr)
    g.reset_tallies()
    t=time.time(); g.run_histories(0, 8000000); g.synchronize(); dt=time.time()-t
    c=g.counters(); r=c.get('reserved',[0,0,0,0])
    e=struct.unpack('d', struct.pack('Q', r[2]))[0]
    ir=r[3]&0xffffffff; iq=(r[3]>>32)-2
    med = prob['region_med'][ir] if ir < len(prob['region_med']) else None
    print('thr',thr,'time %.3f'%dt,'max chain',r[0],'n>20000',r[1],'example e',e,'ir',ir,'iq',iq,'med',med,'rhof', prob['region_rhof'][ir] if med is not None else None, 'esteps/hist', c['electron_steps']/8e6)
PY
python /tmp/dr.py

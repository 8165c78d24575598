cat > /tmp/tr.py <<'PY'
import sys, time
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
g.set_option('pool_size', 1<<21); g.set_option('electron_iters', 1); g.set_option('max_cross', 16); g.set_option('check_every', 50); g.set_option('trace', 1)
t=time.time(); g.run_histories(0, 20000000); g.synchronize(); print('time', time.time()-t)
PY
python /tmp/tr.py 2>&1 | tail -70

set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/matrad_probe.py 1000000 2>&1 | tail -3
python - <<'PY'
import sys
sys.path.insert(0,'.')
import bench, torch, time
from ompmc_b200.api import GpuTransport
for ns in (20, 5):
    prob, ph, w = bench.build_workload('prostate6mv', ns)
    g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
    g.run_histories(0, 200000); g.synchronize()
    n = 2000000
    t = time.time(); g.run_histories(10000000, n); g.synchronize(); dt = time.time() - t
    print('nsplit', ns, 'wavefront %.3g hist/s' % (n / dt), flush=True)
    g.set_option('kernel', 0)
    n = 200000
    t = time.time(); g.run_histories(20000000, n); g.synchronize(); dt = time.time() - t
    print('nsplit', ns, 'lockstep %.3g hist/s' % (n / dt), flush=True)
    g.close()
PY

cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0,'.')
from oracle.gen_fixtures import golden_problem
from ompmc_b200.api import GpuTransport
prob, ph, cfg = golden_problem('golden_water700_6MV')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1); g.set_option('pool_size', 4096)
g.run_histories(0, 2000); g.synchronize(); print('ok', g.counters())
PY
compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in \|^=========                in" | head -60

cat > /tmp/diag.py <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from test_gpu_wavefront import CASES, make_problem
from ompmc_b200.api import GpuTransport
g = GpuTransport(0)
for idx in (0, 1):
    prob, ph = make_problem(CASES[idx][1])
    g.load_problem(prob)
    n = 4000000
    def run(tag, kernel, **opt):
        g.set_option('kernel', kernel)
        for k, v in opt.items(): g.set_option(k, v)
        g.reset_tallies(); g.run_histories(0, n); g.synchronize()
        c = g.counters(); e = g.get_endep()[1:].sum()
        print(CASES[idx][0], tag, 'psteps/h %.3f esteps/h %.3f dep/h %.3f edep/h %.5f' % (c['photon_steps']/n, c['electron_steps']/n, c['deposits']/n, e/n), flush=True)
    run('lockstep', 0)
    run('wave march nodrain', 1, photon_tracking=0, drain_threshold=0)
    run('wave march drain-all', 1, photon_tracking=0, drain_threshold=1<<30, pool_size=1<<22)
    run('wave woodcock', 1, photon_tracking=1, drain_threshold=32768)
PY
python /tmp/diag.py

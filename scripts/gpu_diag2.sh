cat > /tmp/diag2.py <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from test_gpu_wavefront import CASES, make_problem
from ompmc_b200 import problem as P
from ompmc_b200.api import GpuTransport
g = GpuTransport(0)
def run(name, tag, n, kernel, first=0, **opt):
    g.set_option('kernel', kernel)
    for k, v in opt.items(): g.set_option(k, v)
    g.reset_tallies(); g.run_histories(first, n); g.synchronize()
    c = g.counters(); e = g.get_endep()[1:].sum()
    print(name, tag, 'n', n, 'psteps/h %.4f esteps/h %.3f dep/h %.4f edep/h %.6f' % (c['photon_steps']/n, c['electron_steps']/n, c['deposits']/n, e/n), flush=True)
# 1) primary flight only: mono 6 MeV photons, pcut = 5.9 MeV, ecut = 10 MeV: everything but the primary is absorbed on the spot
cfg = dict(CASES[0][1]); cfg['spec'] = None; cfg['mono'] = 6.0
media = P.load_blob(P.golden(cfg["mset"])); ph = cfg["ph"]()
prob = P.build_problem(media, ph, ecut=10.0, pcut=5.9, collimator=cfg["coll"], ssd=cfg["ssd"], charge=0, cdfinv=None, mono_energy=6.0)
g.load_problem(prob)
run('primary-only', 'lockstep', 1000000, 0)
run('primary-only', 'march', 1000000, 1, photon_tracking=0)
run('primary-only', 'woodcock', 1000000, 1, photon_tracking=1)
# 2) high statistics: march vs woodcock
for idx in (0, 1):
    prob, ph = make_problem(CASES[idx][1])
    g.load_problem(prob)
    for rep in range(2):
        run(CASES[idx][0], 'march', 40000000, 1, first=rep*100000000, photon_tracking=0)
        run(CASES[idx][0], 'woodcock', 40000000, 1, first=rep*100000000, photon_tracking=1)
PY
python /tmp/diag2.py

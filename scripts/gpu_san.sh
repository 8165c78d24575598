# compute-sanitizer on small runs of the production path: memcheck + racecheck of the wave kernels, pipelined batches over the
# ring of dose grids, nsplit 20, the multi-beamlet pass, the sampler hooks ("compute-sanitizer on small runs")
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0,'.')
import numpy as np
from oracle.gen_fixtures import golden_problem
from ompmc_b200.api import GpuTransport
from tests.test_matrad import matrad_problem
from tests import sampler_cases as S
prob, ph, cfg = golden_problem('golden_tissue4_6MV')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1); g.set_option('pool_size', 4096)
g.run_histories(0, 2000); g.synchronize(); print('single call ok', g.counters()['histories'])
g.reset_tallies()
for ib in range(7):
    g.run_batch(ib * 1500, 1500)
a, a2, e = g.get_tallies(); print('7 pipelined batches ok', a[1:].sum() > 0)
g.set_nsplit(20); g.reset_tallies(); g.run_batch(0, 300); g.run_batch(300, 300); g.synchronize(); g.set_nsplit(1); print('nsplit 20 ok')
mp, mph, nb = matrad_problem(nbix=(2, 2), angles=(0.0, 120.0))
g.load_problem(mp); g.set_option('kernel', 1)
jc, ir, val = g.run_beamlets(0, 2000, 4, 0, nb, 0.05, mph.med_densities); print('beamlet pass ok', jc[-1])
sp, sph = S.problem_tissue4()
g.load_problem(sp); g.set_option('kernel', 1)
for which in (S.MSDIST, S.SSCAT, S.COMPTON, S.MOLLER):
    pass
o = g.test_samplers(S.MSDIST, S.msdist_inputs(S.MSDIST_GROUPS[4], 4096, True), first_history=5); print('samplers ok', np.isfinite(o).all())
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in \|^=========                in" | tail -25
done

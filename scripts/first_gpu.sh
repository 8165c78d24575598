set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
from ompmc_b200 import problem as P
from ompmc_b200.api import GpuTransport
media = P.load_blob(P.golden('media_700_water.blob'))
ph = P.water_phantom('H2O700ICRU')
prob = P.build_problem(media, ph, ecut=0.7, pcut=0.01, collimator=(-5,5,-5,5), ssd=100.0, cdfinv=(media['cdfinv1_mohan6'], media['cdfinv2_mohan6']))
g = GpuTransport(0)
g.load_problem(prob)
for n in (100000, 1000000, 4000000):
    g.reset_tallies()
    t=time.time(); g.run_histories(0, n); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print(n, 'hist', dt, 's', n/dt, 'hist/s', {k:v/n for k,v in c.items()})
PY

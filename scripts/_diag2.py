import sys
sys.path.insert(0, '.')
import numpy as np
from ompmc_b200.api import GpuTransport
from tests import sampler_cases as S
prob, ph = S.problem_tissue4()
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
for v in (0, 1, 2):
    for gi, grp in enumerate(S.MSDIST_GROUPS):
        inp = S.msdist_inputs(grp, 8192, False)
        print('variant', v, 'group', gi, flush=True)
        o = g.test_samplers(S.MSDIST | (v << 8), inp, first_history=10_000_000 * (gi + 1))

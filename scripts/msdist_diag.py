"""Diagnostic for tests/test_gpu_production_samplers.py::test_msdist: dumps the outputs of the three device variants of the
condensed-history step (production block-draw fp32, word-by-word fp32, fp64) for every group, to be compared offline with the
oracle.  usage (GPU box): python scripts/msdist_diag.py"""
import sys
sys.path.insert(0, '.')
import numpy as np
from ompmc_b200.api import GpuTransport
from tests import sampler_cases as S
prob, ph = S.problem_tissue4()
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
N = 20000
out = {}
for v in range(3):
    for gi, grp in enumerate(S.MSDIST_GROUPS):
        inp = S.msdist_inputs(grp, N, False)
        o = g.test_samplers(S.MSDIST | (v << 8), inp, first_history=10_000_000 * (gi + 1))
        out[f'v{v}_g{gi}_f'] = o[:, :6].astype(np.float32)
        out[f'v{v}_g{gi}_omc'] = 1.0 - o[:, 6]
    print('variant', v, 'done', flush=True)
np.savez_compressed('gpurun_out/msdist_diag.npz', **out)
print('saved')

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
N = 50000
for v in range(3):
    out = {}
    for gi, grp in enumerate(S.MSDIST_GROUPS):
        inp = S.msdist_inputs(grp, N, False)
        print('variant', v, 'group', gi, flush=True)
        out[f'g{gi}'] = g.test_samplers(S.MSDIST | (v << 8), inp, first_history=10_000_000 * (gi + 1)).astype(np.float64)
    np.savez_compressed(f'gpurun_out/msdist_diag_v{v}.npz', **out)
    print('saved variant', v, flush=True)

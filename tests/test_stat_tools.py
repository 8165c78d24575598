"""The statistics the GPU parity tests rely on, checked on synthetic samples (CPU): the shift-tolerant KS distance of
tests/test_gpu_production_samplers.py and the gamma index / batch statistics of tests/parity_tools.py."""
import numpy as np

from tests.parity_tools import batch_stats, gamma_pass
from tests.test_gpu_production_samplers import ks_distance_tol


def test_tolerant_ks_forgives_atom_shifts_but_not_real_differences():
    rng = np.random.default_rng(1)
    n = 100_000
    atoms = np.array([0.0, 1.0e-3, 2.5e-3, 7.0e-3])

    def sample(shift, scale=1.0):
        # 30 % of the mass on four atoms (as mscat()'s tabulated u values and the no-scattering amplitude put it), the rest continuous
        x = rng.exponential(5.0e-3 * scale, n)
        k = rng.random(n) < 0.3
        x[k] = rng.choice(atoms, k.sum()) * (1.0 + shift)
        return x
    a, b = sample(0.0), sample(1.0e-6)                      # atoms displaced by fp32-level rounding
    from scipy import stats
    assert stats.ks_2samp(a, b).statistic > 0.05           # the plain distance reports the atoms' mass ...
    assert ks_distance_tol(a, b) < 0.008                   # ... the tolerant one sees two samples of one distribution
    assert ks_distance_tol(a, sample(0.0, scale=1.05)) > 0.012      # a 5 % wider continuous part is still seen
    assert ks_distance_tol(a, sample(0.02)) > 0.02         # and so are atoms displaced by 2 %


def test_batch_statistics_and_gamma_index():
    rng = np.random.default_rng(2)
    nb, shape = 40, (12, 10, 8)
    truth = 1.0 + np.indices(shape).sum(axis=0) / 10.0
    batches = truth[None] * (1.0 + 0.01 * rng.standard_normal((nb,) + shape))
    accum = np.concatenate([[0.0], batches.sum(axis=0).reshape(-1)])
    accum2 = np.concatenate([[0.0], (batches ** 2).sum(axis=0).reshape(-1)])
    mean, var = batch_stats(accum, accum2, nb)
    assert np.allclose(mean, batches.mean(axis=0).reshape(-1))
    assert np.allclose(var, batches.var(axis=0, ddof=1).reshape(-1) / nb, rtol=1e-9)
    ref = truth
    ok, nvox, gmax = gamma_pass(ref, ref * 1.002, (10.0, 10.0, 10.0))          # 0.2 % off everywhere: passes 1 %
    assert ok == 1.0 and nvox == ref.size and gmax < 0.5
    bad, _, _ = gamma_pass(ref, ref * 1.03, (10.0, 10.0, 10.0))                # 3 % off on a 10 mm grid: 1 mm of distance cannot save it
    assert bad < 0.5

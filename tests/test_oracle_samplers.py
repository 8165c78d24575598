"""Pins the oracle's restated samplers to the UNMODIFIED reference at function level (CPU).

ref_test_samplers() (oracle/ref_harness.c) calls the reference's own computeDrange / computeEloss / msdist (+ mscat +
spinRejection) / sscat / compton / moller (src/ompmc.c:3979, 4016, 3787, 3170, 1670, 4359) on explicit inputs with the
per-history Philox stream; orc_test_samplers() does the same with the restatement of oracle/omc_oracle.c.  Same draw order,
same arithmetic -> bit-identical outputs.  The GPU test tests/test_gpu_production_samplers.py then compares the production
CUDA samplers with the oracle (the reference library only exists where /root/reference was present at build time).
"""
import numpy as np
import pytest

from tests import sampler_cases as S


@pytest.fixture(scope="module")
def pair(oracle_lib, ref_lib):
    prob, ph = S.problem_tissue4((24, 24, 24), (0.8, 0.8, 0.8))
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    ref_lib.load_problem(prob)
    return oracle_lib, ref_lib, prob, ph


def test_drange_eloss_bit_exact(pair):
    orc, ref, prob, ph = pair
    for which, inp in ((S.DRANGE, S.drange_inputs(prob, 2000)), (S.ELOSS, S.eloss_inputs(prob, 3000))):
        a, b = orc.test_samplers(which, inp), ref.test_samplers(which, inp)
        assert np.array_equal(a, b)
        assert np.isfinite(a).all() and (a[:, 0] > 0).all()


def test_msdist_bit_exact(pair):
    orc, ref, prob, ph = pair
    for gi, grp in enumerate(S.MSDIST_GROUPS):
        for tilted in (False, True):
            inp = S.msdist_inputs(grp, 300, tilted)
            a, b = orc.test_samplers(S.MSDIST, inp, first=1000 * gi), ref.test_samplers(S.MSDIST, inp, first=1000 * gi)
            assert np.array_equal(a, b), grp
            assert np.allclose(np.linalg.norm(a[:, 4:7], axis=1), 1.0, atol=1e-9)


def test_sscat_compton_moller_bit_exact(pair):
    orc, ref, prob, ph = pair
    for gi, grp in enumerate(S.sscat_groups(prob)):
        inp = np.tile(np.asarray(grp, dtype=np.float64), (500, 1))
        assert np.array_equal(orc.test_samplers(S.SSCAT, inp, first=77 + gi), ref.test_samplers(S.SSCAT, inp, first=77 + gi))
    d = np.array([0.6, 0.0, 0.8])
    for e in S.COMPTON_ENERGIES:
        inp = np.tile(np.array([e, *d]), (500, 1))
        a, b = orc.test_samplers(S.COMPTON, inp, first=5), ref.test_samplers(S.COMPTON, inp, first=5)
        assert np.array_equal(a, b)
        assert np.allclose(a[:, 0] + a[:, 4] - S.RM, e, rtol=1e-12)
    for imed, ek in S.MOLLER_ENERGIES:
        inp = np.tile(np.array([imed, ek + S.RM, *d]), (500, 1))
        a, b = orc.test_samplers(S.MOLLER, inp, first=9), ref.test_samplers(S.MOLLER, inp, first=9)
        assert np.array_equal(a, b)
        assert np.allclose(a[:, 0] + a[:, 4], ek + 2 * S.RM, rtol=1e-12)


def test_photon_optical_depth_bit_exact(pair):
    orc, ref, prob, ph = pair
    rng = np.random.default_rng(3)
    n = 400
    pos = np.column_stack([rng.uniform(ph.xbounds[0] + 0.01, ph.xbounds[-1] - 0.01, n), rng.uniform(ph.ybounds[0] + 0.01, ph.ybounds[-1] - 0.01, n),
                           rng.uniform(ph.zbounds[0] + 0.01, ph.zbounds[-1] - 0.01, n)])
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    e = np.exp(rng.uniform(np.log(0.02), np.log(20.0), n))
    s = rng.uniform(0.0, 25.0, n)
    ta, ra = orc.test_photon_tau(e, np.column_stack([pos, d]), s)
    tb, rb = ref.test_photon_tau(e, np.column_stack([pos, d]), s)
    assert np.array_equal(ta, tb) and np.array_equal(ra, rb)
    assert (ta >= 0).all() and (ra == 0).any() and (ra > 0).any()

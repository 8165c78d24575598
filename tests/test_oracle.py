"""CPU tests (no GPU): the oracle restatement against the reference's golden runs and, where
oracle/_ref was built (needs /root/reference), against the unmodified reference itself.

The reference ships no golden vectors of its own; tests/golden/golden_*.npz were written by
running the reference with the per-history Philox stream (oracle/gen_fixtures.py), so equality with
them == equality with the reference.  gcc -O2 on both sides: results are required to be BIT-EXACT.
"""
import numpy as np
import pytest

from oracle.gen_fixtures import GOLDEN_RUNS, golden_problem
from ompmc_b200 import problem as P


def test_philox_known_answers(oracle_lib):
    """Philox4x32-10 known-answer vectors of Random123 (kat_vectors), through the history-keyed stream."""
    import ctypes as C
    # the stream helper fixes counter = (block, stream=0, hist_lo, hist_hi): check via a direct block call
    lib = C.CDLL(oracle_lib.path)
    # block function is static inline; verify through the stream API instead: ctr=(0,0,0,0), key=(0,0)
    oracle_lib._f("set_rng")(1, 0, 0)
    a = oracle_lib.test_rng(0, 4)
    want = np.array([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], dtype=np.float64) / 2.0 ** 32
    assert np.array_equal(a, want)
    oracle_lib._f("set_rng")(1, 0xa4093822 - (1 << 32), 0x299f31d0)
    # second block of a stream = counter word 0 incremented: only check determinism + range here
    b = oracle_lib.test_rng((0x03707344 << 32) | 0x13198a2e, 8)
    assert (b >= 0).all() and (b < 1).all() and len(np.unique(b)) == 8
    oracle_lib._f("set_rng")(1, 97, 33)


@pytest.mark.parametrize("name", list(GOLDEN_RUNS))
def test_oracle_reproduces_reference_golden(oracle_lib, name):
    prob, ph, cfg = golden_problem(name)
    z = np.load(P.golden(name + ".npz"))
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    rec = oracle_lib.run_histories(int(z["first"]), int(z["nhist"]), records=True)
    gold = z["records"]
    for f in ("ndraws", "ir_start", "ndeposit"):
        assert np.array_equal(rec[f], gold[f]), f
    assert np.array_equal(rec["edep"], gold["edep"])
    assert np.array_equal(oracle_lib.get_endep(), z["endep"])


def test_oracle_openmp_matches_serial(oracle_lib):
    """Philox streams are keyed by history id: thread count changes only the fp64 summation order."""
    prob, ph, cfg = golden_problem("golden_tissue4_6MV")
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    oracle_lib.set_num_threads(1)
    r1 = oracle_lib.run_histories(0, 1500, records=True)
    g1 = oracle_lib.get_endep()
    oracle_lib.reset_score()
    oracle_lib.set_num_threads(4)
    r4 = oracle_lib.run_histories(0, 1500, records=True)
    g4 = oracle_lib.get_endep()
    oracle_lib.set_num_threads(1)
    assert np.array_equal(r1, r4)
    np.testing.assert_allclose(g1, g4, rtol=1e-12, atol=1e-300)


def test_ranmar_bit_exact_vs_reference(oracle_lib, ref_lib):
    """RANMAR restatement (src/omc_random.c:58-187): integer state machine, must be bit-exact.  The
    reference side is exercised through a serial RANMAR transport run below."""
    a = oracle_lib.test_ranmar(97, 33, 1000)
    assert (a >= 0).all() and (a < 1).all()
    assert np.array_equal(a * 2 ** 24, np.round(a * 2 ** 24))       # 24-bit lattice
    # Marsaglia defaults for out-of-range seeds (:75-82)
    assert np.array_equal(oracle_lib.test_ranmar(0, 40000, 50), oracle_lib.test_ranmar(1802, 9373, 50))


@pytest.mark.parametrize("name", ["golden_water521_250kV", "golden_water700_6MV_ns5", "golden_tissue4_6MV"])
def test_oracle_vs_reference_ranmar_run(oracle_lib, ref_lib, name):
    """Serial run with the reference's own generator: whole energy grid bit-identical."""
    prob, ph, cfg = golden_problem(name)
    n = 1500
    ref_lib.load_problem(prob)
    ref_lib.set_rng("ranmar")
    ref_lib.reset_score()
    ref_lib.run_histories(0, n)
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("ranmar")
    oracle_lib.run_histories(0, n)
    assert np.array_equal(oracle_lib.get_endep(), ref_lib.get_endep())


def test_reference_blob_load_reproduces_golden(ref_lib):
    """ref_load_problem() (tables from the committed blob, no data files) == the reference's own init chain."""
    name = "golden_water700_e6MeV"
    prob, ph, cfg = golden_problem(name)
    z = np.load(P.golden(name + ".npz"))
    ref_lib.load_problem(prob)
    ref_lib.set_rng("philox")
    ref_lib.reset_score()
    rec = ref_lib.run_histories(int(z["first"]), int(z["nhist"]), records=True)
    assert np.array_equal(rec["ndraws"], z["records"]["ndraws"])
    assert np.array_equal(ref_lib.get_endep(), z["endep"])


def test_geometry_oracle_vs_reference(oracle_lib, ref_lib):
    prob, ph, _ = golden_problem("golden_tissue4_6MV")
    ref_lib.load_problem(prob)
    oracle_lib.load_problem(prob)
    rng = np.random.default_rng(3)
    n = 5000
    ir = rng.integers(0, ph.nreg, n).astype(np.int32)
    ix = (ir - 1) % ph.isize; iz = (ir - 1 - ix) // (ph.isize * ph.jsize); iy = ((ir - 1 - ix) - iz * ph.isize * ph.jsize) // ph.isize
    f = rng.random((n, 3)); f[:500] = np.round(f[:500])
    q = np.column_stack([ph.xbounds[ix] + f[:, 0] * np.diff(ph.xbounds)[ix], ph.ybounds[iy] + f[:, 1] * np.diff(ph.ybounds)[iy],
                         ph.zbounds[iz] + f[:, 2] * np.diff(ph.zbounds)[iz], rng.normal(size=(n, 3))])
    q[:, 3:] /= np.linalg.norm(q[:, 3:], axis=1, keepdims=True)
    us = rng.exponential(0.3, n)
    a = oracle_lib.test_geometry(q, ir, us)
    b = ref_lib.test_geometry(q, ir, us)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_work_counts(oracle_lib):
    prob, ph, cfg = golden_problem("golden_water700_6MV")
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    oracle_lib.reset_score()
    oracle_lib.run_histories(0, 500)
    w = oracle_lib.work_per_history()
    assert w["ausgab"] > 1 and w["hownear"] > 1 and w["pwlf"] > w["hownear"]

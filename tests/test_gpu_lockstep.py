"""GPU parity tests of the lock-step kernel, all through the C-ABI (ompmc_b200/api.py -> libompmc_b200.so).

Parity statement: relative to the reference run with the synthetic spinms.data (oracle/gen_fixtures.py).
The GPU consumes the same per-history Philox stream as the instrumented reference, so a history either
follows the reference draw for draw (same number of draws, same number of deposits, same start region,
deposited energy equal to rounding) or -- rarely -- diverges because CUDA's libm differs from glibc's in
the last ulp and flips a rejection test.  Integer bookkeeping is compared bit-exactly, energies to 1e-9.
"""
import numpy as np
import pytest

from oracle.gen_fixtures import GOLDEN_RUNS, golden_problem
from ompmc_b200 import problem as P

pytestmark = pytest.mark.gpu

# fraction of histories allowed to leave lock-step because of last-ulp libm differences
MAX_DIVERGED = 0.002
EDEP_RTOL = 1e-9


@pytest.mark.parametrize("name", list(GOLDEN_RUNS))
def test_lockstep_vs_reference_golden(gpu, name):
    prob, ph, cfg = golden_problem(name)
    z = np.load(P.golden(name + ".npz"))
    gold = z["records"]
    first, n = int(z["first"]), int(z["nhist"])
    gpu.load_problem(prob)
    gpu.set_option("kernel", 0)
    gpu.reset_tallies()
    rec = gpu.run_histories(first, n, records=True)
    grid = gpu.get_endep()
    assert (rec["flags"] & 1).sum() == 0, "stack overflow on device"
    # bit-exact integer bookkeeping: every history starts in the reference's voxel
    assert np.array_equal(rec["ir_start"], gold["ir_start"])
    same = (rec["ndraws"] == gold["ndraws"]) & (rec["ndeposit"] == gold["ndeposit"])
    diverged = (~same).mean()
    assert diverged <= MAX_DIVERGED, f"{(~same).sum()} of {n} histories left lock-step"
    rel = np.abs(rec["edep"][same] - gold["edep"][same]) / np.maximum(np.abs(gold["edep"][same]), 1e-30)
    assert rel.max() <= EDEP_RTOL
    # whole grid: identical up to the diverged histories
    tot_ref, tot = z["endep"].sum(), grid.sum()
    assert abs(tot - tot_ref) <= (1e-9 + 2.0 * diverged) * tot_ref
    if diverged == 0.0:
        np.testing.assert_allclose(grid, z["endep"], rtol=1e-9, atol=1e-12 * z["endep"].max())
    c = gpu.counters()
    assert c["histories"] == n and c["errors"] == 0
    assert c["rng_draws"] == int(rec["ndraws"].sum())
    assert c["deposits"] == int(rec["ndeposit"].sum())


def test_rng_matches_oracle(gpu, oracle_lib):
    prob, _, _ = golden_problem("golden_water521_250kV")
    gpu.load_problem(prob, seeds=(97, 33))
    oracle_lib.load_problem(prob, seeds=(97, 33))
    for hist in (0, 1, 12345, 2 ** 33 + 7):
        a = gpu.test_rng(hist, 67)
        b = oracle_lib.test_rng(hist, 67)
        assert np.array_equal(a, b)
        assert (a >= 0).all() and (a < 1).all()


def test_geometry_bit_exact(gpu, oracle_lib):
    """howfar()/hownear(): irnew, idisc bit-exact; ustep/tperp to 1 ulp (omc_dosxyz.c:187-334)."""
    prob, ph, _ = golden_problem("golden_tissue4_6MV")
    gpu.load_problem(prob)
    oracle_lib.load_problem(prob)
    rng = np.random.default_rng(7)
    n = 20000
    ir = rng.integers(0, ph.nreg, n).astype(np.int32)
    ix = (ir - 1) % ph.isize; iz = (ir - 1 - ix) // (ph.isize * ph.jsize); iy = ((ir - 1 - ix) - iz * ph.isize * ph.jsize) // ph.isize
    f = rng.random((n, 3))
    # a tenth of the particles sit exactly on a voxel face (ties between axes, zero distances)
    f[: n // 10] = np.round(f[: n // 10])
    x = ph.xbounds[ix] + f[:, 0] * (ph.xbounds[ix + 1] - ph.xbounds[ix])
    y = ph.ybounds[iy] + f[:, 1] * (ph.ybounds[iy + 1] - ph.ybounds[iy])
    zc = ph.zbounds[iz] + f[:, 2] * (ph.zbounds[iz + 1] - ph.zbounds[iz])
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[n // 10: n // 5, rng.integers(0, 3)] = 0.0        # axis-parallel directions
    q = np.column_stack([x, y, zc, d])
    ustep = rng.exponential(0.3, n)
    ustep[::7] = 1e8
    a = gpu.test_geometry(q, ir, ustep)
    b = oracle_lib.test_geometry(q, ir, ustep)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    np.testing.assert_allclose(a[2], b[2], rtol=4e-16, atol=0)
    np.testing.assert_allclose(a[3], b[3], rtol=4e-16, atol=0)


def test_batch_statistics_and_conservation(gpu, oracle_lib):
    """accumEndep() semantics (omc_dosxyz.c:696-717): accum = sum of batch grids, accum2 = sum of squares;
    energy conservation: deposited (incl. region 0 = escaped charged-particle energy) <= source energy."""
    prob, ph, cfg = golden_problem("golden_water700_6MV")
    gpu.load_problem(prob)
    gpu.reset_tallies()
    nb, per = 4, 500
    grids = []
    for ib in range(nb):
        gpu.run_histories(1000 + ib * per, per)
        grids.append(gpu.get_endep())
        gpu.accum_batch()
    a, a2, ensrc = gpu.get_tallies()
    g = np.array(grids)
    np.testing.assert_allclose(a, g.sum(0), rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(a2, (g * g).sum(0), rtol=1e-13, atol=1e-300)
    assert gpu.get_endep().sum() == 0.0
    assert 0.0 < a.sum() <= ensrc
    # same histories through the oracle: identical source energy (sum order differs -> 1e-12)
    oracle_lib.load_problem(prob)
    oracle_lib.set_rng("philox")
    oracle_lib.run_histories(1000, nb * per)
    _, _, ensrc_o = oracle_lib.get_accum()
    assert abs(ensrc - ensrc_o) <= 1e-12 * ensrc_o


def test_scheduling_independence(gpu):
    """History id -> RNG stream: splitting a range over launches / changing the grid changes nothing
    beyond fp64 summation order."""
    prob, ph, cfg = golden_problem("golden_tissue4_6MV")
    gpu.load_problem(prob)
    gpu.reset_tallies()
    r1 = gpu.run_histories(5000, 1200, records=True)
    g1 = gpu.get_endep()
    gpu.reset_tallies()
    gpu.set_option("max_blocks", 3)
    gpu.set_option("threads_per_block", 64)
    ra = gpu.run_histories(5000, 500, records=True)
    rb = gpu.run_histories(5500, 700, records=True)
    g2 = gpu.get_endep()
    gpu.set_option("max_blocks", 0)
    gpu.set_option("threads_per_block", 128)
    r2 = np.concatenate([ra, rb])
    assert np.array_equal(r1["ndraws"], r2["ndraws"]) and np.array_equal(r1["ndeposit"], r2["ndeposit"])
    assert np.array_equal(r1["edep"], r2["edep"])
    np.testing.assert_allclose(g1, g2, rtol=1e-12, atol=1e-300)


def test_error_paths(gpu):
    """Loud failures instead of the reference's printf+exit."""
    from ompmc_b200.api import OmcGpuError
    prob, _, _ = golden_problem("golden_water700_6MV_ns5")
    gpu.load_problem(prob)
    gpu.reset_tallies()
    gpu.set_option("stack_depth", 3)            # far too shallow for nsplit = 5
    gpu.run_histories(0, 256)
    with pytest.raises(OmcGpuError):
        gpu.synchronize()                       # "Stack overflow ... Increase MXSTACK!" (src/ompmc.c:1937-1940)
    gpu.set_option("stack_depth", 0)
    gpu.reset_tallies()
    with pytest.raises(OmcGpuError):
        gpu.set_option("no_such_option", 1)

"""GPU tests of the wavefront (production) kernels through the C-ABI.

The wavefront kernels use the same physics device functions as the lock-step kernel but other random
sub-streams (one per particle) and fp32 dose atomics, so agreement is STATISTICAL: per-voxel z-scores
from the batch method against the lock-step kernel, which itself is in lock-step with the
reference (tests/test_gpu_lockstep.py).  Integer bookkeeping (history count) and the source sampling
(primary particles use stream 0 of their history, like the lock-step kernel) are exact.
"""
import numpy as np
import pytest

from oracle.gen_fixtures import golden_problem
from ompmc_b200 import problem as P
from ompmc_b200.api import OmcGpuError

pytestmark = pytest.mark.gpu


def batches(gpu, kernel, nb, per, first=0):
    gpu.set_option("kernel", kernel)
    gpu.reset_tallies()
    for ib in range(nb):
        gpu.run_batch(first + ib * per, per)
    a, a2, ensrc = gpu.get_tallies()
    a, a2 = a[1:], a2[1:]                                        # region 0 = outside the phantom
    mean = a / nb
    var = np.maximum(a2 / nb - mean * mean, 0.0) / (nb - 1)      # variance of the batch mean
    return mean, var, ensrc, gpu.counters()


def zscore_check(m1, v1, m2, v2, frac_dmax=0.2):
    sel = (m1 > frac_dmax * m1.max()) & (v1 + v2 > 0)
    z = (m1[sel] - m2[sel]) / np.sqrt(v1[sel] + v2[sel])
    return z, sel


CASES = [
    ("water6mv", dict(mset="media_700_water.blob", ph=lambda: P.water_phantom("H2O700ICRU", (15, 15, 20), (1.0, 1.0, 1.0)),
                      spec="mohan6", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.7, charge=0, mono=0.0), 10, 200000),
    ("tissue6mv", dict(mset="media_700_tissue4.blob", ph=lambda: P.tissue_phantom((24, 10, 24), (0.8, 0.8, 0.8)),
                       spec="var_6MV", coll=(-4, 4, -3, 3), ssd=90.0, ecut=0.7, charge=0, mono=0.0), 10, 200000),
    ("water250kV", dict(mset="media_521_water.blob", ph=lambda: P.water_phantom("H2O521ICRU", (12, 12, 12), (1.0, 1.0, 1.0)),
                        spec="250", coll=(-3, 3, -3, 3), ssd=100.0, ecut=0.521, charge=0, mono=0.0), 10, 300000),
    ("e-6MeV", dict(mset="media_700_water.blob", ph=lambda: P.water_phantom("H2O700ICRU", (11, 11, 16), (0.8, 0.8, 0.25)),
                    spec=None, coll=(-2, 2, -2, 2), ssd=100.0, ecut=0.7, charge=-1, mono=6.0), 10, 20000),
    ("e+3MeV", dict(mset="media_521_water.blob", ph=lambda: P.water_phantom("H2O521ICRU", (11, 11, 16), (0.8, 0.8, 0.25)),
                    spec=None, coll=(-2, 2, -2, 2), ssd=100.0, ecut=0.521, charge=1, mono=3.0), 10, 20000),
    # uniform photon splitting + Russian roulette of secondary photons (the reference's shipped input uses nsplit = 20)
    ("water6mv-nsplit5", dict(mset="media_700_water.blob", ph=lambda: P.water_phantom("H2O700ICRU", (15, 15, 20), (1.0, 1.0, 1.0)),
                              spec="mohan6", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.7, charge=0, mono=0.0, nsplit=5), 10, 40000),
    ("tissue18mv-nsplit20", dict(mset="media_700_tissue4.blob", ph=lambda: P.tissue_phantom((24, 10, 24), (0.8, 0.8, 0.8)),
                                 spec=None, coll=(-4, 4, -3, 3), ssd=90.0, ecut=0.7, charge=0, mono=18.0, nsplit=20), 10, 5000),
]


def make_problem(cfg):
    media = P.load_blob(P.golden(cfg["mset"]))
    ph = cfg["ph"]()
    cdf = (media["cdfinv1_" + cfg["spec"]], media["cdfinv2_" + cfg["spec"]]) if cfg["spec"] else None
    prob = P.build_problem(media, ph, ecut=cfg["ecut"], pcut=0.01, collimator=cfg["coll"], ssd=cfg["ssd"], charge=cfg["charge"],
                           cdfinv=cdf, mono_energy=cfg["mono"], nsplit=cfg.get("nsplit", 1))
    return prob, ph


@pytest.mark.parametrize("name,cfg,nb,per", CASES, ids=[c[0] for c in CASES])
def test_wavefront_matches_lockstep_statistically(gpu, name, cfg, nb, per):
    prob, ph = make_problem(cfg)
    gpu.load_problem(prob)
    m0, v0, e0, c0 = batches(gpu, 0, nb, per)
    m1, v1, e1, c1 = batches(gpu, 1, nb, per)
    # exact bookkeeping
    assert c0["histories"] == c1["histories"] == nb * per
    assert c1["errors"] == 0
    assert abs(e0 - e1) <= 1e-9 * e0            # same primaries (stream 0 of every history)
    # energy balance: total deposited energy agrees within its statistical error
    tot0, tot1 = m0.sum(), m1.sum()
    s_tot = np.sqrt(v0.sum() + v1.sum()) * 3.0 + 1e-4 * tot0     # voxel sums are positively correlated: generous
    assert abs(tot0 - tot1) < max(s_tot, 0.004 * tot0), (tot0, tot1)
    # per-voxel: north_star criterion "voxels with dose > 20 % of Dmax agree within 2 sigma" (95.4 % expected)
    z, sel = zscore_check(m0, v0, m1, v1)
    assert sel.sum() >= 20
    within2 = (np.abs(z) < 2.0).mean()
    assert within2 >= 0.90, f"{within2:.3f} of {sel.sum()} voxels within 2 sigma"
    assert abs(z.mean()) < 0.35, f"systematic offset: mean z = {z.mean():.3f}"
    assert 0.6 < z.std() < 1.5, f"z spread {z.std():.3f}"


@pytest.mark.parametrize("idx", [0, 1], ids=["water6mv", "tissue6mv"])
def test_wavefront_voxel_march_matches_lockstep_statistically(gpu, idx):
    """photon_tracking = 0: the reference's voxel-to-voxel photon march instead of the default Woodcock flight
    (the parametrised test above runs the default, i.e. Woodcock for nsplit == 1 and the march for nsplit > 1)."""
    name, cfg, nb, per = CASES[idx]
    prob, ph = make_problem(cfg)
    gpu.load_problem(prob)
    m0, v0, e0, c0 = batches(gpu, 0, nb, per)
    gpu.set_option("photon_tracking", 0)
    try:
        m1, v1, e1, c1 = batches(gpu, 1, nb, per)
    finally:
        gpu.set_option("photon_tracking", 1)
    assert c0["histories"] == c1["histories"] == nb * per and c1["errors"] == 0
    assert abs(e0 - e1) <= 1e-9 * e0
    # The march crosses the same voxels as the lock-step kernel.  The reference (and the lock-step kernel) discover
    # that a photon has left on the NEXT howfar() call (irl == 0 -> idisc), one extra counted step per escaping
    # photon (~1 per history here; measured 0.566 = the escape probability for an uncollided 6 MeV primary);
    # the march discards at the crossing itself.
    extra = c0["photon_steps"] - c1["photon_steps"]
    assert 0 < extra < 2 * nb * per and extra < 0.06 * c0["photon_steps"]
    z, sel = zscore_check(m0, v0, m1, v1)
    assert (np.abs(z) < 2.0).mean() >= 0.90
    assert abs(z.mean()) < 0.35 and 0.6 < z.std() < 1.5


def test_woodcock_flight_on_nonuniform_grid(gpu):
    """Voxel look-up of the Woodcock flight by bisection when the planes are not equally spaced: same dose as
    the voxel march on a phantom whose z planes are graded."""
    name, cfg, nb, per = CASES[0]
    media = P.load_blob(P.golden(cfg["mset"]))
    ph = cfg["ph"]()
    z = np.asarray(ph.zbounds, dtype=np.float64)
    ph.zbounds = z[0] + (z - z[0]) * (0.6 + 0.4 * (z - z[0]) / (z[-1] - z[0]))      # graded, monotone
    prob = P.build_problem(media, ph, ecut=cfg["ecut"], pcut=0.01, collimator=cfg["coll"], ssd=cfg["ssd"], charge=0,
                           cdfinv=(media["cdfinv1_" + cfg["spec"]], media["cdfinv2_" + cfg["spec"]]), mono_energy=0.0)
    gpu.load_problem(prob)
    gpu.set_option("photon_tracking", 0)
    try:
        m0, v0, e0, c0 = batches(gpu, 1, nb, per)
    finally:
        gpu.set_option("photon_tracking", 1)
    m1, v1, e1, c1 = batches(gpu, 1, nb, per)
    assert c1["errors"] == 0 and c1["photon_steps"] < 0.5 * c0["photon_steps"]
    zs, sel = zscore_check(m0, v0, m1, v1)
    assert sel.sum() >= 20 and (np.abs(zs) < 2.0).mean() >= 0.90
    assert abs(zs.mean()) < 0.35 and 0.6 < zs.std() < 1.5


def test_wavefront_scheduling_independence(gpu):
    """Per-particle Philox sub-streams: pool size / crossings per wave / launch batching change only the fp32
    summation order.  (The drain kernel continues a particle's stream sequentially through its descendants,
    so WHERE the drain starts changes the draws: it is switched off for the exact comparison and checked
    statistically afterwards.)"""
    prob, ph = make_problem(CASES[1][1])
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    gpu.set_option("drain_threshold", 0)
    res = []
    for pool, cross, every, graph in ((1 << 23, 16, 16, 1), (1 << 14, 7, 3, 0), (1 << 16, 1000, 5, 1)):
        gpu.set_option("pool_size", pool); gpu.set_option("max_cross", cross); gpu.set_option("check_every", every)
        gpu.set_option("max_virtual", {16: 8, 7: 3, 1000: 64}[cross])      # Woodcock flight: tentative collisions per wave
        gpu.set_option("use_graph", graph)
        gpu.reset_tallies()
        gpu.run_histories(0, 50000)
        res.append((gpu.get_endep()[1:], gpu.counters()))
    gpu.set_option("pool_size", 1 << 23); gpu.set_option("max_cross", 16); gpu.set_option("check_every", 16); gpu.set_option("use_graph", 1)
    gpu.set_option("max_virtual", 8)
    g0, c0 = res[0]
    for g, c in res[1:]:
        assert c["deposits"] == c0["deposits"] and c["photon_steps"] == c0["photon_steps"]
        assert c["electron_steps"] == c0["electron_steps"]
        np.testing.assert_allclose(g, g0, rtol=2e-4, atol=1e-4 * g0.max())
        assert abs(g.sum() - g0.sum()) < 1e-5 * g0.sum()
    # with the drain: same configuration twice -> same bookkeeping; vs no drain -> statistically the same
    gpu.set_option("drain_threshold", 20000)
    runs = []
    for _ in range(2):
        gpu.reset_tallies()
        gpu.run_histories(0, 50000)
        runs.append((gpu.get_endep()[1:], gpu.counters()))
    gpu.set_option("drain_threshold", 8192)
    assert runs[0][1] == runs[1][1]
    np.testing.assert_allclose(runs[0][0], runs[1][0], rtol=2e-4, atol=1e-4 * g0.max())
    assert abs(runs[0][0].sum() - g0.sum()) < 0.01 * g0.sum()
    assert runs[0][1]["histories"] == c0["histories"]


def test_batch_pipelining_matches_serial_batches(gpu):
    """omc_gpu_run_batch() pipelines consecutive batches (the next batch is injected while the tail of the previous one
    is still in flight, each particle scoring into the grid of the batch its history id belongs to).  Per-particle RNG
    streams make every history identical to the serial run, so accum AND accum2 (the per-batch squares: a particle
    scored into the wrong batch would show there) agree to fp32 summation order.  Drain off on both sides."""
    prob, ph = make_problem(CASES[1][1])
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    gpu.set_option("drain_threshold", 0)
    nb, per = 6, 30000
    try:
        gpu.reset_tallies()
        for ib in range(nb):                                   # serial: each batch runs to the end, then accumEndep()
            gpu.run_histories(ib * per, per)
            gpu.accum_batch()
        a0, b0, e0 = gpu.get_tallies()
        c0 = gpu.counters()
        gpu.set_option("pool_size", 1 << 15)                   # small pool: injection of a batch spans many waves
        gpu.reset_tallies()
        for ib in range(nb):
            gpu.run_batch(ib * per, per)
        a1, b1, e1 = gpu.get_tallies()
        c1 = gpu.counters()
        # explicit form: start_batch / completed_batches / accum_batch / finish_batches
        gpu.reset_tallies()
        for ib in range(nb):
            gpu.start_batch(ib * per, per)
            assert gpu.completed_batches() == (1 if ib > 0 else 0)
            while gpu.completed_batches():
                gpu.accum_batch()
        gpu.finish_batches()
        assert gpu.completed_batches() == 1
        gpu.accum_batch()
        a2, b2, e2 = gpu.get_tallies()
        # two batches waiting -> a third start is refused
        gpu.reset_tallies()
        gpu.start_batch(0, 1000); gpu.start_batch(1000, 1000); gpu.finish_batches()
        assert gpu.completed_batches() == 2
        with pytest.raises(OmcGpuError):
            gpu.start_batch(2000, 1000)
    finally:
        gpu.set_option("pool_size", 1 << 23); gpu.set_option("drain_threshold", 8192)
        gpu.reset_tallies()
    for a, b, e in ((a1, b1, e1), (a2, b2, e2)):
        assert abs(e - e0) <= 1e-9 * e0
        np.testing.assert_allclose(a[1:], a0[1:], rtol=3e-4, atol=1e-4 * a0.max())
        np.testing.assert_allclose(b[1:], b0[1:], rtol=6e-4, atol=1e-4 * b0.max())
    assert c1["histories"] == c0["histories"] == nb * per
    assert c1["deposits"] == c0["deposits"] and c1["electron_steps"] == c0["electron_steps"]


def test_wavefront_rejects_unsupported(gpu):
    prob, ph = make_problem(dict(CASES[0][1], nsplit=300))
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    with pytest.raises(OmcGpuError):
        gpu.run_histories(0, 100)          # nsplit > 255: lock-step kernel only
    with pytest.raises(OmcGpuError):
        gpu.run_histories(0, 100, records=True)   # per-history records are a lock-step feature
    gpu.set_option("kernel", 0)


def test_wavefront_queue_overflow_is_loud(gpu):
    prob, ph = make_problem(CASES[0][1])
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    gpu.set_option("pool_cap", 1000)                # queues far too small for the 4M-particle target
    gpu.reset_tallies()
    with pytest.raises(OmcGpuError):
        gpu.run_histories(0, 100000)
        gpu.synchronize()
    gpu.set_option("pool_cap", 0)
    gpu.reset_tallies()
    gpu.run_histories(0, 1000)
    gpu.synchronize()

"""CPU tests of the host-side logic that mirrors the reference user code above the batch loop."""
import os

import numpy as np
import pytest

from ompmc_b200 import problem as P


def test_atoi_batch_plan_matches_reference_bookkeeping():
    """omc_dosxyz.c:1207-1225: atoi() ints, truncating division, nhist < nbatch -> nbatch."""
    assert P.batch_plan("100000", "10") == (100000, 10, 10000)
    assert P.batch_plan("1e9", "10") == (10, 10, 1)                 # atoi("1e9") == 1
    assert P.batch_plan("1005", "10") == (1000, 10, 100)
    assert P.batch_plan("  42abc", "5") == (40, 5, 8)
    assert P.batch_plan("3", "10") == (10, 10, 1)
    assert P.c_atoi("-17x") == -17 and P.c_atoi("x") == 0


def test_blob_roundtrip(tmp_path):
    a = {"x": np.arange(7, dtype=np.float64), "i": np.array([3, -1, 2], dtype=np.int32), "s": np.array([2.5])}
    p = tmp_path / "t.blob"
    P.save_blob(str(p), a)
    b = P.load_blob(str(p))
    assert list(b) == list(a)
    for k in a:
        assert np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype


def test_egsphant_roundtrip(tmp_path):
    ph = P.tissue_phantom((9, 5, 7), (0.3, 0.4, 0.25))
    p = tmp_path / "t.egsphant"
    P.write_egsphant(str(p), ph)
    q = P.read_egsphant(str(p))
    assert q.media == ph.media
    for f in ("xbounds", "ybounds", "zbounds", "med_indices", "med_densities"):
        assert np.array_equal(getattr(q, f), getattr(ph, f)), f
    assert set(np.unique(ph.med_indices)) <= {1, 2, 3, 4}


def test_regions_follow_reference_rules():
    media = P.load_blob(P.golden("media_700_tissue4.blob"))
    ph = P.tissue_phantom((8, 4, 8), (0.5, 0.5, 0.5))
    r = P.init_regions(ph, media, ecut=0.7, pcut=0.01)
    assert r["region_med"][0] == -1 and r["region_rhof"][0] == 0.0
    imed = ph.med_indices - 1
    np.testing.assert_array_equal(r["region_med"][1:], imed)
    np.testing.assert_array_equal(r["region_rhof"][1:], ph.med_densities / media["pegs_rho"][imed])
    assert (r["region_pcut"][1:] >= media["pegs_ap"][imed]).all()
    with pytest.raises(ValueError):
        P.init_regions(ph, media, ecut=0.6, pcut=0.01)              # below AE = 0.7: reference leaves ecut uninitialised


def test_spectrum_tables_match_reference_dump():
    """python initSource() restatement vs the cdfinv tables dumped from the reference (in the media blob)."""
    ref = "/root/reference/spectra"
    if not os.path.isdir(ref):
        pytest.skip("spectra live in /root/reference only")
    media = P.load_blob(P.golden("media_700_water.blob"))
    for key, fn in (("250", "250.spectrum"), ("mohan6", "mohan6.spectrum"), ("var_6MV", "var_6MV.spectrum")):
        c1, c2 = P.spectrum_cdfinv(open(os.path.join(ref, fn)).read())
        assert np.array_equal(c1, media["cdfinv1_" + key]) and np.array_equal(c2, media["cdfinv2_" + key])


def test_accumulate_results_and_3ddose(tmp_path, ref_lib):
    """Batch statistics + dose conversion vs the reference's accumulateResults() (omc_dosxyz.c:719-799)."""
    from oracle.gen_fixtures import golden_problem
    prob, ph, cfg = golden_problem("golden_tissue4_6MV")
    ref_lib.load_problem(prob)
    ref_lib.set_rng("philox")
    ref_lib.reset_score()
    nb, per = 5, 300
    for ib in range(nb):
        ref_lib.run_histories(ib * per, per)
        ref_lib.accum_endep()
    a, a2, _ = ref_lib.get_accum()
    dose, unc = P.accumulate_results(ph, a, a2, per, nb)
    fn = ref_lib._f("accumulate_results"); fn.argtypes = [__import__("ctypes").c_int] * 3
    fn(1, per, nb)
    rd, ru, _ = ref_lib.get_accum()
    np.testing.assert_allclose(dose, rd[1:], rtol=1e-14, atol=0)
    np.testing.assert_allclose(unc, ru[1:], rtol=1e-12, atol=0)
    assert (dose[ph.med_densities < 0.044] == 0).all() and (unc[ph.med_densities < 0.044] == 0.9999999).all()
    p = tmp_path / "o.3ddose"
    P.write_3ddose(str(p), ph, dose, unc)
    dims, bounds, d2, u2 = P.read_3ddose(str(p))
    assert dims == (ph.isize, ph.jsize, ph.ksize)
    np.testing.assert_allclose(d2, dose, rtol=1e-6)
    np.testing.assert_allclose(u2, unc, atol=1e-6)


def test_accumulate_results_without_reference():
    ph = P.water_phantom("H2O700ICRU", (2, 2, 2), (1.0, 1.0, 1.0))
    rng = np.random.default_rng(0)
    g = rng.random((4, ph.nreg))
    g[:, 3] = 0.0
    dose, unc = P.accumulate_results(ph, g.sum(0), (g * g).sum(0), 100, 4)
    e = g[:, 1:].mean(0)
    s = np.sqrt(((g[:, 1:] ** 2).mean(0) - e * e) / 3) / np.where(e > 0, e, 1)
    ok = e > 0
    np.testing.assert_allclose(unc[ok], s[ok], rtol=1e-12)
    assert dose[2] == 0.0 and unc[2] == 0.9999999
    np.testing.assert_allclose(dose[ok], e[ok] * 1.602e-10 / (1.0 * 100), rtol=1e-14)


def test_resample_phantom_keeps_material_map():
    ph = P.tissue_phantom((6, 4, 5), (0.3, 0.3, 0.3))
    q = P.resample_phantom(ph, (3, 3, 3))
    assert (q.isize, q.jsize, q.ksize) == (18, 12, 15)
    np.testing.assert_allclose(np.diff(q.xbounds), 0.1, rtol=1e-9)
    m = q.med_indices.reshape(q.ksize, q.jsize, q.isize)[::3, ::3, ::3]
    assert np.array_equal(m.reshape(-1), ph.med_indices)
    # mass is conserved
    vol = lambda p: (np.diff(p.zbounds)[:, None, None] * np.diff(p.ybounds)[None, :, None] * np.diff(p.xbounds)[None, None, :]).reshape(-1)
    assert abs((vol(q) * q.med_densities).sum() - (vol(ph) * ph.med_densities).sum()) < 1e-9


def test_c_host_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    """ompmc_b200/host/omc_dosxyz_b200.c: the plain-C batch loop + statistics + .3ddose writer on the C-ABI."""
    import subprocess
    import torch
    from ompmc_b200 import build
    from oracle.gen_fixtures import golden_problem
    build.build()
    assert os.path.exists(build.HOST_EXE)
    r = subprocess.run([build.HOST_EXE, "--help"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stdout
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_host.py")
    prob, ph, cfg = golden_problem("golden_water700_6MV")
    blob = str(tmp_path / "p.blob")
    P.save_blob(blob, prob)
    r = subprocess.run([build.HOST_EXE, "-p", blob, "-n", "1000", "-b", "4", "-o", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU transport" in r.stdout
    assert not os.path.exists(str(tmp_path / "o.3ddose"))

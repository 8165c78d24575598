"""Host-side initialisation restated in C (ompmc_b200/host/omc_tables.c, omc_host_input.h) against the
reference itself: the physics tables built from the raw PEGS4 / XCOM / form-factor / msnew / spinms files must equal, bit
for bit, the tables the reference's initMediaData() produced (tests/golden/media_*.blob, dumped from the compiled reference
by oracle/gen_fixtures.py), and an input-file run of the C driver must hand the GPU library exactly the arrays the
reference holds in its globals before the batch loop (the golden problems).  CPU only; needs the reference's DATA files
(/root/reference/{data,pegs4,spectra}, or their copy staged under the git-ignored oracle/_ref/refdata by `make -C oracle refdata`)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import gen_fixtures as G
from ompmc_b200 import api, build, problem as P

pytestmark = pytest.mark.skipif(not G.have_data(), reason="reference data files not present")


@pytest.fixture(scope="module")
def tables_lib():
    build.build_host()
    lib = C.CDLL(build.TABLES_LIB)
    lib.omc_tables_build.restype = C.c_void_p
    lib.omc_tables_build.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int]
    lib.omc_tables_view.restype = C.POINTER(api.MediaTables)
    lib.omc_tables_view.argtypes = [C.c_void_p]
    lib.omc_tables_free.argtypes = [C.c_void_p]
    lib.omc_spectrum_cdfinv.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_char_p, C.c_int]
    return lib


@pytest.fixture(scope="module")
def workdir():
    return G.prepare_workdir()          # data folder = the reference's files + the synthetic spinms.data of the fixtures


def build_tables(lib, work, cfg):
    names = (C.c_char_p * len(cfg["media"]))(*[m.encode() for m in cfg["media"]])
    err = C.create_string_buffer(512)
    h = lib.omc_tables_build((work + "/data/").encode(), (G.DATA + "/pegs4/" + cfg["pegs"]).encode(), (G.DATA + "/pegs4/pgs4form.dat").encode(),
                             len(cfg["media"]), names, err, 512)
    assert h, err.value.decode()
    return h


@pytest.mark.parametrize("name", list(G.MEDIA_SETS))
def test_tables_equal_the_references_bit_for_bit(tables_lib, workdir, name):
    cfg = G.MEDIA_SETS[name]
    h = build_tables(tables_lib, workdir, cfg)
    v = tables_lib.omc_tables_view(h).contents
    blob = P.load_blob(P.golden(name + ".blob"))
    mine = {"nmed": np.array([v.nmed], dtype=np.int32)}
    for fname, typ in api.MediaTables._fields_[1:]:
        val = getattr(v, fname)
        mine[fname] = np.ctypeslib.as_array(val, shape=(blob[fname].size,)).copy() if typ in (api.PD, api.PI) else np.array([val])
    G.sanitize(mine)                                # same treatment of never-initialised / never-read entries as the fixtures
    assert int(mine["nmed"][0]) == len(cfg["media"])
    for k, arr in mine.items():
        assert arr.shape == blob[k].shape, k
        assert np.array_equal(arr, blob[k]), f"{name}: table '{k}' differs in {(arr != blob[k]).sum()} of {arr.size} entries"
    tables_lib.omc_tables_free(h)


def test_missing_medium_and_missing_file_are_reported(tables_lib, workdir):
    names = (C.c_char_p * 1)(b"NOSUCHMEDIUM")
    err = C.create_string_buffer(512)
    args = ((workdir + "/data/").encode(), (G.DATA + "/pegs4/700icru.pegs4dat").encode(), (G.DATA + "/pegs4/pgs4form.dat").encode())
    assert not tables_lib.omc_tables_build(*args, 1, names, err, 512) and b"NOSUCHMEDIUM" in err.value
    names = (C.c_char_p * 1)(b"H2O700ICRU")
    assert not tables_lib.omc_tables_build(b"/nonexistent/", args[1], args[2], 1, names, err, 512) and b"Unable to open" in err.value


@pytest.mark.parametrize("key,fname", list(G.SPECTRA.items()))
def test_spectrum_inverse_cdf(tables_lib, workdir, key, fname):
    blob = P.load_blob(P.golden("media_700_water.blob"))
    c1 = np.zeros(1000); c2 = np.zeros(1000); emax = C.c_double(0)
    err = C.create_string_buffer(256)
    assert tables_lib.omc_spectrum_cdfinv((G.DATA + "/spectra/" + fname).encode(), c1.ctypes.data, c2.ctypes.data, C.byref(emax), err, 256) == 0
    assert np.array_equal(c1, blob["cdfinv1_" + key]) and np.array_equal(c2, blob["cdfinv2_" + key])
    assert emax.value >= (c1 + c2).max() > 0.0


@pytest.mark.parametrize("name", list(G.GOLDEN_RUNS))
def test_input_file_run_builds_the_references_problem(workdir, name):
    """`omc_dosxyz_b200 -i <stem> --dump-problem`: the reference's own .inp / .egsphant / data files in, every array that would
    go to the GPU library out (no GPU needed), equal to the golden problem = the reference's globals before its batch loop."""
    cfg = G.GOLDEN_RUNS[name]
    ph = cfg["ph"]()
    ppath = os.path.join(workdir, name + "_c.egsphant")
    P.write_egsphant(ppath, ph)
    stem = os.path.join(workdir, name + "_c")
    mcfg = G.MEDIA_SETS[cfg["mset"]]
    G.write_inp(stem, phantom=ppath, pegs=mcfg["pegs"], spectrum=G.SPECTRA[cfg["spectrum"]] if cfg["spectrum"] else None, mono=cfg["mono"],
                charge=cfg["charge"], coll=cfg["coll"], ssd=cfg["ssd"], ecut=mcfg["ecut"], pcut=0.01, nsplit=cfg["nsplit"])
    build.build()
    r = subprocess.run([build.HOST_EXE, "-i", stem, "-o", stem, "--dump-problem"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    mine = P.load_blob(stem + ".problem")
    G.sanitize(mine)
    ref, _, _ = G.golden_problem(name)
    extras = {"med_indices", "src_ixinu", "src_iyinu"}          # python-side helper / garbage-dependent and only ever printed
    checked = 0
    for k, v in ref.items():
        if k in extras or k.startswith(("cdfinv1_", "cdfinv2_")):
            continue
        if k in ("src_cdfinv1", "src_cdfinv2") and int(ref["src_spectrum"][0]) == 0:
            continue
        assert k in mine, f"{name}: '{k}' missing from the C driver's problem"
        assert np.array_equal(np.asarray(mine[k]), np.asarray(v)), f"{name}: '{k}' differs"
        checked += 1
    assert checked > 100


def test_input_parser_follows_the_reference(tmp_path):
    """Q16: a '#' anywhere drops the line; keys match by substring and the first match wins; blanks inside path values go."""
    stem = str(tmp_path / "x")
    with open(stem + ".inp", "w") as f:
        f.write("\n".join(["# comment", "ncase = 1000 # a trailing comment drops the whole line", "ncase  = 77", "", "my nbatch key = 4",
                           "nbatch = 9", "phantom file = /no such / file .egsphant", ""]))
    r = subprocess.run([build.HOST_EXE, "-i", stem, "-o", stem, "--dump-problem"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unable to open file: /nosuch/file.egsphant" in r.stdout


@pytest.mark.parametrize("voxel", [(0.2, 0.2, 0.2), (0.1, 0.1, 0.1), (0.45, 0.3, 0.25)])
def test_phantom_resampler_any_ratio(workdir, voxel):
    """BASELINE config 5: `omc_dosxyz_b200 -i <stem> -v "dx dy dz" --dump-problem` resamples the .egsphant to any
    voxel size -- 3 mm -> 2 mm is not an integer split -- keeping the extent, conserving mass exactly (volume-weighted densities),
    medium by largest overlap; identical to the numpy mirror problem.resample_phantom_to(); a 3 -> 1 mm split copies the map."""
    name = "golden_tissue4_6MV"
    cfg = G.GOLDEN_RUNS[name]
    ph = P.tissue_phantom((21, 19, 11), (0.3, 0.3, 0.3), "prostate")
    ppath = os.path.join(workdir, "resample_src.egsphant")
    P.write_egsphant(ppath, ph)
    ph = P.read_egsphant(ppath)                      # (the densities as the file's text holds them)
    stem = os.path.join(workdir, "resample_%g" % voxel[0])
    mcfg = G.MEDIA_SETS[cfg["mset"]]
    G.write_inp(stem, phantom=ppath, pegs=mcfg["pegs"], spectrum=G.SPECTRA[cfg["spectrum"]], mono=cfg["mono"], charge=cfg["charge"],
                coll=cfg["coll"], ssd=cfg["ssd"], ecut=mcfg["ecut"], pcut=0.01, nsplit=1)
    build.build()
    r = subprocess.run([build.HOST_EXE, "-i", stem, "-o", stem, "-v", "%g %g %g" % voxel, "--dump-problem"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    mine = P.load_blob(stem + ".problem")
    want = P.resample_phantom_to(ph, voxel)
    assert (int(mine["isize"][0]), int(mine["jsize"][0]), int(mine["ksize"][0])) == (want.isize, want.jsize, want.ksize)
    np.testing.assert_allclose(mine["xbounds"], want.xbounds, rtol=0, atol=1e-12)
    np.testing.assert_allclose(mine["zbounds"], want.zbounds, rtol=0, atol=1e-12)
    assert np.array_equal(mine["region_med"][1:], want.med_indices - 1)
    np.testing.assert_allclose(mine["med_densities"], want.med_densities, rtol=1e-9)
    # mass conservation
    def mass(p):
        vol = np.multiply.outer(np.diff(p.zbounds), np.multiply.outer(np.diff(p.ybounds), np.diff(p.xbounds))).reshape(-1)
        return float((vol * p.med_densities).sum())
    assert abs(mass(want) - mass(ph)) < 1e-10 * mass(ph)
    if voxel == (0.1, 0.1, 0.1):                     # integer split: the material map is copied voxel for voxel
        split = P.resample_phantom(ph, (3, 3, 3))
        assert np.array_equal(split.med_indices, want.med_indices)
        np.testing.assert_allclose(split.med_densities, want.med_densities, rtol=1e-9)

"""The reference's OWN matRad user code (ucodes/omc_matrad/omc_matrad.c) with its beamlet loop on libompmc_b200.so:
ompmc_b200/host/omc_matrad_dropin.c = the mexFunction() a maintainer builds, compiled here against the mex.h stand-in of
oracle/mexshim (no MATLAB in this image) with a harness main() that builds the five MATLAB inputs (cubeRho, cubeMatIx, ompMCgeo,
ompMCsource, ompMCoptions) from a problem blob -> oracle/_ref/omc_matrad_dropin (git-ignored, built by oracle/Makefile).

CPU: the reference's parseInput() / initPhantom() / initMediaData() / initSource() / initRegions() run on those inputs must leave
in the reference's globals exactly the arrays this repository's host builds for the same plan (tables, regions, spectrum inverse
CDF, beamlet geometry: bit for bit), and without a CUDA device the mex stops through mexErrMsgIdAndTxt -- no CPU fallback.
GPU: its sparse matrix against the plain-C driver's (omc_matrad_b200) on the same blob.  (File named to run last.)"""
import os
import subprocess

import numpy as np
import pytest

from oracle import gen_fixtures as G
from ompmc_b200 import build, problem as P

DROPIN = os.path.join(G.HERE, "_ref", "omc_matrad_dropin")
pytestmark = pytest.mark.skipif(not (os.path.exists(DROPIN) and G.have_data()),
                                reason="oracle/_ref/omc_matrad_dropin or the reference data files not present")


def dropin_cmd(nhist="60003", nbatch="3", rel="0.02"):
    from tests.test_matrad import matrad_problem
    prob, ph, nb = matrad_problem(nbix=(2, 1), angles=(0.0, 100.0, 200.0))
    work = G.prepare_workdir()
    blob = os.path.join(work, "mdrop.blob")
    P.save_blob(blob, prob)
    stem = os.path.join(work, "mdrop")
    cmd = [DROPIN, "-p", blob, "-m", ",".join(ph.media), "--data", work + "/data/", "--pegs", G.DATA + "/pegs4/700icru.pegs4dat",
           "--pgs4form", G.DATA + "/pegs4/pgs4form.dat", "--spectrum", G.DATA + "/spectra/var_6MV.spectrum", "-n", nhist, "-b", nbatch,
           "-t", rel, "--ecut", "0.7", "--pcut", "0.01", "-o", stem]
    return cmd, prob, ph, nb, blob, stem


def test_reference_matrad_init_path_equals_this_hosts_problem():
    cmd, prob, ph, nb, blob, stem = dropin_cmd()
    r = subprocess.run(cmd + ["--dump-problem"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    assert f"Total Number of Beamlets:{nb}" in r.stdout and "Number of media in phantom : 4" in r.stdout
    mine = P.load_blob(stem + ".problem")
    G.sanitize(mine)
    beamlets = P.load_blob(stem + ".beamlets")
    helpers = {"med_indices"}                                  # python-side entry (the reference keeps the cube itself)
    checked = 0
    for k, v in prob.items():
        if k in helpers or k.startswith(("cdfinv1_", "cdfinv2_")):
            continue
        src = beamlets if k.startswith("mr_") else mine
        assert k in src, f"'{k}' missing from the reference's initialised problem"
        assert np.array_equal(np.asarray(src[k]), np.asarray(v)), f"'{k}' differs"
        checked += 1
    assert checked > 110
    assert int(beamlets["mr_nbeams"][0]) == 3


def test_matrad_dropin_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cmd, prob, ph, nb, blob, stem = dropin_cmd()
    if os.path.exists(stem + ".csc"):
        os.remove(stem + ".csc")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode != 0 and "Histories per batch: 20001" in r.stdout
    assert "matRad:matRad_ompInterface:invalid" in r.stderr and "omc_gpu_create" in r.stderr
    assert not os.path.exists(stem + ".csc")


def _read_csc(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"OMCCSC1\0"
        nrows, ncols, nnz = np.fromfile(f, dtype="<i8", count=3)
        jc = np.fromfile(f, dtype="<i8", count=ncols + 1)
        ir = np.fromfile(f, dtype="<i8", count=nnz)
        pr = np.fromfile(f, dtype="<f8", count=nnz)
    return int(nrows), int(ncols), jc, ir, pr


@pytest.mark.gpu
def test_reference_matrad_user_code_on_the_gpu_library(gpu):
    cmd, prob, ph, nb, blob, stem = dropin_cmd()
    build.build()
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    assert "Total number of particle histories: 60003" in r.stdout and "Sparse MC Dij has" in r.stdout
    own = stem + "_own"
    r2 = subprocess.run([build.MATRAD_EXE, "-p", blob, "-n", "60003", "-b", "3", "-t", "0.02", "-o", own], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    a, b = _read_csc(stem + ".csc"), _read_csc(own + ".csc")
    assert a[0] == b[0] == ph.nvox and a[1] == b[1] == nb
    for k in range(nb):
        d0 = np.zeros(ph.nvox); d1 = np.zeros(ph.nvox)
        rows = a[3][a[2][k]:a[2][k + 1]]
        assert np.all(np.diff(rows) > 0)                        # rows ascending as the reference's irl loop writes them
        d0[rows] = a[4][a[2][k]:a[2][k + 1]]
        d1[b[3][b[2][k]:b[2][k + 1]]] = b[4][b[2][k]:b[2][k + 1]]
        assert d0.max() > 0
        np.testing.assert_allclose(d0, d1, rtol=2e-3, atol=0.021 * d1.max())       # (entries next to the threshold may flip)

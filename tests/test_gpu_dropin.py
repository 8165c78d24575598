"""The reference's OWN omc_dosxyz user code with its batch loop on libompmc_b200.so (ompmc_b200/host/omc_dosxyz_dropin.c:
ucodes/omc_dosxyz/omc_dosxyz.c #included from the reference checkout, compiled by oracle/Makefile into the git-ignored
oracle/_ref/omc_dosxyz_dropin together with src/ompmc.c) -- the patch of INTEGRATION.md, run for real.

parseInputFile / initPhantom / initMediaData / initSource / initRegions / accumulateResults / the .3ddose writer are the
reference's; only {initHistory(); shower();} x nperbatch + accumEndep() is the library's.  Checked against this repository's
plain-C host (omc_dosxyz_b200 -i, whose tables are the restatement of ompmc_b200/host/omc_tables.c) on the same input file:
both hand the device the same problem bit for bit (tests/test_tables.py), so the two .3ddose files agree to the summation
order of the dose atomics, and the header lines byte for byte."""
import os
import subprocess

import numpy as np
import pytest

from oracle import gen_fixtures as G
from ompmc_b200 import build, problem as P

DROPIN = os.path.join(G.HERE, "_ref", "omc_dosxyz_dropin")
needs_dropin = pytest.mark.skipif(not (os.path.exists(DROPIN) and G.have_data()),
                                  reason="oracle/_ref/omc_dosxyz_dropin or the staged reference data files not present")


def write_case(name, ncase, nbatch):
    work = G.prepare_workdir()
    cfg = G.GOLDEN_RUNS[name]
    ph = cfg["ph"]()
    ppath = os.path.join(work, name + "_d.egsphant")
    P.write_egsphant(ppath, ph)
    stem = os.path.join(work, name + "_d")
    m = G.MEDIA_SETS[cfg["mset"]]
    G.write_inp(stem, phantom=ppath, pegs=m["pegs"], spectrum=G.SPECTRA[cfg["spectrum"]] if cfg["spectrum"] else None, mono=cfg["mono"],
                charge=cfg["charge"], coll=cfg["coll"], ssd=cfg["ssd"], ecut=m["ecut"], pcut=0.01, nsplit=cfg["nsplit"], ncase=ncase, nbatch=nbatch)
    return work, stem, ph


@needs_dropin
def test_dropin_fails_loudly_without_a_device():
    """No CPU fallback behind the reference's host code either: the reference's error behaviour (message + EXIT_FAILURE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    work, stem, ph = write_case("golden_water700_6MV", 1000, 10)
    r = subprocess.run([DROPIN, "-i", stem, "-o", "dropin_nogpu"], capture_output=True, text=True)
    assert r.returncode != 0 and "Histories per batch: 100" in r.stdout and "omc_gpu_multi_create" in r.stdout
    assert not os.path.exists(os.path.join(work, "dropin_nogpu.3ddose"))


@pytest.mark.gpu
@needs_dropin
@pytest.mark.parametrize("name,kernel,ncase", [("golden_tissue4_6MV", 0, 40005), ("golden_tissue4_6MV", 1, 400005), ("golden_water521_250kV", 1, 200000)])
def test_reference_user_code_on_the_gpu_library(gpu, name, kernel, ncase):
    work, stem, ph = write_case(name, ncase, 8)
    build.build()
    r = subprocess.run([DROPIN, "-i", stem, "-o", "dropin_out", "-k", str(kernel)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    nper = ncase // 8
    assert f"Total number of particle histories: {nper * 8}" in r.stdout and f"Histories per batch: {nper}" in r.stdout
    mine = os.path.join(work, "own_out")
    r2 = subprocess.run([build.HOST_EXE, "-i", stem, "-o", mine, "-k", str(kernel)], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-500:]
    fa, fb = os.path.join(work, "dropin_out.3ddose"), mine + ".3ddose"       # the reference writes into its "output folder"
    dims, bounds, dose, unc = P.read_3ddose(fa)
    dims2, bounds2, dose2, unc2 = P.read_3ddose(fb)
    assert dims == dims2 == (ph.isize, ph.jsize, ph.ksize)
    with open(fa) as a, open(fb) as b:                                        # dimensions + the three boundary lines
        for _ in range(4):
            assert a.readline() == b.readline()
    assert dose.max() > 0
    tol = 2e-6 if kernel == 0 else 5e-4          # %e keeps 7 digits; fp32 atomics order differs between runs
    np.testing.assert_allclose(dose, dose2, rtol=tol, atol=tol * dose2.max())
    np.testing.assert_allclose(unc, unc2, atol=2e-6 if kernel == 0 else 2e-3)

"""Scaled-down BASELINE configurations for the statistical parity fixtures.  TEST INFRASTRUCTURE (oracle/).

Each case is a BASELINE.json configuration with the SAME physics (media set, spectrum, field, SSD, cut-offs,
physical extent of the phantom) on a coarser voxel grid, so that the unmodified reference reaches a per-voxel
noise of ~0.3 % in the high-dose region in minutes of host time and gamma(1 %/1 mm) can be evaluated at the
grid's NATIVE resolution.  ``oracle/gen_parity_fixtures.py`` runs the reference (oracle/_ref, OpenMP, its own
RANMAR generator) on them and commits the batch statistics under tests/golden/parity_*.npz;
tests/test_gpu_parity.py runs the production CUDA kernels on the same problems against those files.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ompmc_b200 import problem as P  # noqa: E402

# nhist = reference histories (nbatch batches); the GPU side runs gpu_mult times as many
CASES = {
    # BASELINE config 1: WATER, 250 kV spectrum, 521icru (the reference's own CPU-runnable case)
    "water250kv": dict(media="media_521_water.blob", phantom=lambda: P.water_phantom("H2O521ICRU", (31, 31, 30), (1.0, 1.0, 1.0)),
                       spectrum="250", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.521, nsplit=1, nhist=300_000_000, nbatch=40),
    # BASELINE config 2: WATER, mohan6 6 MV 10x10 cm2, 700icru
    "water6mv": dict(media="media_700_water.blob", phantom=lambda: P.water_phantom("H2O700ICRU", (31, 31, 30), (1.0, 1.0, 1.0)),
                     spectrum="mohan6", coll=(-5, 5, -5, 5), ssd=100.0, ecut=0.700, nsplit=1, nhist=300_000_000, nbatch=40),
    # BASELINE config 3: TG119-like heterogeneous phantom (lung + bone inserts), var_6MV
    "tg119_6mv": dict(media="media_700_tissue4.blob", phantom=lambda: P.tissue_phantom((50, 50, 32), (1.0, 1.0, 1.0), "tg119"),
                      spectrum="var_6MV", coll=(-5, 5, -5, 5), ssd=90.0, ecut=0.700, nsplit=1, nhist=300_000_000, nbatch=40),
    # BASELINE headline / config 5: PROSTATE-like phantom, var_6MV (the bench workload on 9 mm voxels)
    "prostate6mv": dict(media="media_700_tissue4.blob", phantom=lambda: P.tissue_phantom((61, 61, 30), (0.9, 0.9, 0.9), "prostate"),
                        spectrum="var_6MV", coll=(-5, 5, -5, 5), ssd=90.0, ecut=0.700, nsplit=1, nhist=300_000_000, nbatch=40),
    # the reference's shipped input file (ucodes/omc_dosxyz/input_file.inp): mohan6, 5x5 cm2, SSD 90, 521icru, nsplit = 20
    "water_inp_ns20": dict(media="media_521_water.blob", phantom=lambda: P.water_phantom("H2O521ICRU", (31, 31, 30), (1.0, 1.0, 1.0)),
                           spectrum="mohan6", coll=(-2.5, 2.5, -2.5, 2.5), ssd=90.0, ecut=0.521, nsplit=20, nhist=4_000_000, nbatch=40),
}


def build_case(name: str):
    c = CASES[name]
    media = P.load_blob(P.golden(c["media"]))
    ph = c["phantom"]()
    prob = P.build_problem(media, ph, ecut=c["ecut"], pcut=0.010, collimator=c["coll"], ssd=c["ssd"],
                           cdfinv=(media["cdfinv1_" + c["spectrum"]], media["cdfinv2_" + c["spectrum"]]), nsplit=c["nsplit"])
    return prob, ph, c


def fixture_path(name: str) -> str:
    return os.path.join(ROOT, "tests", "golden", f"parity_{name}.npz")

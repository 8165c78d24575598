"""Sampler-level parity on the GPU ("unit: device fn vs oracle fn with an injected RNG sequence"):
single particles of chosen kind and energy are pushed through shower() on the device (lock-step kernel,
omc_gpu_test_particles) and through the CPU oracle with the same per-history Philox stream.  Every sampler of
SURVEY 8a is the FIRST thing that happens to at least one of these particle classes:
  low / medium / high-energy photons  -> photo, Rayleigh, Compton (both k0 regimes), pair (all three regimes)
  electrons 0.8 .. 50 MeV             -> msdist/mscat (all lambda regimes), sscat, Moller, brems (both BH regimes)
  positrons                           -> Bhabha, annihilation in flight and at rest
"""
import numpy as np
import pytest

from oracle.gen_fixtures import golden_problem

pytestmark = pytest.mark.gpu
RM = 0.5109989461

CLASSES = [
    # (name, charge, kinetic energies [MeV], n per energy)
    ("photon", 0, [0.015, 0.03, 0.08, 0.3, 0.9, 1.5, 2.05, 3.0, 8.0, 20.0, 45.0, 52.0], 400),
    ("electron", -1, [0.8, 1.5, 3.0, 6.0, 15.0, 30.0, 51.0], 60),
    ("positron", 1, [0.3, 1.0, 3.0, 10.0, 25.0], 60),
]


@pytest.mark.parametrize("problem", ["golden_water521_250kV", "golden_tissue4_6MV"])
@pytest.mark.parametrize("name,iq,energies,nper", CLASSES, ids=[c[0] for c in CLASSES])
def test_single_particle_showers_match_oracle(gpu, oracle_lib, problem, name, iq, energies, nper):
    prob, ph, cfg = golden_problem(problem)
    gpu.load_problem(prob)
    oracle_lib.set_num_threads(1)
    oracle_lib.load_problem(prob)
    rng = np.random.default_rng(11)
    n = nper * len(energies)
    ekin = np.repeat(np.asarray(energies, dtype=np.float64), nper)
    e = ekin + (RM if iq != 0 else 0.0)
    # start somewhere inside the phantom, random direction, in the region that contains the point
    ix = rng.integers(1, ph.isize - 1, n); iy = rng.integers(1, ph.jsize - 1, n); iz = rng.integers(0, ph.ksize - 1, n)
    f = rng.random((n, 3)) * 0.98 + 0.01
    pos = np.column_stack([ph.xbounds[ix] + f[:, 0] * np.diff(ph.xbounds)[ix], ph.ybounds[iy] + f[:, 1] * np.diff(ph.ybounds)[iy],
                           ph.zbounds[iz] + f[:, 2] * np.diff(ph.zbounds)[iz]])
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    ir = (1 + ix + iy * ph.isize + iz * ph.isize * ph.jsize).astype(np.int32)
    q = np.column_stack([pos, d])
    first = 31000
    rg = gpu.test_particles(np.full(n, iq), e, q, ir, first_history=first)
    ro = np.zeros(n, dtype=rg.dtype)
    for i in range(n):
        ro[i] = oracle_lib.run_particle(first + i, iq, float(e[i]), q[i], int(ir[i]))
    same = (rg["ndraws"] == ro["ndraws"]) & (rg["ndeposit"] == ro["ndeposit"])
    assert same.mean() >= 0.995, f"{(~same).sum()} of {n} {name} showers left lock-step"
    rel = np.abs(rg["edep"][same] - ro["edep"][same]) / np.maximum(np.abs(ro["edep"][same]), 1e-30)
    assert rel.max() < 1e-9
    assert (rg["flags"] & 1).sum() == 0
    # energy bookkeeping: nothing deposits more than it carried (positrons add 2 m_e c^2)
    budget = ekin + (2 * RM if iq > 0 else 0.0)
    assert (rg["edep"] <= budget * (1 + 1e-12)).all()

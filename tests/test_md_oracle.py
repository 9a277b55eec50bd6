"""oracle/md_oracle.py against the integrator steps and kinetic energies recorded in the reference's own OpenMM trajectory
fixtures (tests/golden/langevin_2olx_pairs.npz, extracted by tests/golden/make_golden.py::md_case from
simulation/testdata/implicit-2olx-traj*-arrays.npz and testdata/output/2olx-traj-arrays.npz)."""
import os

import numpy as np

from oracle import md_oracle as mo
from timewarp_b200.peptides import ATOMIC_MASS, tetrapeptide_2olx

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    g = np.load(os.path.join(GOLDEN, "langevin_2olx_pairs.npz"))
    masses = np.array([ATOMIC_MASS[e] for e in g["elements"]])
    kT = mo.MOLAR_GAS_CONSTANT_R * float(g["temperature_K"])
    return g, masses, float(g["timestep_ps"]), float(g["friction_per_ps"]), kT


def test_fixture_is_the_2olx_peptide_of_the_package():
    g, masses, *_ = _load()
    pep = tetrapeptide_2olx()
    assert list(g["elements"]) == list(pep.elements)
    np.testing.assert_array_equal(pep.masses, masses)


def test_langevin_step_reproduces_recorded_steps():
    """x' = x + dt v' with v' from the update rule and the noise the step must have drawn; the implied noise is N(0,1)."""
    g, masses, dt, friction, kT = _load()
    x0, v0, f0, x1, v1 = (g[k].astype(np.float64) for k in ("x0", "v0", "f0", "x1", "v1"))
    xi = mo.langevin_implied_noise(v0, v1, f0, masses, dt, friction, kT)
    xn, vn = mo.langevin_step(x0, v0, f0, xi, masses, dt, friction, kT)
    np.testing.assert_allclose(vn, v1, rtol=0, atol=1e-9)  # (closes the loop on the algebra)
    np.testing.assert_allclose(xn, x1, rtol=0, atol=2e-7)  # fp32 storage of the fixture: 1 ulp at |x| ~ 2 nm = 2.4e-7
    n = xi.size
    assert abs(xi.mean()) < 4 / np.sqrt(n)
    assert abs(xi.std() - 1.0) < 4 / np.sqrt(2 * n), xi.std()  # pins sqrt(kT (1 - a^2) / m) (a 2 % error in kT or m would fail)
    # per-element variance: the 1/sqrt(m) scaling holds for hydrogens and heavy atoms separately
    is_h = np.asarray(g["elements"]) == "H"
    for sel in (is_h, ~is_h):
        s = xi[:, sel]
        assert abs(s.std() - 1.0) < 4 / np.sqrt(2 * s.size)
    # a wrong rule is rejected: the velocity-Verlet / "middle" position update does not reproduce the recorded positions
    xm, _ = mo.langevin_middle_step(x0, v0, f0, xi, masses, dt, friction, kT)
    assert np.abs(xm - x1).max() > 1e-5


def test_leapfrog_kinetic_energy_matches_openmm():
    """The kinetic energies OpenMM recorded (checked by the reference at simulation/tests/test_md.py:35-47)."""
    g, masses, dt, *_ = _load()
    ke = mo.leapfrog_kinetic_energy(g["ke_velocities"].astype(np.float64), g["ke_forces"].astype(np.float64), masses, dt)
    np.testing.assert_allclose(ke, g["ke_openmm"], rtol=0, atol=2e-5)
    # rounded textbook masses (H 1.008 ...) miss by ~2e-3 kJ/mol: the fixture pins OpenMM's element table
    rounded = np.array([{"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999}[e] for e in g["elements"]])
    ke_r = mo.leapfrog_kinetic_energy(g["ke_velocities"].astype(np.float64), g["ke_forces"].astype(np.float64), rounded, dt)
    assert np.abs(ke_r - g["ke_openmm"]).max() > 1e-3


def test_integrators_agree_in_the_small_step_limit():
    """Both rules integrate the same SDE: one noiseless step differs by O(dt^2)."""
    rng = np.random.default_rng(0)
    m = np.array([1.0, 12.0, 16.0])
    x, v, f = rng.standard_normal((3, 3)), rng.standard_normal((3, 3)), rng.standard_normal((3, 3)) * 100
    z = np.zeros((3, 3))
    for dt in (1e-3, 1e-4):
        xa, va = mo.langevin_step(x, v, f, z, m, dt, 0.3, 2.5)
        xb, vb = mo.langevin_middle_step(x, v, f, z, m, dt, 0.3, 2.5)
        assert np.abs(xa - xb).max() < 60 * dt * dt and np.abs(va - vb).max() < 60 * dt * dt

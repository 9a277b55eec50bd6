"""fp64 numpy oracle of the OpenMM integrator steps the reference's MH driver can interleave with the flow proposals
(`openmm_step`, utils/evaluation_utils.py:439-464 -> `sim.step(n)`; integrators of simulation/md.py:116-123, constraints=None
md.py:171,180).  TEST INFRASTRUCTURE ONLY.

The arithmetic lives in OpenMM 7.7 (third-party, timewarp-environment.yml:22, not on this machine); it is restated here from
OpenMM's documented update rules and PINNED for `LangevinIntegrator` by the reference's own trajectory fixtures
(simulation/testdata/implicit-2olx-traj-cpu-arrays.npz, implicit-2olx-traj-arrays.npz, testdata/output/2olx-traj-arrays.npz:
preset "T1-peptides" = LangevinIntegrator, 310 K, 0.3 / ps, 0.5 fs, simulation/md.py:75-82): they contain consecutive integrator
steps (step 20000 -> 20001, ...) with positions, velocities and forces, for which
  * x' = x + dt v'                                                         holds to fp32 round-off (6e-8 nm),
  * xi = (v' - a v - (1-a)/gamma F/m) sqrt(m) / sqrt(kT (1-a^2))           has mean 0 and unit variance,
  * the kinetic energy OpenMM reports is 1/2 sum m (v + dt/2 F/m)^2        (leapfrog half-step shift) to 1e-5 kJ/mol with
    OpenMM's element masses (H 1.007947, C 12.01078, N 14.00672, O 15.99943).
tests/golden/make_golden.py extracts those frames into tests/golden/langevin_2olx_pairs.npz; tests/test_md_oracle.py checks
this file against them.  `LangevinMiddleIntegrator` is PARITY UNPINNED (no fixture with consecutive frames was generated
with it); it follows the OpenMM >= 7.5 documentation.
"""
from __future__ import annotations

import numpy as np

MOLAR_GAS_CONSTANT_R = 8.31446261815324e-3  # kJ/mol/K (openmm.unit.MOLAR_GAS_CONSTANT_R)


def langevin_constants(dt: float, friction: float, kT: float):
    a = np.exp(-dt * friction)
    fscale = dt if friction == 0 else (1.0 - a) / friction
    return a, fscale, np.sqrt(kT * (1.0 - a * a))


def langevin_step(x, v, forces, xi, masses, dt, friction, kT):
    """One LangevinIntegrator step.  x, v, forces, xi: [..., N, 3]; masses [N]."""
    a, fscale, nscale = langevin_constants(dt, friction, kT)
    m = np.asarray(masses, dtype=np.float64)[:, None]
    v1 = a * v + fscale * forces / m + nscale / np.sqrt(m) * xi
    return x + dt * v1, v1


def langevin_implied_noise(v0, v1, forces0, masses, dt, friction, kT):
    """The standard normals a recorded LangevinIntegrator step must have drawn."""
    a, fscale, nscale = langevin_constants(dt, friction, kT)
    m = np.asarray(masses, dtype=np.float64)[:, None]
    return (v1 - a * v0 - fscale * forces0 / m) * np.sqrt(m) / nscale


def langevin_middle_step(x, v, forces, xi, masses, dt, friction, kT):
    """One LangevinMiddleIntegrator step (forces evaluated at x)."""
    a, _, nscale = langevin_constants(dt, friction, kT)
    m = np.asarray(masses, dtype=np.float64)[:, None]
    v = v + dt * forces / m
    x = x + 0.5 * dt * v
    v = a * v + nscale / np.sqrt(m) * xi
    return x + 0.5 * dt * v, v


def leapfrog_kinetic_energy(v, forces, masses, dt):
    """Kinetic energy as OpenMM reports it for a leapfrog integrator (velocities are half a step behind the positions)."""
    m = np.asarray(masses, dtype=np.float64)[:, None]
    vs = v + 0.5 * dt * forces / m
    return 0.5 * (m * vs * vs).sum((-1, -2))


def numerical_forces(energy_fn, x, h=1e-6):
    """-dU/dx by central differences of an energy function [B,N,3] -> [B] (fp64)."""
    x = np.asarray(x, dtype=np.float64)
    f = np.zeros_like(x)
    for i in range(x.shape[-2]):
        for k in range(3):
            xp, xm = x.copy(), x.copy()
            xp[..., i, k] += h
            xm[..., i, k] -= h
            f[..., i, k] = -(energy_fn(xp) - energy_fn(xm)) / (2 * h)
    return f


def integrate(energy_fn, x, v, noise, masses, dt, friction, kT, middle=False, h=1e-6):
    """n = noise.shape[0] steps with forces from central differences of `energy_fn` (small systems only)."""
    x, v = np.asarray(x, dtype=np.float64), np.asarray(v, dtype=np.float64)
    step = langevin_middle_step if middle else langevin_step
    for xi in np.asarray(noise, dtype=np.float64):
        x, v = step(x, v, numerical_forces(energy_fn, x, h), xi, masses, dt, friction, kT)
    return x, v

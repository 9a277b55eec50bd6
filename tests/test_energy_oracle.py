"""Closed-form / known-answer checks of the fp64 energy oracle (CPU only)."""
import math

import numpy as np

from oracle import energy_oracle as eo
from timewarp_b200.forcefield import ONE_4PI_EPS0, SystemDescription, amber_like_system
from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx


def _empty_system(n, **kw):
    z = lambda *s, dt=np.float64: np.zeros(s, dtype=dt)  # noqa: E731
    excl = np.eye(n, dtype=np.uint8)
    base = dict(
        n_atoms=n, bond_idx=z(0, 2, dt=np.int32), bond_param=z(0, 2), angle_idx=z(0, 3, dt=np.int32), angle_param=z(0, 2),
        torsion_idx=z(0, 4, dt=np.int32), torsion_param=z(0, 3), charge=z(n), sigma=z(n), epsilon=z(n), excluded=excl,
        exception_idx=z(0, 2, dt=np.int32), exception_param=z(0, 3), gb_radius=np.full(n, 0.15), gb_scale=np.full(n, 0.8),
        masses=np.ones(n), use_gb=False,
    )  # fmt: skip
    base.update(kw)
    return SystemDescription(**base)


def test_o2_harmonic_bond():
    # O2 toy system of the reference: k=248940 kJ/mol/nm^2, r0=0.1016 nm (utils/evaluation_utils_o2.py:9-17,33)
    s = _empty_system(2, bond_idx=np.array([[0, 1]], dtype=np.int32), bond_param=np.array([[0.1016, 248940.0]]))
    s.excluded[:] = 1
    x = np.array([[[0, 0, 0], [0.11, 0, 0]]], dtype=np.float64)
    assert np.isclose(eo.potential_energy(s, x)[0], 0.5 * 248940.0 * (0.11 - 0.1016) ** 2, rtol=1e-12)


def test_angle_and_torsion():
    s = _empty_system(4, angle_idx=np.array([[0, 1, 2]], dtype=np.int32), angle_param=np.array([[math.radians(100.0), 400.0]]),
                      torsion_idx=np.array([[0, 1, 2, 3]], dtype=np.int32), torsion_param=np.array([[3.0, 0.5, 2.0]]))
    s.excluded[:] = 1
    phi = math.radians(60.0)
    x = np.array([[[1, 0, 0], [0, 0, 0], [0, 0, 1], [math.cos(phi), math.sin(phi), 1]]], dtype=np.float64)
    t = eo.energy_terms(s, x)[0]
    assert np.isclose(t[1], 0.5 * 400.0 * (math.radians(90.0) - math.radians(100.0)) ** 2, rtol=1e-12)
    assert np.isclose(abs(eo._dihedral(*x[0])), phi, rtol=1e-12)
    assert np.isclose(t[2], 2.0 * (1 + math.cos(3 * eo._dihedral(*x[0]) - 0.5)), rtol=1e-12)


def test_pair_lj_coulomb_reaction_field():
    s = _empty_system(2, charge=np.array([0.5, -0.4]), sigma=np.array([0.3, 0.34]), epsilon=np.array([0.4, 0.9]))
    r = 0.37
    x = np.array([[[0, 0, 0], [r, 0, 0]]], dtype=np.float64)
    sig, eps = 0.32, math.sqrt(0.36)
    lj = 4 * eps * ((sig / r) ** 12 - (sig / r) ** 6)
    # eps_rf = 1 -> k_rf = 0, c_rf = 1/rc: shifted Coulomb
    assert np.isclose(eo.potential_energy(s, x)[0], lj + ONE_4PI_EPS0 * (-0.2) * (1 / r - 1 / 2.0), rtol=1e-12)
    s.reaction_field_eps = 78.3
    krf = (78.3 - 1) / ((2 * 78.3 + 1) * 8.0)
    crf = 3 * 78.3 / ((2 * 78.3 + 1) * 2.0)
    assert np.isclose(eo.potential_energy(s, x)[0], lj + ONE_4PI_EPS0 * (-0.2) * (1 / r + krf * r * r - crf), rtol=1e-12)
    # beyond the cutoff the pair is dropped
    x[0, 1, 0] = 2.5
    assert eo.potential_energy(s, x)[0] == 0.0


def test_gb_single_ion_born_energy():
    # one isolated ion: Born radius = offset radius, E = -1/2 * 138.9 * (1/eps_in - 1/eps_out) q^2 / B  (+ ACE term)
    s = _empty_system(1, charge=np.array([1.0]), use_gb=True, gb_radius=np.array([0.2]), surface_area_energy=0.0, cutoff=0.0)
    e = eo.potential_energy(s, np.zeros((1, 1, 3)))[0]
    born = 0.2 - 0.009
    assert np.isclose(e, -0.5 * ONE_4PI_EPS0 * (1 - 1 / 78.5) / born, rtol=1e-12)
    s.surface_area_energy = 28.3919551
    e2 = eo.potential_energy(s, np.zeros((1, 1, 3)))[0]
    assert np.isclose(e2 - e, 28.3919551 * (0.2 + 0.14) ** 2 * (0.2 / born) ** 6, rtol=1e-12)


def test_gb_far_pair_limit():
    # two far ions: Born radii -> isolated values, pair term -> -138.9 (1-1/78.5) q1 q2 / r
    s = _empty_system(2, charge=np.array([1.0, -1.0]), use_gb=True, gb_radius=np.array([0.15, 0.15]), surface_area_energy=0.0, cutoff=0.0)
    s.excluded[:] = 1
    r = 50.0
    x = np.array([[[0, 0, 0], [r, 0, 0]]], dtype=np.float64)
    e = eo.potential_energy(s, x)[0]
    born = 0.15 - 0.009
    expect = 2 * (-0.5 * ONE_4PI_EPS0 * (1 - 1 / 78.5) / born) + ONE_4PI_EPS0 * (1 - 1 / 78.5) / r
    assert np.isclose(e, expect, rtol=1e-6)


def test_synthetic_system_counts_and_invariance():
    for pep, nb, na in ((alanine_dipeptide(), 21, 36), (tetrapeptide_2olx(), 64, 111)):
        s = amber_like_system(pep)
        assert len(s.bond_idx) == nb and len(s.angle_idx) == na  # tree topologies: N-1 bonds
        assert abs(s.charge.sum()) < 1e-12
        x = pep.coords_nm[None]
        e = eo.potential_energy(s, x)[0]
        assert np.isfinite(e)
        # rigid-motion invariance (size-independent property)
        th = 0.7
        R = np.array([[math.cos(th), -math.sin(th), 0], [math.sin(th), math.cos(th), 0], [0, 0, 1]])
        e_rot = eo.potential_energy(s, x @ R.T + np.array([0.3, -1.0, 2.0]))[0]
        assert np.isclose(e, e_rot, rtol=1e-10, atol=1e-8)

"""openmm.System -> SystemDescription (timewarp_b200.forcefield.system_description_from_openmm).

Without OpenMM the extraction is exercised on OpenMM-shaped stand-ins (same getter names and tuple layouts as the OpenMM 7.7
Python API); with OpenMM installed the last test rebuilds the reference's System (simulation/md.py:149-173, preset
"T1-peptides") and reproduces the golden potential energies of simulation/testdata/implicit-2olx-traj-cpu-arrays.npz at the
reference's own tolerance (simulation/tests/test_md.py:35-47) -- the check that pins oracle/energy_oracle.py to the reference."""
import io
import os

import numpy as np
import pytest

from oracle import energy_oracle as eo
from timewarp_b200.forcefield import amber_like_system, system_description_from_openmm
from timewarp_b200.peptides import tetrapeptide_2olx

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Q:  # a Quantity-like wrapper: the extractor must unwrap it
    def __init__(self, v):
        self._value = v


class HarmonicBondForce:
    def __init__(self, d):
        self.d = d

    def getNumBonds(self):
        return len(self.d.bond_idx)

    def getBondParameters(self, i):
        return int(self.d.bond_idx[i, 0]), int(self.d.bond_idx[i, 1]), _Q(self.d.bond_param[i, 0]), _Q(self.d.bond_param[i, 1])


class HarmonicAngleForce:
    def __init__(self, d):
        self.d = d

    def getNumAngles(self):
        return len(self.d.angle_idx)

    def getAngleParameters(self, i):
        return (*map(int, self.d.angle_idx[i]), _Q(self.d.angle_param[i, 0]), _Q(self.d.angle_param[i, 1]))


class PeriodicTorsionForce:
    def __init__(self, d):
        self.d = d

    def getNumTorsions(self):
        return len(self.d.torsion_idx)

    def getTorsionParameters(self, i):
        return (*map(int, self.d.torsion_idx[i]), int(self.d.torsion_param[i, 0]), _Q(self.d.torsion_param[i, 1]), _Q(self.d.torsion_param[i, 2]))


class NonbondedForce:
    def __init__(self, d):
        self.d = d
        n = d.n_atoms
        live = {(int(a), int(b)): tuple(p) for (a, b), p in zip(d.exception_idx, d.exception_param)}
        self.exc = []
        for a in range(n):
            for b in range(a + 1, n):
                if d.excluded[a, b]:
                    q, s, e = live.get((a, b), live.get((b, a), (0.0, 1.0, 0.0)))
                    self.exc.append((a, b, _Q(q), _Q(s), _Q(e)))

    def getNonbondedMethod(self):
        return 1 if self.d.cutoff > 0 else 0

    def getParticleParameters(self, i):
        return _Q(self.d.charge[i]), _Q(self.d.sigma[i]), _Q(self.d.epsilon[i])

    def getNumExceptions(self):
        return len(self.exc)

    def getExceptionParameters(self, i):
        return self.exc[i]

    def getCutoffDistance(self):
        return _Q(self.d.cutoff)

    def getReactionFieldDielectric(self):
        return self.d.reaction_field_eps


class GBSAOBCForce:
    def __init__(self, d):
        self.d = d

    def getParticleParameters(self, i):
        return _Q(self.d.charge[i]), _Q(self.d.gb_radius[i]), self.d.gb_scale[i]

    def getSoluteDielectric(self):
        return self.d.solute_dielectric

    def getSolventDielectric(self):
        return self.d.solvent_dielectric

    def getSurfaceAreaEnergy(self):
        return _Q(self.d.surface_area_energy / (4.0 * np.pi))  # OpenMM reports sigma; the kernel field is 4 pi sigma


class CMMotionRemover:
    pass


class _System:
    def __init__(self, d, forces):
        self.d, self.forces = d, forces

    def getNumParticles(self):
        return self.d.n_atoms

    def getParticleMass(self, i):
        return _Q(self.d.masses[i])

    def getForces(self):
        return self.forces


def _fake(d):
    return _System(d, [HarmonicBondForce(d), HarmonicAngleForce(d), PeriodicTorsionForce(d), NonbondedForce(d), GBSAOBCForce(d), CMMotionRemover()])


def test_extraction_round_trip():
    pep = tetrapeptide_2olx()
    d = amber_like_system(pep)  # OBC2, cutoff 2 nm, reaction-field eps 1: the "T1-peptides" System shape
    got = system_description_from_openmm(_fake(d), temperature=310.0)
    for f in ("bond_idx", "bond_param", "angle_idx", "angle_param", "torsion_idx", "torsion_param", "charge", "sigma", "epsilon", "excluded",
              "gb_radius", "gb_scale", "masses"):
        np.testing.assert_array_equal(getattr(got, f), getattr(d, f), err_msg=f)
    assert sorted(map(tuple, got.exception_idx.tolist())) == sorted(tuple(sorted(p)) for p in d.exception_idx.tolist())
    for f in ("cutoff", "reaction_field_eps", "use_gb", "gb_alpha", "gb_beta", "gb_gamma", "solute_dielectric", "solvent_dielectric", "n_atoms"):
        assert getattr(got, f) == getattr(d, f), f
    assert abs(got.surface_area_energy - d.surface_area_energy) < 1e-12 and abs(d.surface_area_energy / (4 * np.pi) - 2.25936) < 1e-5
    x = pep.coords_nm[None].astype(np.float64) + 0.01 * np.random.default_rng(0).standard_normal((3, pep.num_atoms, 3))
    np.testing.assert_allclose(eo.potential_energy(got, x), eo.potential_energy(d, x), rtol=1e-12)


def test_unsupported_forces_are_rejected():
    d = amber_like_system(tetrapeptide_2olx())

    class CustomGBForce:
        pass

    with pytest.raises(NotImplementedError, match="CustomGBForce"):
        system_description_from_openmm(_System(d, [CustomGBForce()]))
    pme = NonbondedForce(d)
    pme.getNonbondedMethod = lambda: 4
    with pytest.raises(NotImplementedError, match="PME"):
        system_description_from_openmm(_System(d, [pme]))


def test_reference_golden_energies_with_openmm():
    """Runs where OpenMM (and its Amber XML files) is installed: the reference's own golden-vector check."""
    openmm = pytest.importorskip("openmm")
    from openmm import app, unit

    g = np.load(os.path.join(GOLDEN, "langevin_2olx_pairs.npz"))
    pdb = app.PDBFile(io.StringIO(str(g["state0_pdb"])))
    ff = app.ForceField("amber99sbildn.xml", "amber99_obc.xml")  # simulation/md.py:151-152
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffNonPeriodic, nonbondedCutoff=2.0 * unit.nanometer, constraints=None)
    sysd = system_description_from_openmm(system)
    e = eo.potential_energy(sysd, g["pot_positions"].astype(np.float64))
    np.testing.assert_allclose(e, g["pot_openmm"], rtol=0, atol=1e-3)  # simulation/tests/test_md.py:35-47
    assert openmm is not None

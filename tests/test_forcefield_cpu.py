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
from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx

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


# ------------------------------------------------------------------------------------------------ pinned table (no OpenMM needed)
def _golden_energy():
    return np.load(os.path.join(GOLDEN, "energy_2olx_openmm.npz"))


def _c_terminal_carbon(pep):
    last = max(pep.residue_index)
    return [i for i, (n, r) in enumerate(zip(pep.atom_names, pep.residue_index)) if n == "C" and r == last][0]


def test_amber99sbildn_table_is_consistent():
    from timewarp_b200 import amber99 as A
    from timewarp_b200.forcefield import AMBER99SBILDN_PINNED, amber99sbildn_obc2

    assert AMBER99SBILDN_PINNED
    for name, atoms in A.RESIDUES.items():  # ff94 charge sets sum to the residue's formal charge
        total = sum(q for _, q in atoms.values())
        expect = {"NASN": 1.0, "CGLN": -1.0}.get(name, None)
        if expect is not None:
            assert abs(total - expect) < 1e-9, (name, total)
        elif name in ("ALA", "ASN", "GLN"):
            assert abs(total) < 1e-9, (name, total)
    pep = tetrapeptide_2olx()
    s = amber99sbildn_obc2(pep)
    assert s.n_atoms == 65 and len(s.bond_idx) == 64 and len(s.angle_idx) == 111 and abs(s.charge.sum()) < 1e-9
    assert s.cutoff == 2.0 and s.reaction_field_eps == 1.0 and s.solvent_dielectric == 78.5  # simulation/md.py:166-171 + GB post-processing
    ad = amber99sbildn_obc2(alanine_dipeptide())
    assert ad.n_atoms == 22 and abs(ad.charge.sum()) < 1e-3  # ACE + ALA + NME: 0.0001 from the published rounding of the ACE set


def test_reference_golden_energies():
    """The reference's own check, simulation/tests/test_md.py:35-47, WITHOUT OpenMM: the fp64 oracle fed with the typed-in
    ff99SB-ILDN / OBC2 table reproduces the 40 OpenMM potential energies of implicit-2olx-traj-cpu-arrays.npz.  The reference
    compares two runs of the same single-precision OpenMM platform at atol 1e-3; against an fp64 evaluation the floor is that
    platform's own rounding (measured: mean -0.005, std 0.003, max 0.011 kJ/mol at energies around -1700), hence atol 0.02."""
    from timewarp_b200.forcefield import amber99sbildn_obc2

    g = _golden_energy()
    pep = tetrapeptide_2olx()
    e = eo.potential_energy(amber99sbildn_obc2(pep), g["cpu_positions"].astype(np.float64))
    np.testing.assert_allclose(e, g["cpu_potential"], rtol=0, atol=0.02)
    assert abs((e - g["cpu_potential"]).mean()) < 0.01 and (e - g["cpu_potential"]).std() < 0.006
    # the longer trajectory (other Asn chi1 rotamers; its carboxylate improper lists the two oxygens in the other order)
    sw = amber99sbildn_obc2(pep, improper_choice={_c_terminal_carbon(pep): 0})
    ew = eo.potential_energy(sw, g["wide_positions"].astype(np.float64))
    np.testing.assert_allclose(ew, g["wide_potential"], rtol=0, atol=0.02)
    # and the synthetic table is nowhere near (this is what "unpinned" looked like)
    assert np.abs(eo.potential_energy(amber_like_system(pep), g["cpu_positions"].astype(np.float64)) - g["cpu_potential"]).min() > 100.0


def test_reference_golden_forces():
    """Forces of the same fixture (simulation/tests/test_md.py:45-47: rtol 0.05, atol 1e-2) by central differences of the oracle."""
    from timewarp_b200.forcefield import amber99sbildn_obc2

    g = _golden_energy()
    pep = tetrapeptide_2olx()
    s = amber99sbildn_obc2(pep)
    h = 1e-5
    for frame in (0, 17, 39):
        x = g["cpu_positions"][frame].astype(np.float64)
        n = pep.num_atoms
        xp = np.repeat(x[None], 2 * n * 3, 0)
        for a in range(n):
            for k in range(3):
                xp[2 * (a * 3 + k), a, k] += h
                xp[2 * (a * 3 + k) + 1, a, k] -= h
        e = eo.potential_energy(s, xp)
        f = -((e[0::2] - e[1::2]) / (2 * h)).reshape(n, 3)
        ref = g["cpu_forces"][frame].astype(np.float64)
        np.testing.assert_allclose(f, ref, rtol=0.05, atol=0.2)
        assert np.sqrt(((f - ref) ** 2).mean()) < 0.1  # measured 0.03 kJ/mol/nm rms of ~930 (float32 storage of the reference)

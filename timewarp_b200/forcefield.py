"""System description for the energy kernel (replaces the `openmm.System` the reference builds in
simulation/md.py:128-187 -- OpenMM and the Amber XML parameter files do not exist offline).

`SystemDescription` is the wire format of SURVEY.md Appendix B: bonded term lists, per-atom
nonbonded / GB parameters, exclusion matrix, 1-4 exceptions and the scalar settings of the
reference presets (CutoffNonPeriodic 2.0 nm, reaction-field eps 1 with GB, OBC1/OBC2).

`amber_like_system()` derives angles / torsions / exclusions from a bond graph and assigns
GENERIC Amber-magnitude parameters by element and hybridisation.  It is a SYNTHETIC force field
(parameters are not ff99SB-ILDN / ff14SB; the reference's golden energies cannot be reproduced
with it) -- it gives the kernel a realistic workload with the right term counts.  Real parameter
sets can be supplied by filling a SystemDescription directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

MOLAR_GAS_CONSTANT_R = 8.31446261815324e-3  # kJ/mol/K (openmm.unit.MOLAR_GAS_CONSTANT_R)
ONE_4PI_EPS0 = 138.935456  # OpenMM 7.7 SimTKOpenMMRealType.h

GB_PRESETS = {
    # simulation/md.py:152 ("amber99sbildn.xml","amber99_obc.xml") -> OBC2 ; :159 ("implicit/obc1.xml") -> OBC1
    "obc2": dict(alpha=1.0, beta=0.8, gamma=4.85),
    "obc1": dict(alpha=0.8, beta=0.0, gamma=2.909125),
}


@dataclass
class SystemDescription:
    n_atoms: int
    bond_idx: np.ndarray  # [nb,2] int32
    bond_param: np.ndarray  # [nb,2] r0, k
    angle_idx: np.ndarray  # [na,3]
    angle_param: np.ndarray  # [na,2] theta0, k
    torsion_idx: np.ndarray  # [nt,4]
    torsion_param: np.ndarray  # [nt,3] periodicity, phase, k
    charge: np.ndarray  # [N]
    sigma: np.ndarray  # [N]
    epsilon: np.ndarray  # [N]
    excluded: np.ndarray  # [N,N] uint8
    exception_idx: np.ndarray  # [ne,2]
    exception_param: np.ndarray  # [ne,3] chargeProd, sigma, epsilon
    gb_radius: np.ndarray  # [N]
    gb_scale: np.ndarray  # [N]
    masses: np.ndarray  # [N] dalton
    cutoff: float = 2.0
    reaction_field_eps: float = 1.0
    one_4pi_eps0: float = ONE_4PI_EPS0
    use_gb: bool = True
    gb_alpha: float = 1.0
    gb_beta: float = 0.8
    gb_gamma: float = 4.85
    gb_offset: float = 0.009
    solute_dielectric: float = 1.0
    solvent_dielectric: float = 78.5
    surface_area_energy: float = 28.3919551
    temperature: float = 310.0  # simulation/md.py:78,86

    def getNumParticles(self) -> int:  # openmm.System API used by the reference (openmm_bridge.py:277)
        return self.n_atoms

    def getParticleMass(self, i: int) -> float:
        return float(self.masses[i])

    def as_float32(self) -> "SystemDescription":
        """Round every parameter to fp32 (what the kernel reads) so that an fp64 oracle fed with the
        same description sees identical parameters."""
        import copy

        s = copy.copy(self)
        for k in ("bond_param", "angle_param", "torsion_param", "charge", "sigma", "epsilon", "exception_param", "gb_radius", "gb_scale"):
            setattr(s, k, getattr(self, k).astype(np.float32).astype(np.float64))
        return s


def _neighbors(n: int, bonds: np.ndarray) -> List[List[int]]:
    nb: List[List[int]] = [[] for _ in range(n)]
    for a, b in bonds:
        nb[int(a)].append(int(b))
        nb[int(b)].append(int(a))
    return [sorted(x) for x in nb]


def derive_angles(n: int, bonds: np.ndarray) -> np.ndarray:
    nb = _neighbors(n, bonds)
    out = []
    for j in range(n):
        for a in range(len(nb[j])):
            for b in range(a + 1, len(nb[j])):
                out.append((nb[j][a], j, nb[j][b]))
    return np.array(out, dtype=np.int32).reshape(-1, 3)


def derive_proper_torsions(n: int, bonds: np.ndarray) -> np.ndarray:
    nb = _neighbors(n, bonds)
    out = []
    for j, k in bonds:
        j, k = int(j), int(k)
        for i in nb[j]:
            if i == k:
                continue
            for l in nb[k]:
                if l == j or l == i:
                    continue
                out.append((i, j, k, l))
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def amber_like_system(peptide, gb: str = "obc2", total_charge: float = 0.0) -> SystemDescription:
    """Synthetic Amber-magnitude parameters for a Peptide (timewarp_b200.peptides)."""
    n = peptide.num_atoms
    el = peptide.elements
    bonds = np.asarray(peptide.bonds, dtype=np.int32)
    nb = _neighbors(n, bonds)
    deg = [len(x) for x in nb]

    def is_carbonyl_c(i):
        return el[i] == "C" and deg[i] == 3 and any(el[j] == "O" for j in nb[i])

    # ---- bonds
    bp = []
    for a, b in bonds:
        pair = "".join(sorted((el[a], el[b])))
        if pair == "CH":
            r0, k = 0.1090, 284512.0
        elif pair == "HN":
            r0, k = 0.1010, 363171.2
        elif pair == "CC":
            r0, k = (0.1522, 265265.6) if (is_carbonyl_c(a) or is_carbonyl_c(b)) else (0.1526, 259408.0)
        elif pair == "CN":
            r0, k = (0.1335, 410032.0) if (is_carbonyl_c(a) or is_carbonyl_c(b)) else (0.1449, 282001.6)
        elif pair == "CO":
            c = a if el[a] == "C" else b
            n_o = sum(1 for j in nb[c] if el[j] == "O")
            r0, k = (0.1250, 548940.8) if n_o == 2 else (0.1229, 476976.0)
        elif pair == "CS":
            r0, k = 0.1810, 189953.6
        elif pair == "HS":
            r0, k = 0.1336, 229283.2
        else:
            r0, k = 0.15, 250000.0
        bp.append((r0, k))
    # ---- angles
    angles = derive_angles(n, bonds)
    ap = []
    for i, j, k in angles:
        theta0 = math.radians(109.5 if deg[j] == 4 else (120.0 if deg[j] == 3 else 109.5))
        n_h = (el[i] == "H") + (el[k] == "H")
        kk = 292.88 if n_h == 2 else (418.4 if n_h == 1 else 585.76)
        ap.append((theta0, kk))
    # ---- torsions: generic X-sp3-sp3-X 3-fold, X-sp2-sp2-X (amide) 2-fold, plus sp2 impropers
    torsions = derive_proper_torsions(n, bonds)
    tp = []
    for i, j, k, l in torsions:
        n_paths = (deg[j] - 1) * (deg[k] - 1)
        if deg[j] == 3 and deg[k] == 3:
            tp.append((2.0, math.pi, 41.84 / n_paths))
        elif deg[j] == 4 and deg[k] == 4:
            tp.append((3.0, 0.0, 5.858 / n_paths))
        else:
            tp.append((3.0 if deg[j] == 4 or deg[k] == 4 else 2.0, 0.0 if (deg[j] == 4 or deg[k] == 4) else math.pi, 1.0))
    tors_list = [tuple(t) for t in torsions]
    for c in range(n):
        if deg[c] == 3 and el[c] in ("C", "N"):
            a, b, d = nb[c]
            tors_list.append((a, b, c, d))  # Amber improper: central atom third
            tp.append((2.0, math.pi, 43.932 if is_carbonyl_c(c) else 4.6024))
    torsions = np.array(tors_list, dtype=np.int32).reshape(-1, 4)
    # ---- per-atom nonbonded
    q = np.zeros(n)
    sig = np.zeros(n)
    eps = np.zeros(n)
    rad = np.zeros(n)
    scl = np.zeros(n)
    for i in range(n):
        e = el[i]
        heavy = [j for j in nb[i]]
        if e == "H":
            parent = el[heavy[0]] if heavy else "C"
            if parent == "N":
                q[i], sig[i], eps[i], rad[i] = 0.30, 0.106908, 0.0656888, 0.13
            elif parent == "O" or parent == "S":
                q[i], sig[i], eps[i], rad[i] = 0.40, 0.0, 0.0, 0.12
            else:
                q[i], sig[i], eps[i], rad[i] = 0.08, 0.247135, 0.0656888, 0.12
            scl[i] = 0.85
        elif e == "C":
            if is_carbonyl_c(i):
                q[i], sig[i], eps[i] = 0.60, 0.339967, 0.359824
            else:
                q[i], sig[i], eps[i] = -0.05 - 0.04 * sum(1 for j in nb[i] if el[j] == "H"), 0.339967, 0.457730
            rad[i], scl[i] = 0.17, 0.72
        elif e == "N":
            q[i], sig[i], eps[i], rad[i], scl[i] = -0.45 if deg[i] == 3 else -0.15, 0.325000, 0.711280, 0.155, 0.79
        elif e == "O":
            q[i], sig[i], eps[i], rad[i], scl[i] = -0.57, 0.295992, 0.878640, 0.15, 0.85
        else:  # S
            q[i], sig[i], eps[i], rad[i], scl[i] = -0.11, 0.356359, 1.046000, 0.18, 0.96
    q += (total_charge - q.sum()) / n  # neutralise (or set the net charge) uniformly
    # ---- exclusions (1-2, 1-3) and 1-4 exceptions (Coulomb / 1.2, LJ epsilon / 2)
    excl = np.zeros((n, n), dtype=np.uint8)
    np.fill_diagonal(excl, 1)
    for a, b in bonds:
        excl[a, b] = excl[b, a] = 1
    for i, j, k in angles:
        excl[i, k] = excl[k, i] = 1
    ex_pairs: Dict[Tuple[int, int], None] = {}
    for i, j, k, l in derive_proper_torsions(n, bonds):
        a, b = (int(i), int(l)) if i < l else (int(l), int(i))
        if not excl[a, b]:
            ex_pairs[(a, b)] = None
    ex_idx = np.array(sorted(ex_pairs), dtype=np.int32).reshape(-1, 2)
    for a, b in ex_idx:
        excl[a, b] = excl[b, a] = 1
    ex_par = np.array(
        [(q[a] * q[b] / 1.2, 0.5 * (sig[a] + sig[b]), 0.5 * math.sqrt(eps[a] * eps[b])) for a, b in ex_idx], dtype=np.float64
    ).reshape(-1, 3)
    g = GB_PRESETS[gb]
    return SystemDescription(
        n_atoms=n,
        bond_idx=bonds.astype(np.int32),
        bond_param=np.array(bp, dtype=np.float64).reshape(-1, 2),
        angle_idx=angles,
        angle_param=np.array(ap, dtype=np.float64).reshape(-1, 2),
        torsion_idx=torsions,
        torsion_param=np.array(tp, dtype=np.float64).reshape(-1, 3),
        charge=q, sigma=sig, epsilon=eps, excluded=excl,
        exception_idx=ex_idx, exception_param=ex_par,
        gb_radius=rad, gb_scale=scl, masses=np.asarray(peptide.masses, dtype=np.float64),
        gb_alpha=g["alpha"], gb_beta=g["beta"], gb_gamma=g["gamma"],
    )  # fmt: skip


AMBER99SBILDN_PINNED = True  # the table reproduces the reference's golden energies and forces (tests/test_forcefield_cpu.py)


def _improper_order(el, central, others, matched):
    """OpenMM's ordering of a wildcard improper (app/forcefield.py, "workaround to be more consistent with AMBER"): the
    two atoms matched by wildcards go first -- same element: lower index first; else carbon first, else the heavier --
    then the central atom, then the atom matched by the explicit class."""
    a1, a2 = [o for o in others if o != matched]
    mass = {"H": 1.0, "C": 12.0, "N": 14.0, "O": 16.0, "S": 32.0}
    e1, e2 = el[a1], el[a2]
    if e1 == e2:
        if a1 > a2:
            a1, a2 = a2, a1
    elif e1 != "C" and (e2 == "C" or mass[e1] < mass[e2]):
        a1, a2 = a2, a1
    return (a1, a2, central, matched)


def amber99sbildn_obc2(peptide, improper_choice=None) -> SystemDescription:
    """The System of the reference's "T1-peptides" preset (simulation/md.py:149-173: ForceField("amber99sbildn.xml",
    "amber99_obc.xml"), CutoffNonPeriodic 2.0 nm, no constraints) for a Peptide built from the residues in
    timewarp_b200/amber99.py.  `improper_choice` (dict, diagnostics only) overrides which of two equivalent atoms is taken
    as the explicitly matched one of an improper."""
    from . import amber99 as A

    n = peptide.num_atoms
    el = peptide.elements
    names = peptide.atom_names
    variants = A.variant_names(peptide)
    types = [A.RESIDUES[v][nm][0] for v, nm in zip(variants, names)]
    q = np.array([A.RESIDUES[v][nm][1] for v, nm in zip(variants, names)], dtype=np.float64)
    bonds = np.asarray(peptide.bonds, dtype=np.int32)
    nb = _neighbors(n, bonds)
    # ---- bonds
    bp = []
    for a, b in bonds:
        K, r0 = A._sym(A.BONDS, types[a], types[b])
        bp.append((r0 / 10.0, 2.0 * K * A.KCAL * 100.0))
    # ---- angles
    angles = derive_angles(n, bonds)
    ap = []
    for i, j, k in angles:
        par = A._sym(A.ANGLES, types[i], types[j], types[k])
        if par is None:
            raise KeyError(f"no angle parameters for {types[i]}-{types[j]}-{types[k]}")
        ap.append((math.radians(par[1]), 2.0 * par[0] * A.KCAL))
    # ---- proper torsions
    tors_idx, tors_par = [], []
    for i, j, k, l in derive_proper_torsions(n, bonds):
        key = (names[i], names[j], names[k], names[l])
        same_res = len({peptide.residue_index[a] for a in (i, j, k, l)}) == 1
        base = peptide.residue_names[i]
        terms = None
        if same_res:
            terms = A.ILDN.get((base, key)) or A.ILDN.get((base, key[::-1]))
        if terms is None:
            terms = A.proper_terms((types[i], types[j], types[k], types[l]))
        for pk, phase, per in terms:
            if pk != 0.0:
                tors_idx.append((i, j, k, l))
                tors_par.append((float(per), math.radians(phase) if phase != 180.0 else 3.14159265359, pk * A.KCAL))
    # ---- impropers (X-X-C-O, X-O2-C-O2, X-X-N-H)
    choice = improper_choice or {}
    for c in range(n):
        if len(nb[c]) != 3:
            continue
        t = types[c]
        others = list(nb[c])
        if t == "C":
            o2 = [a for a in others if types[a] == "O2"]
            o = [a for a in others if types[a] == "O"]
            if len(o2) == 2:
                matched = o2[choice.get(c, 1)]  # (CA, O, C, OXT): the order in the fixture the reference's test reads; the
                # 140-frame fixture under testdata/output/ was written with the two oxygens in the other order
                a1 = [a for a in others if a not in o2][0]
                a2 = [a for a in o2 if a != matched][0]
                tors_idx.append(_improper_order(el, c, [a1, a2, matched], matched))
                tors_par.append((2.0, 3.14159265359, A.IMPROPER_C_O2 * A.KCAL))
            elif len(o) == 1:
                tors_idx.append(_improper_order(el, c, others, o[0]))
                tors_par.append((2.0, 3.14159265359, A.IMPROPER_C_O * A.KCAL))
        elif t == "N":
            hs = [a for a in others if types[a] == "H"]
            if hs:
                matched = hs[choice.get(c, len(hs) - 1)]
                tors_idx.append(_improper_order(el, c, others, matched))
                backbone = sorted(types[a] for a in others) == ["C", "CT", "H"]
                tors_par.append((2.0, 3.14159265359, (A.IMPROPER_N_H_BACKBONE if backbone else A.IMPROPER_N_H) * A.KCAL))
    # ---- nonbonded
    sig = np.array([A.LJ[t][0] * 2.0 * 2.0 ** (-1.0 / 6.0) / 10.0 for t in types])
    eps = np.array([A.LJ[t][1] * A.KCAL for t in types])
    excl = np.zeros((n, n), dtype=np.uint8)
    np.fill_diagonal(excl, 1)
    for a, b in bonds:
        excl[a, b] = excl[b, a] = 1
    for i, j, k in angles:
        excl[i, k] = excl[k, i] = 1
    ex_pairs: Dict[Tuple[int, int], None] = {}
    for i, j, k, l in derive_proper_torsions(n, bonds):
        a, b = (int(i), int(l)) if i < l else (int(l), int(i))
        if not excl[a, b]:
            ex_pairs[(a, b)] = None
    ex_idx = np.array(sorted(ex_pairs), dtype=np.int32).reshape(-1, 2)
    for a, b in ex_idx:
        excl[a, b] = excl[b, a] = 1
    ex_par = np.array([(q[a] * q[b] * A.COULOMB14, 0.5 * (sig[a] + sig[b]), A.LJ14 * math.sqrt(eps[a] * eps[b])) for a, b in ex_idx],
                      dtype=np.float64).reshape(-1, 3)
    # ---- GB-OBC2
    rad = np.array([A.gb_radius(el[i], len(nb[i]), el[nb[i][0]]) for i in range(n)])
    scl = np.array([A.GB_SCALE[e] for e in el])
    g = GB_PRESETS["obc2"]
    return SystemDescription(
        n_atoms=n, bond_idx=bonds.astype(np.int32), bond_param=np.array(bp, dtype=np.float64).reshape(-1, 2),
        angle_idx=angles, angle_param=np.array(ap, dtype=np.float64).reshape(-1, 2),
        torsion_idx=np.array(tors_idx, dtype=np.int32).reshape(-1, 4), torsion_param=np.array(tors_par, dtype=np.float64).reshape(-1, 3),
        charge=q, sigma=sig, epsilon=eps, excluded=excl, exception_idx=ex_idx, exception_param=ex_par,
        gb_radius=rad, gb_scale=scl, masses=np.asarray(peptide.masses, dtype=np.float64),
        gb_alpha=g["alpha"], gb_beta=g["beta"], gb_gamma=g["gamma"], solvent_dielectric=78.5,
    )  # fmt: skip


# ------------------------------------------------------------------------------------------------
# openmm.System -> SystemDescription (where OpenMM exists: the reference builds the System in simulation/md.py:128-187)
def _val(q):
    """Plain float of an OpenMM Quantity in its default MD unit system (nm, ps, kJ/mol, radian, elementary charge, dalton)."""
    if hasattr(q, "value_in_unit_system"):
        try:
            import openmm.unit as u  # noqa: WPS433 (only reachable where OpenMM is installed)

            return float(q.value_in_unit_system(u.md_unit_system))
        except ImportError:
            pass
    return float(getattr(q, "_value", q))


def system_description_from_openmm(system, temperature: float = 310.0) -> SystemDescription:
    """Read the forces of an `openmm.System` created by `simulation.md.get_system` (simulation/md.py:128-187) for the
    implicit-solvent presets: HarmonicBondForce, HarmonicAngleForce, PeriodicTorsionForce, NonbondedForce (NoCutoff /
    CutoffNonPeriodic) and GBSAOBCForce; CMMotionRemover carries no energy.  Forces are dispatched on their class NAME, so
    anything exposing the same getters works (tests/test_forcefield_cpu.py feeds an OpenMM-shaped stand-in).  Explicit-solvent
    systems (PME) and CustomGBForce-based GB models (`implicit/obc1.xml` of OpenMM >= 7.6) are rejected, not approximated."""
    n = int(system.getNumParticles())
    kw: Dict[str, object] = dict(
        n_atoms=n, masses=np.array([_val(system.getParticleMass(i)) for i in range(n)], dtype=np.float64),
        bond_idx=np.zeros((0, 2), np.int32), bond_param=np.zeros((0, 2)), angle_idx=np.zeros((0, 3), np.int32),
        angle_param=np.zeros((0, 2)), torsion_idx=np.zeros((0, 4), np.int32), torsion_param=np.zeros((0, 3)),
        charge=np.zeros(n), sigma=np.ones(n), epsilon=np.zeros(n), excluded=np.zeros((n, n), np.uint8),
        exception_idx=np.zeros((0, 2), np.int32), exception_param=np.zeros((0, 3)), gb_radius=np.zeros(n), gb_scale=np.zeros(n),
        use_gb=False, cutoff=0.0, temperature=float(temperature))
    for force in system.getForces():
        name = type(force).__name__
        if name == "HarmonicBondForce":
            p = [force.getBondParameters(i) for i in range(force.getNumBonds())]
            kw["bond_idx"] = np.array([[a, b] for a, b, _, _ in p], np.int32).reshape(-1, 2)
            kw["bond_param"] = np.array([[_val(r0), _val(k)] for _, _, r0, k in p], np.float64).reshape(-1, 2)
        elif name == "HarmonicAngleForce":
            p = [force.getAngleParameters(i) for i in range(force.getNumAngles())]
            kw["angle_idx"] = np.array([[a, b, c] for a, b, c, _, _ in p], np.int32).reshape(-1, 3)
            kw["angle_param"] = np.array([[_val(t0), _val(k)] for _, _, _, t0, k in p], np.float64).reshape(-1, 2)
        elif name == "PeriodicTorsionForce":
            p = [force.getTorsionParameters(i) for i in range(force.getNumTorsions())]
            kw["torsion_idx"] = np.array([[a, b, c, d] for a, b, c, d, _, _, _ in p], np.int32).reshape(-1, 4)
            kw["torsion_param"] = np.array([[float(per), _val(ph), _val(k)] for _, _, _, _, per, ph, k in p], np.float64).reshape(-1, 3)
        elif name == "NonbondedForce":
            method = int(force.getNonbondedMethod())  # 0 NoCutoff, 1 CutoffNonPeriodic, 2 CutoffPeriodic, 3 Ewald, 4 PME, 5 LJPME
            if method not in (0, 1):
                raise NotImplementedError("periodic / PME NonbondedForce (explicit solvent) is outside the energy kernel's scope")
            pp = [force.getParticleParameters(i) for i in range(n)]
            kw["charge"] = np.array([_val(q) for q, _, _ in pp])
            kw["sigma"] = np.array([_val(s) for _, s, _ in pp])
            kw["epsilon"] = np.array([_val(e) for _, _, e in pp])
            excluded = np.eye(n, dtype=np.uint8)  # (self pairs are never interactions; amber_like_system marks them too)
            ex_idx, ex_par = [], []
            for i in range(force.getNumExceptions()):
                a, b, qq, sig, eps = force.getExceptionParameters(i)
                excluded[a, b] = excluded[b, a] = 1  # every exception pair leaves the regular pair loop
                if _val(qq) != 0.0 or _val(eps) != 0.0:
                    ex_idx.append((a, b))
                    ex_par.append((_val(qq), _val(sig), _val(eps)))
            kw["excluded"] = excluded
            kw["exception_idx"] = np.array(ex_idx, np.int32).reshape(-1, 2)
            kw["exception_param"] = np.array(ex_par, np.float64).reshape(-1, 3)
            kw["cutoff"] = _val(force.getCutoffDistance()) if method == 1 else 0.0
            kw["reaction_field_eps"] = float(force.getReactionFieldDielectric())
        elif name == "GBSAOBCForce":  # OBC2 (amber99_obc.xml): alpha, beta, gamma = 1, 0.8, 4.85
            pp = [force.getParticleParameters(i) for i in range(n)]
            kw["gb_radius"] = np.array([_val(r) for _, r, _ in pp])
            kw["gb_scale"] = np.array([float(s) for _, _, s in pp])
            kw.update(use_gb=True, gb_alpha=1.0, gb_beta=0.8, gb_gamma=4.85, solute_dielectric=float(force.getSoluteDielectric()),
                      solvent_dielectric=float(force.getSolventDielectric()),
                      # OpenMM's ACE term is 4 pi sigma (r + 0.14)^2 (r / B)^6 with sigma = getSurfaceAreaEnergy() (2.25936 kJ/mol/nm^2
                      # by default); the kernel's `surface_area_energy` is the prefactor 4 pi sigma (28.3919551)
                      surface_area_energy=4.0 * math.pi * _val(force.getSurfaceAreaEnergy()))
        elif name == "CMMotionRemover":
            continue
        else:
            raise NotImplementedError(f"{name}: no counterpart in the energy kernel (see SystemDescription)")
    return SystemDescription(**kw)

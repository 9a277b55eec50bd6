"""Synthetic peptide inputs (no dataset is reachable offline).

Two topologies are provided, both hand-derived from standard Amber residue templates using
the atom names in the reference's fixture PDBs:

* ``alanine_dipeptide()`` : ACE-ALA-NME, 22 atoms  (reference: simulation/testdata/alanine-dipeptide.pdb)
* ``tetrapeptide_2olx()`` : ASN-ASN-GLN-GLN zwitterion, 65 atoms (reference: testdata/output/2olx-traj-state0.pdb,
  coordinates = MD frame 0 of testdata/output/2olx-traj-arrays.npz)

Element vocabulary follows the reference's ``ELEMENT_VOCAB`` (dataloader.py:24-25): C,H,N,O,S -> 0..4.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

ELEMENT_VOCAB: Dict[str, int] = {"C": 0, "H": 1, "N": 2, "O": 3, "S": 4}
# OpenMM's element masses (dalton): what `system.getParticleMass` returns in the reference (evaluate.py:303-305).  H, C, N, O are
# pinned to 1e-5 kJ/mol by the kinetic energies recorded in simulation/testdata/implicit-2olx-traj-cpu-arrays.npz
# (tests/test_md_oracle.py); S is OpenMM's tabulated value.
ATOMIC_MASS: Dict[str, float] = {"H": 1.007947, "C": 12.01078, "N": 14.00672, "O": 15.99943, "S": 32.0655}

# intra-residue bonds by atom name
_BACKBONE = [("N", "H"), ("N", "CA"), ("CA", "HA"), ("CA", "C"), ("C", "O"), ("CA", "CB")]
_TEMPLATES: Dict[str, List[Tuple[str, str]]] = {
    "ACE": [("CH3", "1HH3"), ("CH3", "2HH3"), ("CH3", "3HH3"), ("CH3", "C"), ("C", "O")],
    "NME": [("N", "H"), ("N", "CH3"), ("CH3", "1HH3"), ("CH3", "2HH3"), ("CH3", "3HH3")],
    "ALA": _BACKBONE + [("CB", "1HB"), ("CB", "2HB"), ("CB", "3HB")],
    "ASN": _BACKBONE
    + [("CB", "HB2"), ("CB", "HB3"), ("CB", "CG"), ("CG", "OD1"), ("CG", "ND2"), ("ND2", "HD21"), ("ND2", "HD22")],
    "GLN": _BACKBONE
    + [
        ("CB", "HB2"),
        ("CB", "HB3"),
        ("CB", "CG"),
        ("CG", "HG2"),
        ("CG", "HG3"),
        ("CG", "CD"),
        ("CD", "OE1"),
        ("CD", "NE2"),
        ("NE2", "HE21"),
        ("NE2", "HE22"),
    ],
}
_NTERM_EXTRA = [("N", "H2"), ("N", "H3")]
_CTERM_EXTRA = [("C", "OXT")]


@dataclass
class Peptide:
    name: str
    atom_names: List[str]
    residue_names: List[str]
    residue_index: List[int]
    coords_nm: np.ndarray  # [V,3] float64
    bonds: np.ndarray  # [E,2] int64, each bond once

    @property
    def num_atoms(self) -> int:
        return len(self.atom_names)

    @property
    def elements(self) -> List[str]:
        return [next(c for c in n if c.isalpha()) for n in self.atom_names]

    @property
    def atom_types(self) -> np.ndarray:
        return np.array([ELEMENT_VOCAB[e] for e in self.elements], dtype=np.int64)

    @property
    def masses(self) -> np.ndarray:
        return np.array([ATOMIC_MASS[e] for e in self.elements], dtype=np.float64)


def _build_bonds(atom_names: Sequence[str], residue_names: Sequence[str], residue_index: Sequence[int]) -> np.ndarray:
    res_ids = sorted(set(residue_index))
    idx = {(r, n): i for i, (r, n) in enumerate(zip(residue_index, atom_names))}
    bonds: List[Tuple[int, int]] = []
    for pos, r in enumerate(res_ids):
        rname = next(rn for rn, ri in zip(residue_names, residue_index) if ri == r)
        tmpl = list(_TEMPLATES[rname])
        if (r, "H2") in idx:
            tmpl += _NTERM_EXTRA
        if (r, "OXT") in idx:
            tmpl += _CTERM_EXTRA
        for a, b in tmpl:
            assert (r, a) in idx and (r, b) in idx, (rname, a, b)
            bonds.append((idx[(r, a)], idx[(r, b)]))
        if pos + 1 < len(res_ids):  # peptide bond C(i) - N(i+1)
            bonds.append((idx[(r, "C")], idx[(res_ids[pos + 1], "N")]))
    return np.array(sorted(set(tuple(sorted(b)) for b in bonds)), dtype=np.int64)


_AD_NAMES = "1HH3 CH3 2HH3 3HH3 C O N H CA HA CB 1HB 2HB 3HB C O N H CH3 1HH3 2HH3 3HH3".split()
_AD_RES = ["ACE"] * 6 + ["ALA"] * 10 + ["NME"] * 6
_AD_RESID = [1] * 6 + [2] * 10 + [3] * 6
_AD_XYZ = [
    [0.2, 0.1, 0.0], [0.2, 0.209, 0.0], [0.1486, 0.2454, 0.089], [0.1486, 0.2454, -0.089],
    [0.3427, 0.2641, 0.0], [0.4391, 0.1877, 0.0], [0.3555, 0.397, 0.0], [0.2733, 0.4556, 0.0],
    [0.4853, 0.4614, 0.0], [0.5408, 0.4316, 0.089], [0.5661, 0.4221, -0.1232], [0.5123, 0.4521, -0.2131],
    [0.663, 0.4719, -0.1206], [0.5809, 0.3141, -0.1241], [0.4713, 0.6129, 0.0], [0.3601, 0.6653, 0.0],
    [0.5846, 0.6835, 0.0], [0.6737, 0.6359, 0.0], [0.5846, 0.8284, 0.0], [0.4819, 0.8648, 0.0],
    [0.636, 0.8648, 0.089], [0.636, 0.8648, -0.089],
]  # fmt: skip

_2OLX_NAMES = (
    "N H H2 H3 CA HA C O CB HB2 HB3 CG OD1 ND2 HD21 HD22 "
    "N H CA HA C O CB HB2 HB3 CG OD1 ND2 HD21 HD22 "
    "N H CA HA C O CB HB2 HB3 CG HG2 HG3 CD OE1 NE2 HE21 HE22 "
    "N H CA HA C O CB HB2 HB3 CG HG2 HG3 CD OE1 NE2 HE21 HE22 OXT"
).split()
_2OLX_RES = ["ASN"] * 30 + ["GLN"] * 35
_2OLX_RESID = [1] * 16 + [2] * 14 + [3] * 17 + [4] * 18
_2OLX_XYZ = [
    [0.44950, 0.02266, 0.21321], [0.35338, 0.04583, 0.18853], [0.48829, -0.03035, 0.13837],
    [0.50722, 0.10876, 0.22344], [0.45003, -0.06597, 0.33134], [0.40479, -0.15851, 0.29901],
    [0.36665, 0.00276, 0.44456], [0.33628, 0.12391, 0.44226], [0.60168, -0.07711, 0.38383],
    [0.59621, -0.13367, 0.47555], [0.64507, 0.01671, 0.39956], [0.68723, -0.14995, 0.28228],
    [0.80315, -0.11439, 0.25454], [0.62177, -0.24679, 0.20669], [0.68928, -0.28939, 0.13707],
    [0.53617, -0.29029, 0.23344], [0.34345, -0.06859, 0.55828], [0.36518, -0.16684, 0.55696],
    [0.26610, -0.00521, 0.67003], [0.17740, 0.04389, 0.62596], [0.34533, 0.11054, 0.73747],
    [0.46596, 0.12702, 0.73167], [0.24382, -0.11630, 0.77733], [0.34193, -0.14886, 0.81965],
    [0.18926, -0.07723, 0.86497], [0.16283, -0.23110, 0.72390], [0.12630, -0.24033, 0.60692],
    [0.11622, -0.31508, 0.81375], [0.06147, -0.39289, 0.76917], [0.15179, -0.32548, 0.90415],
    [0.26447, 0.19238, 0.79827], [0.16881, 0.16444, 0.81792], [0.31258, 0.30479, 0.88748],
    [0.40998, 0.33140, 0.83751], [0.36931, 0.25571, 1.03037], [0.32948, 0.14652, 1.07338],
    [0.20515, 0.42084, 0.88500], [0.16043, 0.43212, 0.78800], [0.25161, 0.51454, 0.91835],
    [0.08786, 0.38957, 0.97613], [0.12354, 0.39952, 1.08026], [0.04184, 0.29078, 0.95468],
    [-0.02666, 0.49127, 0.95486], [-0.08570, 0.49284, 0.84648], [-0.04279, 0.58152, 1.05535],
    [-0.12874, 0.63733, 1.05244], [0.02321, 0.58779, 1.13410], [0.44727, 0.34114, 1.10070],
    [0.46524, 0.43042, 1.05329], [0.47527, 0.33296, 1.23664], [0.47288, 0.22937, 1.26689],
    [0.37344, 0.41380, 1.31209], [0.28033, 0.35436, 1.36486], [0.62615, 0.37332, 1.26119],
    [0.64461, 0.36635, 1.37113], [0.65096, 0.47286, 1.23279], [0.72561, 0.28525, 1.17817],
    [0.70915, 0.29087, 1.06645], [0.82657, 0.32698, 1.19769], [0.72156, 0.14054, 1.22117],
    [0.80566, 0.10314, 1.30188], [0.62960, 0.06136, 1.15968], [0.62940, -0.03945, 1.17331],
    [0.55225, 0.10583, 1.11602], [0.37286, 0.53909, 1.32316],
]  # fmt: skip


def alanine_dipeptide() -> Peptide:
    return Peptide(
        "alanine-dipeptide", list(_AD_NAMES), list(_AD_RES), list(_AD_RESID),
        np.array(_AD_XYZ, dtype=np.float64), _build_bonds(_AD_NAMES, _AD_RES, _AD_RESID),
    )  # fmt: skip


def tetrapeptide_2olx() -> Peptide:
    return Peptide(
        "2olx", list(_2OLX_NAMES), list(_2OLX_RES), list(_2OLX_RESID),
        np.array(_2OLX_XYZ, dtype=np.float64), _build_bonds(_2OLX_NAMES, _2OLX_RES, _2OLX_RESID),
    )  # fmt: skip


def get_peptide(name: str) -> Peptide:
    if name in ("alanine-dipeptide", "AD", "ad"):
        return alanine_dipeptide()
    if name in ("2olx", "4AA", "tetrapeptide"):
        return tetrapeptide_2olx()
    raise KeyError(name)

"""How the ff99SB-ILDN / OBC2 table of timewarp_b200/amber99.py was tied to the reference (authoring container only: reads
the OpenMM fixtures under /root/reference).  Not product code, not a test: a record of the identification.

A differentiable fp64 torch restatement of the energy (same functional forms as oracle/energy_oracle.py) gives forces by
autograd; the residual against the fixtures' OpenMM forces is then examined per atom and explained by linear least squares
over candidate corrections.  What it established, in order:
  1. GB radii: mbondi2 leaves 44 kJ/mol/nm rms; the per-environment set now in amber99.gb_radius leaves 17.7, all of it on
     the atoms of the two Asn side-chain torsions.
  2. Backbone amide improper: k = 1.1001 kcal/mol (side-chain NH2 stays 1.0).
  3. ILDN Asn series on C-CA-CB-CG and CA-CB-CG-ND2: six cosine harmonics each, identical for both Asn residues, phases
     exactly 0 / 180 degrees, 7th / 8th harmonics and N-CA-CB-CG / CA-CB-CG-OD1 corrections zero to 1e-4 (printed below).
  4. Solvent dielectric 78.5 and the Amber-form constants k (1 + cos) bring the absolute energy to -0.003 +- 0.004 kJ/mol.

    python tools/ff_identify.py          # prints the joint force + energy fit of step 3 and the final residuals
"""
import copy
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from timewarp_b200 import amber99 as A  # noqa: E402
from timewarp_b200.forcefield import amber99sbildn_obc2  # noqa: E402
from timewarp_b200.peptides import tetrapeptide_2olx  # noqa: E402

torch.set_default_dtype(torch.float64)
REF = "/root/reference"
FILES = ["simulation/testdata/implicit-2olx-traj-cpu-arrays.npz", "simulation/testdata/implicit-2olx-traj-arrays.npz",
         "testdata/output/2olx-traj-arrays.npz", "testdata/smallest_molecule/2olx-traj-arrays.npz"]
T = lambda a: a if torch.is_tensor(a) else torch.tensor(np.asarray(a))  # noqa: E731


def dihedral(x, i, j, k, l):
    d0, d1, d2 = x[:, i] - x[:, j], x[:, k] - x[:, j], x[:, k] - x[:, l]
    c1, c2 = torch.cross(d0, d1, dim=-1), torch.cross(d1, d2, dim=-1)
    cs = (c1 * c2).sum(-1) / torch.sqrt((c1 * c1).sum(-1) * (c2 * c2).sum(-1))
    phi = torch.acos(cs.clamp(-1 + 1e-14, 1 - 1e-14))
    return torch.where((d0 * c2).sum(-1) < 0, -phi, phi)


def energy(s, x):
    """Potential energy [B] of SystemDescription `s` at x [B,N,3]; differentiable w.r.t. x."""
    N = s.n_atoms
    i, j = T(s.bond_idx).long().T
    e = (0.5 * T(s.bond_param[:, 1]) * (torch.linalg.norm(x[:, i] - x[:, j], dim=-1) - T(s.bond_param[:, 0])) ** 2).sum(-1)
    i, j, k = T(s.angle_idx).long().T
    a, c = x[:, i] - x[:, j], x[:, k] - x[:, j]
    th = torch.acos(((a * c).sum(-1) / torch.sqrt((a * a).sum(-1) * (c * c).sum(-1))).clamp(-1, 1))
    e = e + (0.5 * T(s.angle_param[:, 1]) * (th - T(s.angle_param[:, 0])) ** 2).sum(-1)
    i, j, k, l = T(s.torsion_idx).long().T
    nper, ph, kk = T(s.torsion_param).T
    e = e + (kk * (1 + torch.cos(nper * dihedral(x, i, j, k, l) - ph))).sum(-1)
    iu = np.triu_indices(N, 1)
    iu0, iu1 = T(iu[0]).long(), T(iu[1]).long()
    r2 = ((x[:, iu0] - x[:, iu1]) ** 2).sum(-1)
    ru = r2.sqrt()
    rc, erf = s.cutoff, s.reaction_field_eps
    krf, crf = (erf - 1) / ((2 * erf + 1) * rc**3), 3 * erf / ((2 * erf + 1) * rc)
    incl = T(s.excluded[iu[0], iu[1]] == 0)[None] & (ru <= rc)
    q = T(s.charge)
    sig, eps = 0.5 * (T(s.sigma)[iu0] + T(s.sigma)[iu1]), torch.sqrt(T(s.epsilon)[iu0] * T(s.epsilon)[iu1])
    sr6 = (sig**2 / r2) ** 3
    e = e + torch.where(incl, 4 * eps * (sr6 * sr6 - sr6) + s.one_4pi_eps0 * q[iu0] * q[iu1] * (1 / ru + krf * r2 - crf), 0.0).sum(-1)
    i, j = T(s.exception_idx).long().T
    r2e = ((x[:, i] - x[:, j]) ** 2).sum(-1)
    qq, se, ee = T(s.exception_param).T
    sr6 = (se**2 / r2e) ** 3
    e = e + (4 * ee * (sr6 * sr6 - sr6) + s.one_4pi_eps0 * qq / r2e.sqrt()).sum(-1)
    eye = torch.eye(N, dtype=bool)[None]
    r2f = ((x[:, :, None] - x[:, None]) ** 2).sum(-1)
    rf = (r2f + eye * 1.0).sqrt()
    rad = T(s.gb_radius)
    orad = rad - s.gb_offset
    srj, ori = (orad * T(s.gb_scale))[None, None, :], orad[None, :, None]
    l_ij, u_ij = 1 / torch.maximum(ori.expand_as(rf), (rf - srj).abs()), 1 / (rf + srj)
    term = l_ij - u_ij + 0.25 * rf * (u_ij**2 - l_ij**2) + 0.5 / rf * torch.log(u_ij / l_ij) + 0.25 * srj**2 / rf * (l_ij**2 - u_ij**2)
    term = term + torch.where(ori < (srj - rf), 2 * (1 / ori - l_ij), 0.0)
    ssum = torch.where((ori < rf + srj) & (~eye) & (rf <= rc), term, 0.0).sum(-1) * 0.5 * orad[None]
    born = 1 / (1 / orad[None] - torch.tanh(s.gb_alpha * ssum - s.gb_beta * ssum**2 + s.gb_gamma * ssum**3) / rad[None])
    e = e + (s.surface_area_energy * (rad + 0.14)[None] ** 2 * (rad[None] / born) ** 6).sum(-1)
    pre = -s.one_4pi_eps0 * (1 / s.solute_dielectric - 1 / s.solvent_dielectric)
    a2 = born[:, :, None] * born[:, None, :]
    r2m = torch.where(eye, 0.0, r2f)
    qqm = pre * q[None, :, None] * q[None, None, :]
    g = qqm / torch.sqrt(r2m + a2 * torch.exp(-r2m / (4 * a2))) - torch.where(~eye, qqm / rc, 0.0)
    return e + 0.5 * torch.where(r2m <= rc * rc, g, 0.0).sum((-1, -2))


def load(rel):
    d = np.load(os.path.join(REF, rel))
    return torch.tensor(d["positions"].astype(np.float64)), torch.tensor(d["energies"][:, 0]), torch.tensor(d["forces"].astype(np.float64))


def system_for(pep, rel, base):
    """The 140-frame fixture (and the 2-frame one derived from it) carry the other carboxylate improper order."""
    if "output" in rel or "smallest" in rel:
        c = [i for i, (n, r) in enumerate(zip(pep.atom_names, pep.residue_index)) if n == "C" and r == max(pep.residue_index)][0]
        return amber99sbildn_obc2(pep, improper_choice={c: 0}) if base is None else base(pep, {c: 0})
    return amber99sbildn_obc2(pep) if base is None else base(pep, None)


def main():
    pep = tetrapeptide_2olx()
    saved = dict(A.ILDN)

    def without_asn_series(pep, choice):
        A.ILDN[("ASN", ("C", "CA", "CB", "CG"))] = [(0.0, 0.0, 1)]
        A.ILDN[("ASN", ("CA", "CB", "CG", "ND2"))] = [(0.0, 0.0, 1)]
        try:
            return amber99sbildn_obc2(pep, improper_choice=choice)
        finally:
            A.ILDN.update(saved)

    quads = {"C-CA-CB-CG": [(6, 4, 8, 11), (20, 18, 22, 25)], "CA-CB-CG-ND2": [(4, 8, 11, 13), (18, 22, 25, 27)]}
    sign = {"C-CA-CB-CG": [1, -1, 1, -1, 1, -1], "CA-CB-CG-ND2": [-1, -1, -1, 1, 1, -1]}  # phase 0 -> +, 180 -> -
    G, R, GE, RE = [], [], [], []
    for rel in FILES:
        X, E, F = load(rel)
        s = system_for(pep, rel, without_asn_series)
        x = X.clone().requires_grad_(True)
        e = energy(s, x)
        (g,) = torch.autograd.grad(e.sum(), x)
        cols, ecols = [], []
        for name, qs in quads.items():
            idx = torch.tensor(qs)
            for n in range(1, 7):
                x = X.clone().requires_grad_(True)
                ee = (sign[name][n - 1] + torch.cos(n * dihedral(x, idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]))).sum(-1)  # Amber form k (1 + cos)
                (gg,) = torch.autograd.grad(ee.sum(), x)
                cols.append((-gg).reshape(-1)), ecols.append(ee.detach())
        G.append(torch.stack(cols, 1)), R.append((F + g).reshape(-1)), GE.append(torch.stack(ecols, 1)), RE.append(E - e.detach())
    G, R, GE, RE = torch.cat(G), torch.cat(R), torch.cat(GE), torch.cat(RE)
    w = 10.0  # energies are ten times less noisy than the float32 forces
    t = torch.linalg.lstsq(torch.cat([G, w * GE]), torch.cat([R, w * RE])[:, None]).solution[:, 0] / A.KCAL
    print("joint fit over", len(RE), "frames: cosine coefficients in kcal/mol (negative = phase 180)")
    print("  C-CA-CB-CG  :", t[:6].numpy().round(4))
    print("  CA-CB-CG-ND2:", t[6:].numpy().round(4))
    print("table        :", {k: v for k, v in A.ILDN.items()})
    for rel in FILES:
        X, E, F = load(rel)
        s = system_for(pep, rel, None)
        x = X.clone().requires_grad_(True)
        e = energy(s, x)
        (g,) = torch.autograd.grad(e.sum(), x)
        dE = E - e.detach()
        print(f"{rel:60s} frames {len(E):3d}  dE mean {dE.mean():+.4f} std {dE.std():.4f} max {dE.abs().max():.4f} kJ/mol   "
              f"force residual rms {(F + g).pow(2).mean().sqrt():.4f} of {F.pow(2).mean().sqrt():.0f} kJ/mol/nm")


if __name__ == "__main__":
    sys.exit(main())

"""fp64 numpy oracle of the implicit-solvent Amber potential energy.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against the reference: the arithmetic lives in OpenMM 7.7 (third-party, pinned in
timewarp-environment.yml:22, absent from /root/reference and from this machine) driven by Amber
XML parameter files that are not on disk either, so the reference's golden energies
(simulation/testdata/implicit-2olx-traj-cpu-arrays.npz, checked by simulation/tests/test_md.py:35-47)
cannot be reproduced here.  This file restates OpenMM's published functional forms for the forces
that simulation/md.py:149-173 creates:

  HarmonicBondForce     1/2 k (r - r0)^2
  HarmonicAngleForce    1/2 k (theta - theta0)^2
  PeriodicTorsionForce  k (1 + cos(n phi - phase))
  NonbondedForce        4 eps ((sig/r)^12 - (sig/r)^6) + ONE_4PI_EPS0 q1 q2 (1/r + k_rf r^2 - c_rf),
                        Lorentz-Berthelot, CutoffNonPeriodic (pairs beyond the cutoff dropped),
                        k_rf = (eps_rf-1)/((2 eps_rf+1) rc^3), c_rf = 3 eps_rf/((2 eps_rf+1) rc);
                        exceptions evaluated without cutoff / reaction field
  GBSAOBCForce          Onufriev-Bashford-Case Born radii + Still pair energy + ACE surface term,
                        as in OpenMM's ReferenceObc (cutoff variant: pair term shifted by -qq/rc)

It is written independently of the CUDA kernel (vectorised numpy, different loop structure) and is
the checker for tests/ and the CPU baseline of bench.py; closed-form cases are in tests/test_energy_oracle.py.
"""
from __future__ import annotations

import numpy as np


def _dihedral(p0, p1, p2, p3):
    d0, d1, d2 = p0 - p1, p2 - p1, p2 - p3
    c1, c2 = np.cross(d0, d1), np.cross(d1, d2)
    cs = (c1 * c2).sum(-1) / np.sqrt((c1 * c1).sum(-1) * (c2 * c2).sum(-1))
    phi = np.arccos(np.clip(cs, -1.0, 1.0))
    return np.where((d0 * c2).sum(-1) < 0, -phi, phi)


def energy_terms(sysd, coords) -> np.ndarray:
    """coords [B,N,3] (nm) -> [B,5] = bond, angle, torsion, nonbonded(+exceptions), GB/SA in kJ/mol."""
    x = np.asarray(coords, dtype=np.float64)
    if x.ndim == 2:
        x = x[None]
    B, N, _ = x.shape
    out = np.zeros((B, 5))
    # bonds
    if len(sysd.bond_idx):
        i, j = sysd.bond_idx[:, 0], sysd.bond_idx[:, 1]
        r = np.linalg.norm(x[:, i] - x[:, j], axis=-1)
        out[:, 0] = (0.5 * sysd.bond_param[:, 1] * (r - sysd.bond_param[:, 0]) ** 2).sum(-1)
    # angles
    if len(sysd.angle_idx):
        i, j, k = sysd.angle_idx.T
        a, c = x[:, i] - x[:, j], x[:, k] - x[:, j]
        cs = (a * c).sum(-1) / np.sqrt((a * a).sum(-1) * (c * c).sum(-1))
        th = np.arccos(np.clip(cs, -1.0, 1.0))
        out[:, 1] = (0.5 * sysd.angle_param[:, 1] * (th - sysd.angle_param[:, 0]) ** 2).sum(-1)
    # torsions
    if len(sysd.torsion_idx):
        i, j, k, l = sysd.torsion_idx.T
        phi = _dihedral(x[:, i], x[:, j], x[:, k], x[:, l])
        n, ph, kk = sysd.torsion_param.T
        out[:, 2] = (kk * (1.0 + np.cos(n * phi - ph))).sum(-1)
    # nonbonded
    diff = x[:, :, None, :] - x[:, None, :, :]
    r2 = (diff * diff).sum(-1)
    iu = np.triu_indices(N, 1)
    r2u = r2[:, iu[0], iu[1]]
    ru = np.sqrt(r2u)
    use_cut = sysd.cutoff > 0
    rc = sysd.cutoff
    erf = sysd.reaction_field_eps
    krf = (erf - 1.0) / ((2.0 * erf + 1.0) * rc**3) if use_cut else 0.0
    crf = 3.0 * erf / ((2.0 * erf + 1.0) * rc) if use_cut else 0.0
    incl = (sysd.excluded[iu[0], iu[1]] == 0)[None, :]
    if use_cut:
        incl = incl & (ru <= rc)
    sig = 0.5 * (sysd.sigma[iu[0]] + sysd.sigma[iu[1]])
    eps = np.sqrt(sysd.epsilon[iu[0]] * sysd.epsilon[iu[1]])
    sr6 = (sig**2 / r2u) ** 3
    e_pair = 4.0 * eps * (sr6 * sr6 - sr6) + sysd.one_4pi_eps0 * sysd.charge[iu[0]] * sysd.charge[iu[1]] * (1.0 / ru + krf * r2u - crf)
    out[:, 3] = np.where(incl, e_pair, 0.0).sum(-1)
    if len(sysd.exception_idx):
        i, j = sysd.exception_idx.T
        r2e = r2[:, i, j]
        qq, s, e = sysd.exception_param.T
        sr6 = (s**2 / r2e) ** 3
        out[:, 3] += (4.0 * e * (sr6 * sr6 - sr6) + sysd.one_4pi_eps0 * qq / np.sqrt(r2e)).sum(-1)
    # GB / SA
    if sysd.use_gb:
        r = np.sqrt(r2)
        rad = sysd.gb_radius
        orad = rad - sysd.gb_offset  # [N]
        srj = (orad * sysd.gb_scale)[None, None, :]  # scaled radius of j
        ori = orad[None, :, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            rsr = r + srj
            l_ij = 1.0 / np.maximum(ori, np.abs(r - srj))
            u_ij = 1.0 / rsr
            term = (l_ij - u_ij + 0.25 * r * (u_ij**2 - l_ij**2) + 0.5 / r * np.log(u_ij / l_ij)
                    + 0.25 * srj**2 / r * (l_ij**2 - u_ij**2))
            term = term + np.where(ori < (srj - r), 2.0 * (1.0 / ori - l_ij), 0.0)
        valid = (ori < rsr) & (~np.eye(N, dtype=bool))[None]
        if use_cut:
            valid = valid & (r <= rc)
        s = np.where(valid, term, 0.0).sum(-1) * 0.5 * orad[None, :]
        th = np.tanh(sysd.gb_alpha * s - sysd.gb_beta * s**2 + sysd.gb_gamma * s**3)
        born = 1.0 / (1.0 / orad[None, :] - th / rad[None, :])  # [B,N]
        e_gb = np.zeros(B)
        if sysd.surface_area_energy != 0.0:
            sa = sysd.surface_area_energy * (rad + 0.14)[None, :] ** 2 * (rad[None, :] / born) ** 6
            e_gb += np.where(born > 0, sa, 0.0).sum(-1)
        pre = (-sysd.one_4pi_eps0 * (1.0 / sysd.solute_dielectric - 1.0 / sysd.solvent_dielectric)
               if (sysd.solute_dielectric != 0 and sysd.solvent_dielectric != 0) else 0.0)
        a2 = born[:, :, None] * born[:, None, :]
        den = np.sqrt(r2 + a2 * np.exp(-r2 / (4.0 * a2)))
        qq = pre * sysd.charge[None, :, None] * sysd.charge[None, None, :]
        g = qq / den
        offd = ~np.eye(N, dtype=bool)[None]
        if use_cut:
            g = g - np.where(offd, qq / rc, 0.0)
            g = np.where((r2 <= rc * rc), g, 0.0)
        # pairs j >= i: off-diagonal counted once, diagonal halved
        e_gb += 0.5 * np.where(offd, g, 0.0).sum((-1, -2)) + 0.5 * np.einsum("bii->b", g)
        out[:, 4] = e_gb
    return out


def potential_energy(sysd, coords) -> np.ndarray:
    return energy_terms(sysd, coords).sum(-1)

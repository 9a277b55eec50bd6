"""TEST INFRASTRUCTURE -- CPU restatement of the reference's Metropolis-Hastings drivers.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path
(timewarp_b200/) never does.

Restates, line by line, on top of the flow oracle (oracle/flow_oracle.py) and the energy oracle
(oracle/energy_oracle.py):

* `sample_with_model`   utils/evaluation_utils.py:468-745 -- ONE chain, S parallel proposals per iteration, the chain
  keeps every state up to and including the first accepted proposal (:668-713), optional adaptive S through the
  exponential moving average of the acceptance (:685-697), ChainStats arrays (:721-743).  The integrator options
  (`openmm_on_current` / `openmm_on_proposal`) and `rotate` are not restated (they need OpenMM / only type-check for
  V == 3).
* `mh_step`             the body of that loop for B independent chains with one proposal each (:589-668), the unit the
  product's `MHChains.step()` implements.
* `explore_step`        exploration.py:229-250 (threshold acceptance + chirality veto), one step of P chains.
* `compute_kinetic_energy` :416-436, `compute_num_proposal_steps` :32-64, `compute_chirality_sign` /
  `check_symmetry_change` utils/chirality.py:41-80.

Random numbers come from a `draws` object (`randn(shape)`, `rand(n)`), in exactly the order the reference consumes its
generator: initial velocities (:530-531), then per iteration the velocity resample (:590-592), the two latent draws of
`conditional_sample_with_logp` (flow.py:274-275, coordinates first) and the S uniforms (:668).  `TorchDraws(generator)`
draws on the CPU; the GPU tests pass a replay object that draws from the CUDA generator so both sides see one stream.

Parity status: pinned at the level the reference pins it (tests/test_evaluation_utils.py:112-138 checks shapes only) plus
the flow / chirality oracles underneath, which are pinned by golden vectors of the unmodified reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import energy_oracle as eo
from . import flow_oracle as fo


class TorchDraws:
    """CPU generator with the reference's draw calls."""

    def __init__(self, generator: torch.Generator):
        self.g = generator

    def randn(self, shape):
        return torch.randn(tuple(shape), generator=self.g)

    def rand(self, n):
        return torch.rand(n, generator=self.g)


def compute_num_proposal_steps(current_acceptance_probability, target_acceptance_per_step=0.9, max_num_proposal_steps=100):
    """utils/evaluation_utils.py:32-64."""
    probability_of_rejection = min(max(1 - current_acceptance_probability, 1e-3), 1 - 1e-3)
    with np.errstate(all="ignore"):
        n = np.nan_to_num(np.log(1 - target_acceptance_per_step) / np.log(probability_of_rejection), nan=np.inf)
    return max(int(np.ceil(min(n, max_num_proposal_steps))), 1)


def compute_kinetic_energy(velocs, masses, random_velocs=False, kbT=None):
    """utils/evaluation_utils.py:416-436."""
    if random_velocs:
        return 0.5 * ((velocs**2.0).sum(-1)).sum(-1)
    assert kbT, "Requires kbT to compute energy"
    return 0.5 * (masses * (velocs**2.0).sum(-1)).sum(-1) / kbT


def compute_chirality_sign(coords, chirality_centers):
    """utils/chirality.py:41-62."""
    d = coords[:, chirality_centers[:, 1:], :] - coords[:, chirality_centers[:, [0]], :]
    return torch.sign(torch.einsum("ijk,ijk->ij", d[:, :, 0], torch.cross(d[:, :, 1], d[:, :, 2], dim=-1)))


def check_symmetry_change(coords, chirality_centers, reference_signs):
    """utils/chirality.py:65-80."""
    return (compute_chirality_sign(coords, chirality_centers) != reference_signs.to(coords)).any(dim=-1)


def potential_energy_kT(sysd32, coords, kbT):
    """openmm_potential_energy_torch(coords) / kbT, squeezed (evaluation_utils.py:628,635): fp64 energy oracle, fp32 result
    like the bridge (openmm_bridge.py:292-307 returns the input dtype)."""
    e = eo.potential_energy(sysd32, coords.detach().cpu().numpy().astype(np.float64))
    return torch.from_numpy(np.asarray(e)).to(torch.float32) / kbT


@dataclass
class StepRecord:
    y_coords: torch.Tensor
    y_velocs: torch.Tensor
    p_xy: torch.Tensor
    p_yx: torch.Tensor
    e_pot_x: torch.Tensor
    e_pot_y: torch.Tensor
    e_kin_x: torch.Tensor
    e_kin_y: torch.Tensor
    exponent: torch.Tensor
    p_acc: torch.Tensor


def proposal_terms(sd, o, sysd32, kbT, atom_types, x_coords, x_velocs, mask, S, z_coords, z_velocs, masses=None, random_velocs=True,
                   chirality_centers=None, reference_signs=None, distance_mode="direct", e_pot_x=None) -> StepRecord:
    """Everything between the proposal and the acceptance probability (evaluation_utils.py:609-665) for conditioning states
    `x_coords [B,V,3]` and S proposals each (S == 1 or B == 1 like flow.py:326-331); `z_*` are the scaled latents."""
    B = x_coords.shape[0]
    yc, yv, p_xy = fo.conditional_sample_with_logp(sd, o, atom_types, x_coords, x_velocs, mask, S, z_coords, z_velocs,
                                                   distance_mode=distance_mode)  # :609-617
    if B == 1:
        yc, yv = yc.squeeze(1), yv.squeeze(1)  # :618-619
        x_rep, xv_rep = x_coords.repeat(S, 1, 1), x_velocs.repeat(S, 1, 1)  # :620-621
        at_rep, mask_rep = atom_types.repeat(S, 1), mask.repeat(S, 1)
    else:
        assert S == 1
        yc, yv = yc[0], yv[0]
        x_rep, xv_rep, at_rep, mask_rep = x_coords, x_velocs, atom_types, mask
    if e_pot_x is None:
        e_pot_x = potential_energy_kT(sysd32, x_rep, kbT)  # :628
    e_kin_x = compute_kinetic_energy(xv_rep, masses, random_velocs=random_velocs, kbT=kbT)  # :629
    e_kin_y = compute_kinetic_energy(yv, masses, random_velocs=random_velocs, kbT=kbT)  # :632
    e_pot_y = potential_energy_kT(sysd32, yc, kbT)  # :635
    if chirality_centers is not None and reference_signs is not None:  # :638-642
        e_pot_y = e_pot_y.clone()
        e_pot_y[check_symmetry_change(yc, chirality_centers, reference_signs)] += 2000
    energy = (e_pot_y - e_pot_x) + (e_kin_y - e_kin_x)  # :644-646
    sgn = 1.0 if random_velocs else -1.0
    p_yx = fo.log_likelihood(sd, o, at_rep, yc, sgn * yv, x_rep, sgn * xv_rep, mask_rep, distance_mode=distance_mode)  # :648-657
    p_xy = p_xy.reshape(p_yx.shape)  # :659
    exp = energy + p_xy - p_yx  # :663
    p_acc = torch.min(torch.tensor(1.0), torch.exp(-exp))  # :665
    return StepRecord(yc, yv, p_xy, p_yx, e_pot_x, e_pot_y, e_kin_x, e_kin_y, exp, p_acc)


def mh_step(sd, o, sysd32, kbT, atom_types, x_coords, mask, draws, masses=None, x_velocs=None, random_velocs=True, resample_velocs=True,
            chirality_centers=None, reference_signs=None, accept=True, distance_mode="direct", e_pot_x=None):
    """One iteration of the loop body (:589-713) applied to B independent chains with one proposal each.  Returns
    (new_coords, new_velocs, accepted [B] bool, u [B], StepRecord).  Draw order: velocities, latent coordinates, latent
    velocities, uniforms."""
    B, V = x_coords.shape[:2]
    if random_velocs and resample_velocs:
        x_velocs = draws.randn((B, V, 3))  # :590-592
    zc = draws.randn((1, B, V, 3)) * torch.exp(sd["coords_prior_log_scale"])  # flow.py:274-277
    zv = draws.randn((1, B, V, 3)) * torch.exp(sd["velocs_prior_log_scale"])
    rec = proposal_terms(sd, o, sysd32, kbT, atom_types, x_coords, x_velocs, mask, 1, zc, zv, masses, random_velocs, chirality_centers,
                         reference_signs, distance_mode, e_pot_x)
    u = draws.rand(B)  # :668
    acc = (u < rec.p_acc) if accept else torch.ones(B, dtype=torch.bool)
    new_c = torch.where(acc[:, None, None], rec.y_coords, x_coords)
    new_v = torch.where(acc[:, None, None], rec.y_velocs, x_velocs)
    return new_c, new_v, acc, u, rec


def sample_with_model(sd, o, sysd32, kbT, atom_types, atom_coords, atom_velocs, masked_elements, masses, num_samples, draws,
                      accept=False, random_velocs=False, resample_velocs=False, num_proposal_steps=1, adaptive_parallelism=False,
                      acceptance_rate_smoothing_factor=0.01, reference_signs=None, chirality_centers=None, distance_mode="direct",
                      trace=None, rotate=False):
    """utils/evaluation_utils.py:468-745 for one chain (batch size 1, :517).  Returns the reference's 4-tuple
    `(sampled_coords, sampled_velocs, accepted, stats)` with `stats` a dict of the ChainStats arrays (:721-731).  If `trace`
    is a list, one dict per iteration (S, u, p_acc, first_acc_idx) is appended for the lock-step comparison."""
    assert atom_coords.shape[0] == 1, "only batch-size of 1 is supported"  # :517
    keys = ("acceptance_indicator", "acceptance", "p_xy", "p_yx", "exponent", "energies_pot", "energies_kin", "energies_pot_delta",
            "energies_kin_delta")
    st = {k: [] for k in keys}
    x_coords = atom_coords.to(torch.float32).contiguous()  # :529
    V = x_coords.shape[1]
    x_velocs = draws.randn(x_coords.shape) if random_velocs else atom_velocs.to(torch.float32).contiguous()  # :530-533
    sampled_coords = [x_coords.numpy().copy()]  # :567-568
    sampled_velocs = [x_velocs.numpy().copy()]
    accepted = 0
    current_acceptance_probability = 1e-3  # :576
    max_num_proposal_steps = num_proposal_steps  # :578
    if adaptive_parallelism:  # :579-585
        num_proposal_steps = compute_num_proposal_steps(current_acceptance_probability, max_num_proposal_steps=max_num_proposal_steps)
    i = 0
    while i < num_samples:  # :589
        S = num_proposal_steps
        if random_velocs and resample_velocs:
            x_velocs = draws.randn(x_velocs.shape)  # :590-592
        if rotate:  # :604-607: Q = random_rotation_matrix() (scipy Rotation.random(), numpy's global generator), applied per atom
            from scipy.spatial.transform import Rotation as _R  # (the reference's `(Q @ x.T).T` on [1, V, 3] only type-checks for V == 3)

            Q = torch.tensor(_R.random().as_matrix(), dtype=x_coords.dtype)
            x_coords, x_velocs = x_coords @ Q.T, x_velocs @ Q.T
        zc = draws.randn((S, 1, V, 3)) * torch.exp(sd["coords_prior_log_scale"])  # flow.py:274-277
        zv = draws.randn((S, 1, V, 3)) * torch.exp(sd["velocs_prior_log_scale"])
        rec = proposal_terms(sd, o, sysd32, kbT, atom_types, x_coords, x_velocs, masked_elements, S, zc, zv, masses, random_velocs,
                             chirality_centers, reference_signs, distance_mode)
        x_rep, xv_rep = x_coords.repeat(S, 1, 1), x_velocs.repeat(S, 1, 1)  # :620-621
        y_coords, y_velocs, p_acc = rec.y_coords, rec.y_velocs, rec.p_acc
        u = None
        if accept:
            u = draws.rand(S)
            accepted_samples = u.to(p_acc) < p_acc  # :668
            acc_idx = accepted_samples.nonzero(as_tuple=True)[0]  # :669
            did_not_accept = len(acc_idx) == 0
            if did_not_accept:
                first_acc_idx = S - 1  # :671-672
            else:
                first_acc_idx = int(acc_idx[0])  # :674
                x_rep[first_acc_idx] = y_coords[first_acc_idx]  # :675-676
                xv_rep[first_acc_idx] = y_velocs[first_acc_idx]
                accepted += 1
            first_acc_idx = min(first_acc_idx, num_samples - i)  # :681 (an index bound by a count: kept as in the reference)
            st["acceptance_indicator"].append(accepted_samples[: first_acc_idx + 1].numpy())  # :683
            current_acceptance_probability = (
                acceptance_rate_smoothing_factor * (1 - did_not_accept)
                + (1 - acceptance_rate_smoothing_factor) ** first_acc_idx * current_acceptance_probability)  # :686-690
            if adaptive_parallelism:  # :691-697
                num_proposal_steps = compute_num_proposal_steps(current_acceptance_probability, max_num_proposal_steps=max_num_proposal_steps)
        elif S == 1:  # :698-705
            x_rep, xv_rep = y_coords, y_velocs
            accepted += 1
            first_acc_idx = 0
            st["acceptance_indicator"].append(np.array([True]))
        else:
            raise ValueError("Number of proposals has to be one if everything is accepted!")  # :707
        k = first_acc_idx + 1
        sampled_coords.append(x_rep[:k].numpy().copy())  # :709-710
        sampled_velocs.append(xv_rep[:k].numpy().copy())
        x_coords = x_rep[first_acc_idx].unsqueeze(0)  # :712-713
        x_velocs = xv_rep[first_acc_idx].unsqueeze(0)
        i += k  # :717-719
        st["acceptance"].append(p_acc.numpy()[:k])  # :721-728
        st["p_xy"].append(rec.p_xy.numpy()[:k])
        st["p_yx"].append(rec.p_yx.numpy()[:k])
        st["exponent"].append(rec.exponent.numpy()[:k])
        st["energies_pot"].append(rec.e_pot_y.numpy()[:k])
        st["energies_kin"].append(rec.e_kin_y.numpy()[:k])
        st["energies_pot_delta"].append((rec.e_pot_y - rec.e_pot_x).numpy()[:k])
        st["energies_kin_delta"].append((rec.e_kin_y - rec.e_kin_x).numpy()[:k])
        if trace is not None:
            trace.append(dict(S=S, u=None if u is None else u.numpy().copy(), p_acc=p_acc.numpy().copy(), first_acc_idx=first_acc_idx,
                              exponent=rec.exponent.numpy().copy()))
    stats = {k: np.concatenate(v, axis=0) for k, v in st.items()}  # :733-743
    return np.concatenate(sampled_coords, axis=0), np.concatenate(sampled_velocs, axis=0), accepted, stats


def explore_step(sd, o, sysd32, atom_types, y, y_velocs, energies, mask, draws, threshold=300.0, chirality_centers=None,
                 reference_signs=None, distance_mode="direct"):
    """exploration.py:229-250, one step of P chains: sample (:126-134), energy in kJ/mol (:239), +10000 on a chirality change
    (:240-242), keep the old state where E_new - E_old > threshold (:243-246), resample velocities (:250).  Returns
    (y, energies, accepted [P] bool, y_new, e_new, next_velocs)."""
    P, V = y.shape[:2]
    zc = draws.randn((1, P, V, 3)) * torch.exp(sd["coords_prior_log_scale"])
    zv = draws.randn((1, P, V, 3)) * torch.exp(sd["velocs_prior_log_scale"])
    y_new, _, _ = fo.conditional_sample_with_logp(sd, o, atom_types, y, y_velocs, mask, 1, zc, zv, distance_mode=distance_mode)
    y_new = y_new[0]
    e_new = torch.from_numpy(np.asarray(eo.potential_energy(sysd32, y_new.numpy().astype(np.float64)))).to(torch.float32)
    if chirality_centers is not None and reference_signs is not None:
        e_new = e_new.clone()
        e_new[check_symmetry_change(y_new, chirality_centers, reference_signs)] += 10000
    reject = (e_new - energies) > threshold
    y_out = torch.where(reject[:, None, None], y, y_new)
    e_out = torch.where(reject, energies, e_new)
    return y_out, e_out, ~reject, y_new, e_new, draws.randn((P, V, 3))

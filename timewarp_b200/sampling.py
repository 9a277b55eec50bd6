"""Sampling drivers on top of the CUDA flow + energy kernels.

* `sample_with_model`  -- the reference's MH chain (utils/evaluation_utils.py:468-745): ONE chain,
  S parallel proposals per iteration, first-accept truncation, optional adaptive S.  Same argument
  names, same return tuple `(sampled_coords, sampled_velocs, accepted, ChainStats)`.
* `MHChains`           -- B independent chains, one proposal per chain per step, everything on the
  device (no host sync in the loop; the BASELINE "1024 parallel chains" workload).
* `sample_on_batches`  -- one-step acceptance on dataset pairs (utils/evaluation_utils.py:190-353).
* `explore`            -- exploration.py:124-138,229-250: energy-threshold acceptance + chirality veto.

The OpenMM-integrator options of the reference (`openmm_on_current/proposal`, `sim`) are not
available (no OpenMM) and raise if requested.
"""
from __future__ import annotations

import pickle
from dataclasses import astuple, dataclass
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .chirality import check_symmetry_change


def compute_num_proposal_steps(current_acceptance_probability: float, target_acceptance_per_step: float = 0.9,
                               max_num_proposal_steps: int = 100) -> int:
    """utils/evaluation_utils.py:32-64."""
    p_rej = min(max(1 - current_acceptance_probability, 1e-3), 1 - 1e-3)
    with np.errstate(all="ignore"):
        steps = np.nan_to_num(np.log(1 - target_acceptance_per_step) / np.log(p_rej), nan=np.inf)
    return max(int(np.ceil(min(steps, max_num_proposal_steps))), 1)


@dataclass
class ChainStats:
    """utils/evaluation_utils.py:67-114 (same fields, same pickle round trip)."""

    acceptance_indicator: np.ndarray
    acceptance: np.ndarray
    p_xy: np.ndarray
    p_yx: np.ndarray
    exponent: np.ndarray
    energies_pot: np.ndarray
    energies_kin: np.ndarray
    energies_pot_delta: np.ndarray
    energies_kin_delta: np.ndarray

    def __len__(self):
        return len(self.acceptance)

    def __getitem__(self, key):
        return ChainStats(*map(lambda x: x[key], astuple(self)))

    def thin(self, step):
        return ChainStats(*map(lambda x: x[0 : x.shape[0] : step], astuple(self)))

    def save(self, path):
        with open(path, "wb") as f:
            pickle.dump(self, f)

    @staticmethod
    def load(path):
        with open(path, "rb") as f:
            return pickle.load(f)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def compute_kinetic_energy(velocs: Tensor, masses: Optional[Tensor], random_velocs: bool = False, kbT: Optional[float] = None) -> Tensor:
    """utils/evaluation_utils.py:416-436 -> [B]."""
    v = velocs.to(torch.float32).contiguous()
    B, V = v.shape[:2]
    out = torch.empty(B, dtype=torch.float32, device=v.device)
    if random_velocs:
        m, inv = None, 0.0
    else:
        assert kbT, "Requires kbT to compute energy"
        m, inv = masses.to(device=v.device, dtype=torch.float32).contiguous(), 1.0 / kbT
        # the kernel reads one mass per atom; the reference also broadcasts [B, V] masses (evaluation_utils.py:434)
        assert m.numel() == V, f"compute_kinetic_energy: masses must hold one entry per atom ({V}), got shape {tuple(masses.shape)}"
    _lib.check(_lib.load().tw_kinetic_energy(_lib.ptr(v), _lib.ptr(m), inv, B, V, _lib.ptr(out), _stream(v.device)), "tw_kinetic_energy")
    return out


def mh_accept(e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u, x_coords=None, x_velocs=None, y_coords=None, y_velocs=None,
              want_first=False):
    """Device-side MH decision (utils/evaluation_utils.py:663-689).  Returns (exponent, p_acc, accepted[uint8], first_idx|None)."""
    n = u.shape[0]
    dev = u.device
    V = y_coords.shape[-2] if y_coords is not None else 1
    ex = torch.empty(n, dtype=torch.float32, device=dev)
    pa = torch.empty(n, dtype=torch.float32, device=dev)
    acc = torch.empty(n, dtype=torch.uint8, device=dev)
    first = torch.empty(1, dtype=torch.int32, device=dev) if want_first else None
    for t in (x_coords, x_velocs):  # updated in place: no converted copy possible
        assert t is None or (t.is_contiguous() and t.dtype == torch.float32)
    # converted inputs are held in `keep` until the launch has been enqueued: a temporary freed right after data_ptr()
    # could be handed to the next conversion by the caching allocator (e.g. fp64 energies of fp64 coordinates)
    keep = [None if t is None else t.to(torch.float32).contiguous()
            for t in (e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u, y_coords, y_velocs)]
    for t in keep[:7]:
        assert t is not None and t.numel() == n, "mh_accept: every energy / density / uniform needs one entry per proposal"
    p = [_lib.ptr(t) for t in keep]
    _lib.check(
        _lib.load().tw_mh_accept(p[0], p[1], p[2], p[3], p[4], p[5], p[6], n, V, _lib.ptr(x_coords), _lib.ptr(x_velocs), p[7], p[8],
                                 _lib.ptr(ex), _lib.ptr(pa), _lib.ptr(acc), _lib.ptr(first), _stream(dev)),
        "tw_mh_accept",
    )
    del keep
    return ex, pa, acc, first


class MHChains:
    """B independent Metropolis-Hastings chains advanced in lock-step on one GPU.

    One `step()` = the body of the reference loop (utils/evaluation_utils.py:589-665) with
    num_proposal_steps == 1 applied to every chain: resample velocities, propose through the flow
    (reverse pass), evaluate potential/kinetic energies and the reverse-move density (forward
    pass), accept/reject.  The potential energy of the current state is carried between steps
    (the reference re-evaluates it every iteration, :628 -- same value).  Nothing in `step()`
    synchronises with the host."""

    def __init__(self, model, energy, atom_types: Tensor, masked_elements: Tensor, x_coords: Tensor, x_velocs: Optional[Tensor] = None,
                 masses: Optional[Tensor] = None, random_velocs: bool = True, resample_velocs: bool = True,
                 chirality_centers: Optional[Tensor] = None, reference_signs: Optional[Tensor] = None, accept: bool = True,
                 sim=None, num_openmm_steps: int = 0, openmm_on_current: bool = False, openmm_on_proposal: bool = False):
        assert x_coords.device.type == "cuda", "MHChains runs on CUDA only"
        self.model, self.energy = model, energy
        self.B, self.V = x_coords.shape[:2]
        self.atom_types = atom_types.contiguous()
        self.mask = masked_elements.contiguous()
        self.x = x_coords.to(torch.float32).clone().contiguous()
        self.random_velocs, self.resample_velocs, self.accept = random_velocs, resample_velocs, accept
        self.xv = torch.randn_like(self.x) if (random_velocs or x_velocs is None) else x_velocs.to(torch.float32).clone().contiguous()
        self.masses = masses
        self.kbT = float(energy.kbT)
        self.centers, self.ref_signs = chirality_centers, reference_signs
        self.e_pot_x = (energy(self.x) / self.kbT).squeeze(-1).contiguous()
        dev = self.x.device
        self.n_accepted = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.n_steps = 0
        self.last = {}
        self._graph, self._graph_last = None, None
        self._empty_adj = torch.zeros(0, 2, dtype=torch.long, device=dev)
        self._empty_ebi = torch.zeros(0, dtype=torch.long, device=dev)
        # integrator steps inside the chain (evaluation_utils.py:594-602,623-626): `sim` is a timewarp_b200.md.Simulation; every
        # chain takes its `num_openmm_steps` steps in the same kernel launch
        self.sim, self.num_openmm_steps = sim, int(num_openmm_steps)
        self.openmm_on_current = bool(openmm_on_current and sim is not None and num_openmm_steps > 0)
        self.openmm_on_proposal = bool(openmm_on_proposal and sim is not None and num_openmm_steps > 0)
        if self.openmm_on_current or self.openmm_on_proposal:
            assert masses is not None, "integrator steps need the masses (velocity scale sqrt(kT / m), evaluation_utils.py:556)"
            self._velocs_std = torch.sqrt(self.kbT / masses.to(dev, torch.float32))[None, :, None]

    @torch.no_grad()  # the reference samples under no_grad (evaluation_utils.py:468)
    def step(self):
        """One MH iteration of every chain.  After `capture_graph()` the iteration is replayed as ONE CUDA graph
        (≈210 kernel launches, no host work between them); the state tensors keep their addresses either way."""
        if self._graph is not None:
            self._graph.replay()
            self.last = self._graph_last  # the graph's static output tensors
        else:
            self._step_impl()
        self.n_steps += 1
        return self.last["accepted"]

    @torch.no_grad()
    def _step_impl(self):
        m = self.model
        if self.random_velocs and self.resample_velocs:
            self.xv.normal_()  # :590-592 (same draw as torch.randn_like, in place: the state keeps its address)
        if self.openmm_on_current:  # :594-602
            if self.random_velocs:
                xn, _ = self.sim.step(self.x, self.xv * self._velocs_std, self.num_openmm_steps)
            else:
                xn, vn = self.sim.step(self.x, self.xv, self.num_openmm_steps)
                self.xv.copy_(vn)
            self.x.copy_(xn)
            self.e_pot_x.copy_((self.energy(self.x) / self.kbT).squeeze(-1))  # the carried energy belongs to the moved state
        y, yv, p_xy = m.conditional_sample_with_logp(
            atom_types=self.atom_types, x_coords=self.x, x_velocs=self.xv, adj_list=self._empty_adj,
            edge_batch_idx=self._empty_ebi, masked_elements=self.mask, num_samples=1)  # :609-617
        y, yv, p_xy = y[0], yv[0], p_xy[0]
        if self.openmm_on_proposal:  # :623-626
            y, _ = self.sim.step(y, yv * self._velocs_std, self.num_openmm_steps)
        e_kin_x = compute_kinetic_energy(self.xv, self.masses, self.random_velocs, self.kbT)  # :629
        e_kin_y = compute_kinetic_energy(yv, self.masses, self.random_velocs, self.kbT)  # :632
        e_pot_y = (self.energy(y) / self.kbT).squeeze(-1)  # :635
        if self.centers is not None and self.ref_signs is not None:
            e_pot_y = e_pot_y + 2000.0 * check_symmetry_change(y, self.centers, self.ref_signs)  # :638-642
        sgn = 1.0 if self.random_velocs else -1.0  # :651-653
        p_yx = m.log_likelihood(
            atom_types=self.atom_types, y_coords=self.x, y_velocs=sgn * self.xv, x_coords=y, x_velocs=sgn * yv,
            adj_list=self._empty_adj, edge_batch_idx=self._empty_ebi, masked_elements=self.mask)  # :648-657
        u = torch.rand(self.B, device=self.x.device)  # :668
        if self.accept:
            ex, pa, acc, _ = mh_accept(self.e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u, self.x, self.xv, y, yv)
            accb = acc.view(torch.bool)
            self.e_pot_x.copy_(torch.where(accb, e_pot_y, self.e_pot_x))
        else:  # accept everything (:698-705)
            ex, pa, acc, _ = mh_accept(self.e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u)
            self.x.copy_(y), self.xv.copy_(yv), self.e_pot_x.copy_(e_pot_y)
            accb = torch.ones_like(acc, dtype=torch.bool)
        self.n_accepted += accb
        self.last = dict(exponent=ex, acceptance=pa, accepted=accb, p_xy=p_xy, p_yx=p_yx, e_pot_y=e_pot_y, e_kin_y=e_kin_y)
        return accb

    @torch.no_grad()
    def capture_graph(self, warmup: int = 2):
        """Capture one iteration into a CUDA graph (torch.cuda.graph: the torch RNG draws stay graph-safe, outputs live
        in the graph's private pool).  `warmup` eager iterations run first so that every lazy initialisation (weight
        packing, workspace, kernel attributes) happens outside the capture.  Returns self."""
        if self.openmm_on_current or self.openmm_on_proposal:
            # the integrator's Philox offset advances on the host between calls: a replayed graph would reuse one noise stream
            raise NotImplementedError("capture_graph is not available with integrator steps inside the chain")
        for _ in range(max(int(warmup), 1)):
            self.step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._step_impl()
        self._graph, self._graph_last = graph, self.last
        return self

    def release_graph(self):
        self._graph = None

    def acceptance_rate(self) -> Tensor:
        return self.n_accepted.to(torch.float32) / max(self.n_steps, 1)


@torch.no_grad()
def sample_with_model(batch, model, device, openmm_potential_energy_torch, masses: Tensor, num_samples: int, accept: bool = False,
                      random_velocs: bool = False, resample_velocs: bool = False, initialize_randomly: bool = False,
                      num_openmm_steps: int = 0, sim=None, openmm_on_proposal: bool = False, openmm_on_current: bool = False,
                      num_proposal_steps: int = 1, adaptive_parallelism: bool = False, acceptance_rate_smoothing_factor: float = 0.01,
                      rotate: bool = False, reference_signs: Optional[Tensor] = None, chirality_centers: Optional[Tensor] = None,
                      disable_tqdm: Optional[bool] = True):
    """The reference MH sampler, utils/evaluation_utils.py:468-745 (line numbers cited inline).
    One 4-byte device->host read per iteration (the first accepted index) is inherent to the
    reference's variable-length bookkeeping; everything else stays on the device."""
    assert batch.atom_coords.size(0) == 1, "only batch-size of 1 is supported"  # :517
    from .md import openmm_step  # `sim` is a timewarp_b200.md.Simulation (integrator steps run in the tw_langevin_steps kernel)
    energy = openmm_potential_energy_torch
    x_coords = batch.atom_coords.to(device).to(torch.float32).contiguous()
    x_velocs = torch.randn_like(x_coords) if random_velocs else batch.atom_velocs.to(device).to(torch.float32).contiguous()  # :530-533
    masked_elements = batch.masked_elements.to(device)
    adj_list = batch.adj_list.to(device)
    edge_batch_idx = batch.edge_batch_idx.to(device)
    atom_types = batch.atom_types.to(device)
    masses = masses.to(device) if masses is not None else None
    if initialize_randomly:  # :540-553
        x_coords, x_velocs = model.conditional_sample(
            atom_types=atom_types, x_coords=torch.randn_like(x_coords), x_velocs=torch.randn_like(x_velocs), adj_list=adj_list,
            edge_batch_idx=edge_batch_idx, masked_elements=masked_elements, num_samples=1)
        x_coords, x_velocs = x_coords.squeeze(0), x_velocs.squeeze(0)
    kbT = energy.kbT  # :555
    md_steps = num_openmm_steps > 0 and sim is not None
    if md_steps and (openmm_on_current or openmm_on_proposal):
        velocs_std = torch.sqrt(kbT / masses)[None, :, None]  # :556
    if openmm_on_current and md_steps:  # :559-565
        if random_velocs:
            x_coords, _ = openmm_step(sim, x_coords, x_velocs * velocs_std, num_steps=num_openmm_steps)
        else:
            x_coords, x_velocs = openmm_step(sim, x_coords, x_velocs, num_steps=num_openmm_steps)
    sampled_coords = [x_coords.cpu().numpy()]
    sampled_velocs = [x_velocs.cpu().numpy()]
    accepted = 0
    current_acceptance_probability = 1e-3  # :574
    max_num_proposal_steps = num_proposal_steps
    if adaptive_parallelism:
        num_proposal_steps = compute_num_proposal_steps(current_acceptance_probability, max_num_proposal_steps=max_num_proposal_steps)
    stats = {k: [] for k in ("acceptance_indicator", "acceptance", "p_xy", "p_yx", "exponent", "energies_pot", "energies_kin",
                             "energies_pot_delta", "energies_kin_delta")}
    i = 0
    while i < num_samples:  # :589
        S = num_proposal_steps
        if random_velocs and resample_velocs:
            x_velocs = torch.randn_like(x_velocs)  # :590-592
        if openmm_on_current and md_steps:  # :594-602
            if random_velocs:
                x_coords, _ = openmm_step(sim, x_coords, x_velocs * velocs_std, num_steps=num_openmm_steps)
            else:
                x_coords, x_velocs = openmm_step(sim, x_coords, x_velocs, num_steps=num_openmm_steps)
        if rotate:  # :604-607: one Haar-random rotation of the current state per iteration (drawn like the reference draws it:
            # scipy's Rotation.random(), numpy's global generator).  The reference writes `(Q @ x_coords.T).T` on the [1, V, 3]
            # tensors, which only type-checks for V == 3; this is the per-atom rotation that line intends.
            from .equivariance import random_rotation_matrix

            Q = random_rotation_matrix(device=device, dtype=x_coords.dtype)
            x_coords = (x_coords @ Q.T).contiguous()
            x_velocs = (x_velocs @ Q.T).contiguous()
        y_coords, y_velocs, p_xy = model.conditional_sample_with_logp(
            atom_types=atom_types, x_coords=x_coords, x_velocs=x_velocs, adj_list=adj_list, edge_batch_idx=edge_batch_idx,
            masked_elements=masked_elements, num_samples=S)  # :609-617
        y_coords, y_velocs = y_coords.squeeze(1).contiguous(), y_velocs.squeeze(1).contiguous()
        x_rep, xv_rep = x_coords.repeat(S, 1, 1), x_velocs.repeat(S, 1, 1)  # :620-621
        if openmm_on_proposal and md_steps:  # :623-626 (the reference's openmm_step handles S == 1 only; here every proposal steps)
            y_coords, _ = openmm_step(sim, y_coords, y_velocs * velocs_std, num_steps=num_openmm_steps)
        e_pot_x = (energy(x_coords) / kbT).squeeze(-1).repeat(S)  # :628 (S identical evaluations in the reference)
        e_kin_x = compute_kinetic_energy(xv_rep, masses, random_velocs=random_velocs, kbT=kbT)
        e_kin_y = compute_kinetic_energy(y_velocs, masses, random_velocs=random_velocs, kbT=kbT)
        e_pot_y = (energy(y_coords) / kbT).squeeze(-1)  # :635
        if chirality_centers is not None and reference_signs is not None:
            e_pot_y = e_pot_y + 2000.0 * check_symmetry_change(y_coords, chirality_centers, reference_signs)  # :638-642
        sgn = 1.0 if random_velocs else -1.0
        p_yx = model.log_likelihood(
            atom_types=atom_types.repeat(S, 1), y_coords=x_rep, y_velocs=sgn * xv_rep, x_coords=y_coords, x_velocs=sgn * y_velocs,
            adj_list=adj_list, edge_batch_idx=edge_batch_idx, masked_elements=masked_elements.repeat(S, 1))  # :648-657
        p_xy = p_xy.reshape(p_yx.shape)
        if accept:
            u = torch.rand(S, device=device)  # :668
            exp_, p_acc, acc, first = mh_accept(e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u, want_first=True)
            first_idx = int(first.item())  # the reference's .cpu() at :674
            did_not_accept = first_idx < 0
            if did_not_accept:
                first_acc_idx = S - 1  # :671-672
            else:
                first_acc_idx = first_idx
                x_rep[first_acc_idx] = y_coords[first_acc_idx]  # :675-676
                xv_rep[first_acc_idx] = y_velocs[first_acc_idx]
                accepted += 1
            first_acc_idx = min(first_acc_idx, num_samples - i)  # :681
            stats["acceptance_indicator"].append(acc[: first_acc_idx + 1].view(torch.bool).cpu().numpy())
            current_acceptance_probability = (
                acceptance_rate_smoothing_factor * (1 - did_not_accept)
                + (1 - acceptance_rate_smoothing_factor) ** first_acc_idx * current_acceptance_probability)  # :686-690
            if adaptive_parallelism:
                num_proposal_steps = compute_num_proposal_steps(current_acceptance_probability, max_num_proposal_steps=max_num_proposal_steps)
        elif S == 1:  # :698-705
            u = torch.zeros(S, device=device)
            exp_, p_acc, acc, _ = mh_accept(e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u)
            x_rep, xv_rep = y_coords, y_velocs
            accepted += 1
            first_acc_idx = 0
            stats["acceptance_indicator"].append(np.array([True]))
        else:
            raise ValueError("Number of proposals has to be one if everything is accepted!")  # :707
        sampled_coords.append(x_rep[: first_acc_idx + 1].cpu().numpy())  # :709-710
        sampled_velocs.append(xv_rep[: first_acc_idx + 1].cpu().numpy())
        x_coords = x_rep[first_acc_idx].unsqueeze(0).contiguous()  # :712-713
        x_velocs = xv_rep[first_acc_idx].unsqueeze(0).contiguous()
        i += first_acc_idx + 1
        k = first_acc_idx + 1
        stats["acceptance"].append(p_acc.cpu().numpy()[:k])
        stats["p_xy"].append(p_xy.cpu().numpy()[:k])
        stats["p_yx"].append(p_yx.cpu().numpy()[:k])
        stats["exponent"].append(exp_.cpu().numpy()[:k])
        stats["energies_pot"].append(e_pot_y.cpu().numpy()[:k])
        stats["energies_kin"].append(e_kin_y.cpu().numpy()[:k])
        stats["energies_pot_delta"].append((e_pot_y - e_pot_x).cpu().numpy()[:k])
        stats["energies_kin_delta"].append((e_kin_y - e_kin_x).cpu().numpy()[:k])
    chain_stats = ChainStats(**{k: np.concatenate(v, axis=0) for k, v in stats.items()})
    return np.concatenate(sampled_coords, axis=0), np.concatenate(sampled_velocs, axis=0), accepted, chain_stats


@torch.no_grad()
def sample_on_batches(batches, model, device, openmm_potential_energy_torch, data_augmentation, masses, random_velocs=False):
    """One-step acceptance statistics on dataset pairs -- utils/evaluation_utils.py:190-353 (called from
    evaluate.py:355-375).  Per batch: 1x conditional_sample(S=1) + 4x log_likelihood (p_xy, p_yx of the model's own
    sample, and both on the training target, :241-311) + the energies of x and y (:267-269).  Same argument names,
    same draws from the device generator in the same order, same 11-tuple of numpy arrays.  `batch` needs the
    DenseMolDynBatch attributes used at :229-252 (atom_types, atom_coords, atom_velocs, atom_coord_targets,
    atom_veloc_targets, adj_list, edge_batch_idx, masked_elements)."""
    energy = openmm_potential_energy_torch
    masses = masses.to(device) if masses is not None else None
    cols = {k: [] for k in ("y_c", "y_v", "t_c", "t_v", "c_c", "c_v", "acc", "p_xy", "p_yx", "p_xy_tr", "p_yx_tr", "e_pot", "e_kin")}
    for batch in batches:
        if data_augmentation:  # :227-229: one random translation + rotation per batch
            from .dataloader import DenseMolDynBatch
            from .equivariance import transform_batch

            assert isinstance(batch, DenseMolDynBatch)
            batch = transform_batch(batch)
        x_coords = batch.atom_coords.to(device).to(torch.float32).contiguous()  # :229
        y_coord_targets = batch.atom_coord_targets.to(device).to(torch.float32).contiguous()
        if random_velocs:  # :232-234
            x_velocs = torch.randn_like(x_coords)
            y_veloc_targets = torch.randn_like(y_coord_targets)
        else:
            x_velocs = batch.atom_velocs.to(device).to(torch.float32).contiguous()
            y_veloc_targets = batch.atom_veloc_targets.to(device).to(torch.float32).contiguous()
        kw = dict(atom_types=batch.atom_types.to(device), adj_list=batch.adj_list.to(device),
                  edge_batch_idx=batch.edge_batch_idx.to(device), masked_elements=batch.masked_elements.to(device))
        y_coords, y_velocs = model.conditional_sample(x_coords=x_coords, x_velocs=x_velocs, num_samples=1, **kw)  # :239-247
        y_coords, y_velocs = y_coords.squeeze(0), y_velocs.squeeze(0)
        p_xy = model.log_likelihood(x_coords=x_coords, x_velocs=x_velocs, y_coords=y_coords, y_velocs=y_velocs, **kw)  # :250-259
        kbT = energy.kbT
        e_kin = (compute_kinetic_energy(y_velocs, masses, random_velocs=random_velocs, kbT=kbT)
                 - compute_kinetic_energy(x_velocs, masses, random_velocs=random_velocs, kbT=kbT))  # :262-264
        e_pot = ((energy(y_coords) - energy(x_coords)) / kbT).view(-1)  # :265-268
        sgn = 1.0 if random_velocs else -1.0
        p_yx = model.log_likelihood(y_coords=x_coords, y_velocs=sgn * x_velocs, x_coords=y_coords, x_velocs=sgn * y_velocs, **kw)  # :272-281
        exp_ = e_pot + e_kin + p_xy - p_yx  # :285
        p_acc = torch.clamp(torch.exp(-exp_), max=1.0)  # :286  (NaN exponent -> NaN, like torch.min)
        p_acc = torch.where(torch.isnan(exp_), exp_, p_acc)
        p_xy_tr = model.log_likelihood(x_coords=x_coords, x_velocs=x_velocs, y_coords=y_coord_targets, y_velocs=y_veloc_targets, **kw)  # :289-298
        p_yx_tr = model.log_likelihood(x_coords=y_coord_targets, x_velocs=sgn * y_veloc_targets, y_coords=x_coords,
                                       y_velocs=sgn * x_velocs, **kw)  # :300-309
        for k, t in (("acc", p_acc), ("p_xy", p_xy), ("p_yx", p_yx), ("p_xy_tr", p_xy_tr), ("p_yx_tr", p_yx_tr), ("e_pot", e_pot),
                     ("e_kin", e_kin), ("y_c", y_coords), ("y_v", y_velocs), ("c_c", x_coords), ("c_v", x_velocs)):
            cols[k].append(t.cpu().numpy())
        cols["t_c"].append(batch.atom_coord_targets.cpu().numpy())
        cols["t_v"].append(batch.atom_veloc_targets.cpu().numpy())
    arr = {k: np.array(v) for k, v in cols.items()}
    sq = lambda a: a.squeeze(1)  # noqa: E731  (:341-346: batches of one datapoint)
    return (sq(arr["y_c"]), sq(arr["y_v"]), sq(arr["t_c"]), sq(arr["t_v"]), sq(arr["c_c"]), sq(arr["c_v"]), arr["p_yx"], arr["p_xy"],
            arr["p_yx_tr"], arr["p_xy_tr"], arr["acc"])


@torch.no_grad()
def explore(model, energy, atom_types: Tensor, masked_elements: Tensor, x_coords: Tensor, x_velocs: Tensor, num_steps: int,
            num_chains: int, threshold: float = 300.0, chirality_centers: Optional[Tensor] = None,
            reference_signs: Optional[Tensor] = None, keep_trajectory: bool = True):
    """exploration.py:229-250: `num_chains` copies of one start state; per step one flow sample and
    one energy per chain, reject where E_new - E_old > threshold (kJ/mol) or chirality flipped (+10000).
    Returns (positions [steps*chains,V,3] or final [chains,V,3], energies, accept_counts[chains])."""
    dev = x_coords.device
    P = num_chains
    empty_adj = torch.zeros(0, 2, dtype=torch.long, device=dev)
    empty_ebi = torch.zeros(0, dtype=torch.long, device=dev)
    at = atom_types.repeat(P, 1).contiguous() if atom_types.shape[0] == 1 else atom_types
    mask = masked_elements.repeat(P, 1).contiguous() if masked_elements.shape[0] == 1 else masked_elements
    energies = energy(x_coords).repeat(P, 1).squeeze(-1).contiguous() if x_coords.shape[0] == 1 else energy(x_coords).squeeze(-1).contiguous()
    y = (x_coords.repeat(P, 1, 1) if x_coords.shape[0] == 1 else x_coords).to(torch.float32).clone().contiguous()
    yv = (x_velocs.repeat(P, 1, 1) if x_velocs.shape[0] == 1 else x_velocs).to(torch.float32).clone().contiguous()
    V = y.shape[1]
    traj, etraj = [], []
    n_acc = torch.zeros(P, dtype=torch.int64, device=dev)
    lib = _lib.load()
    for _ in range(num_steps):
        y_new, _v = model.conditional_sample(atom_types=at, x_coords=y, x_velocs=yv, adj_list=empty_adj, edge_batch_idx=empty_ebi,
                                             masked_elements=mask, num_samples=1)  # :126-134
        y_new = y_new.squeeze(0).contiguous()
        e_new = energy(y_new).squeeze(-1)  # :239
        if chirality_centers is not None and reference_signs is not None:
            e_new = e_new + 10000.0 * check_symmetry_change(y_new, chirality_centers, reference_signs)  # :240-242
        e_new = e_new.to(torch.float32).contiguous()
        acc = torch.empty(P, dtype=torch.uint8, device=dev)
        _lib.check(lib.tw_threshold_accept(_lib.ptr(y), _lib.ptr(energies), _lib.ptr(y_new), _lib.ptr(e_new), float(threshold), P, V,
                                           _lib.ptr(acc), _stream(dev)), "tw_threshold_accept")  # :243-246
        n_acc += acc
        if keep_trajectory:
            traj.append(y.clone())
            etraj.append(energies.clone())
        yv = torch.randn_like(y)  # :250
    if keep_trajectory:
        return torch.cat(traj, 0), torch.cat(etraj, 0), n_acc
    return y, energies, n_acc


def sample_trajectory(batch, model, device, openmm_potential_energy_torch, masses: Tensor, output_dir: str, protein: str,
                      num_samples: int, saving_interval: int, mh: bool = True, random_velocities: bool = True,
                      resample_velocities: bool = True, initialize_randomly: bool = False, sim=None, openmm_on_current: bool = False,
                      openmm_on_proposal: bool = False, num_openmm_steps: int = 0, num_proposal_steps: int = 1,
                      adaptive_parallelism: bool = False, conserve_chirality: bool = False, thin: int = 10) -> int:
    """The chunked, resumable sampling loop of sample_trajectory.py:217-281: `num_samples // saving_interval` calls of
    sample_with_model, each written to `{output_dir}/{protein}_trajectory_model_{i}.npz` (positions[::10], wall time) -- the
    files the paper notebooks read; a chain whose directory already holds chunks resumes from the last saved position."""
    import os
    from timeit import default_timer as timer

    from .chirality import compute_chirality_sign, find_chirality_centers

    num_iters = num_samples // saving_interval
    assert num_iters > 0, "num_samples must be larger than saving_interval."  # :221
    os.makedirs(output_dir, exist_ok=True)
    chirality_centers = reference_signs = None
    if conserve_chirality:  # :228-232
        chirality_centers = find_chirality_centers(batch.adj_list, batch.atom_types).to(device)
        reference_signs = compute_chirality_sign(batch.atom_coords.to(device), chirality_centers)
    try:  # :235-241
        n_saved_iterations = len(os.listdir(output_dir))
        npz = np.load(os.path.join(output_dir, f"{protein}_trajectory_model_{n_saved_iterations - 1}.npz"))
        batch.atom_coords = torch.from_numpy(npz["positions"][-1:])
    except FileNotFoundError:
        n_saved_iterations = 0
    needs_sim = openmm_on_proposal or openmm_on_current
    for i in range(n_saved_iterations, num_iters):
        start = timer()
        sampled_coords, _, _, _ = sample_with_model(
            batch, model, device, openmm_potential_energy_torch, masses, saving_interval, mh, random_velocs=random_velocities,
            resample_velocs=resample_velocities, initialize_randomly=initialize_randomly, sim=sim if needs_sim else None,
            openmm_on_current=openmm_on_current, openmm_on_proposal=openmm_on_proposal, num_openmm_steps=num_openmm_steps,
            num_proposal_steps=num_proposal_steps, adaptive_parallelism=adaptive_parallelism,
            reference_signs=reference_signs, chirality_centers=chirality_centers, disable_tqdm=True)
        duration = timer() - start
        np.savez(os.path.join(output_dir, f"{protein}_trajectory_model_{i}.npz"), positions=sampled_coords[::thin], time=duration)  # :270-278
        batch.atom_coords = torch.from_numpy(sampled_coords[-1:])  # :279
    return 0

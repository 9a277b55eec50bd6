"""Energy-based training losses on the GPU path: losses.py:24-149 (energies) and :558-664
(`EnergyLoss`).  The loss differentiates THROUGH the sampler: `conditional_sample_with_logp` runs the taped sampling pass
(tw_flow_sample_train / tw_flow_sample_backward) and the potential energy returns -force as its gradient
(tw_peptide_energy), so no OpenMM call and no eager PyTorch network evaluation sits in the training step.

`AcceptanceLoss` (losses.py:358-555) additionally differentiates the reverse-move density w.r.t. its CONDITIONING state (the
proposal): tw_flow_log_likelihood_backward_inputs returns that gradient (conditioner inputs + attention scores + centring)."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor


def compute_kinetic_energy(velocs: Tensor, masses: Optional[Tensor], kbT: Optional[float], random_velocs: bool = False) -> Tensor:
    """losses.py:24-45 (differentiable torch expression; the MH drivers use the tw_kinetic_energy kernel instead)."""
    if random_velocs:
        return 0.5 * ((velocs**2.0).sum(-1)).sum(-1)
    return 0.5 * (masses * (velocs**2.0).sum(-1)).sum(-1) / kbT


class EnergyProvider:
    """The part of `OpenMMProvider` (utils/openmm/openmm_provider.py:134-140) the losses use: an energy module and the
    masses per protein name, one temperature."""

    def __init__(self, energies: Dict[str, Callable[[Tensor], Tensor]], masses: Optional[Dict[str, Tensor]] = None):
        self._energies, self._masses = dict(energies), dict(masses or {})
        kbTs = {float(e.kbT) for e in self._energies.values()}
        assert len(kbTs) == 1, "all systems of a provider share one temperature"
        self.kbT = kbTs.pop()

    def get_potential_energy_module(self, name: str):
        return self._energies[name]

    def get_masses(self, name: str) -> Tensor:
        return self._masses[name]


def compute_potential_energy(coords: Tensor, pdb_names: Sequence[str], masked_elements: Tensor, provider: EnergyProvider,
                             segments: Optional[Sequence[int]] = None) -> Tensor:
    """losses.py:48-99: U / kT per sample [B]; padding atoms are dropped before the energy call; contiguous segments of one
    protein are evaluated as one batch."""
    def single(coord, protein, mask):
        pot = provider.get_potential_energy_module(protein)
        return pot(coord[~mask, :].view(coord.size(0), -1, 3)).squeeze(-1) / provider.kbT

    if segments is not None:
        parts = [single(coords[segments[i]:segments[i + 1]], pdb_names[segments[i]], masked_elements[segments[i]:segments[i + 1]])
                 for i in range(len(segments) - 1)]
    else:
        parts = [single(c[None], n, m[None]) for n, c, m in zip(pdb_names, coords, masked_elements)]
    return torch.hstack(parts)


def compute_energy(coords, velocs, pdb_names, masked_elements, provider: EnergyProvider, random_velocs: bool = False,
                   masses: Optional[Tensor] = None, segments: Optional[Sequence[int]] = None) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """losses.py:101-149: (kinetic + potential) / kT and its two parts, [B] each."""
    if masses is None and not random_velocs:
        ms = [provider.get_masses(n).to(coords.device) for n in pdb_names]
        width = masked_elements.size(-1)
        masses = torch.stack([torch.nn.functional.pad(m, (0, width - m.size(0)), "constant", 0) for m in ms])
    kinetic = compute_kinetic_energy(velocs, masses, provider.kbT, random_velocs=random_velocs)
    potential = compute_potential_energy(coords, pdb_names, masked_elements, provider, segments=segments)
    return kinetic + potential, (potential, kinetic)


class EnergyLoss:
    """losses.py:558-583: configuration of the energy loss."""

    def __init__(self, openmm_provider: EnergyProvider, random_velocs: bool = True, num_samples: int = 1):
        self.openmm_provider, self.random_velocs, self.num_samples = openmm_provider, random_velocs, num_samples


def energy_loss(loss: EnergyLoss, model, batch, device: Optional[Union[str, torch.device]] = None, logger=None) -> Tensor:
    """`get_loss(EnergyLoss, ...)`, losses.py:586-664: mean over the batch of (E(y)/kT + log p(y|x)) / n_atoms for
    y ~ p(.|x), averaged over `num_samples` draws.  `batch` carries atom_types / atom_coords / atom_velocs / adj_list /
    edge_batch_idx / masked_elements / names / segments like `DenseMolDynBatch`."""
    to = (lambda t: t.to(device, non_blocking=True)) if device is not None else (lambda t: t)
    x_coords, atom_types, masked_elements = to(batch.atom_coords), to(batch.atom_types), to(batch.masked_elements)
    adj_list, edge_batch_idx = to(batch.adj_list), to(batch.edge_batch_idx)
    x_velocs = torch.randn_like(x_coords).contiguous() if loss.random_velocs else to(batch.atom_velocs)
    num_atoms = (~masked_elements).sum(dim=-1)
    total = torch.tensor(0.0, device=x_coords.device)
    for _ in range(loss.num_samples):
        y_coords, y_velocs, logp_xy = model.conditional_sample_with_logp(
            atom_types=atom_types, x_coords=x_coords, x_velocs=x_velocs, adj_list=adj_list, edge_batch_idx=edge_batch_idx,
            masked_elements=masked_elements, num_samples=1, logger=logger)
        y_coords, y_velocs = y_coords.squeeze(0), y_velocs.squeeze(0)
        energy, _ = compute_energy(y_coords, y_velocs, batch.names, masked_elements, loss.openmm_provider,
                                   random_velocs=loss.random_velocs, segments=getattr(batch, "segments", None))
        total = total + ((energy + logp_xy.reshape(energy.shape)) / num_atoms).mean()
    return total / loss.num_samples


class AcceptanceLoss:
    """losses.py:358-393: configuration of the acceptance loss.  `chirality_checker(batch, y_coords, masked_elements) -> bool[B]`
    replaces the reference's `CiralityChecker(openmm_provider.pdb_dirs)` (it reads PDB files; here any callable, e.g. built on
    timewarp_b200.chirality.check_symmetry_change); required when `high_energy_threshold != -1`."""

    def __init__(self, openmm_provider: EnergyProvider, random_velocs: bool = True, beta: float = 0.0, clamp: bool = False,
                 num_samples: int = 1, high_energy_threshold: float = -1, chirality_checker: Optional[Callable] = None):
        self.openmm_provider, self.random_velocs, self.beta, self.clamp = openmm_provider, random_velocs, beta, clamp
        self.num_samples, self.high_energy_threshold, self.chirality_checker = num_samples, high_energy_threshold, chirality_checker
        if high_energy_threshold != -1 and chirality_checker is None:
            raise ValueError("high_energy_threshold needs a chirality_checker (losses.py:391-393)")


def acceptance_loss(loss: AcceptanceLoss, model, batch, device: Optional[Union[str, torch.device]] = None, logger=None) -> Tensor:
    """`get_loss(AcceptanceLoss, ...)`, losses.py:396-555: the negative log MH acceptance of a proposal y ~ p(.|x),
        (E(y) - E(x)) / kT + log p(y|x) - log p(x|y)      [+ beta * log p(y|x)],
    per atom, averaged over the batch and `num_samples` draws; optionally clamped at 0 (= min(1, acceptance)) and with
    high-energy / chirality-flipping proposals dropped."""
    to = (lambda t: t.to(device, non_blocking=True)) if device is not None else (lambda t: t)
    x_coords, atom_types, masked_elements = to(batch.atom_coords), to(batch.atom_types), to(batch.masked_elements)
    adj_list, edge_batch_idx = to(batch.adj_list), to(batch.edge_batch_idx)
    random_velocs = loss.random_velocs
    x_velocs = torch.randn_like(x_coords).contiguous() if random_velocs else to(batch.atom_velocs)
    num_atoms_all = (~masked_elements).sum(dim=-1)
    segments = getattr(batch, "segments", None)
    masses = None
    if not random_velocs:  # :464-471
        ms = [loss.openmm_provider.get_masses(n).to(x_coords.device) for n in batch.names]
        masses = torch.stack([torch.nn.functional.pad(m, (0, atom_types.size(-1) - m.size(0)), "constant", 0) for m in ms])
    total = torch.tensor(0.0, device=x_coords.device)
    for _ in range(loss.num_samples):
        num_atoms = num_atoms_all
        y_coords, y_velocs, logp_xy = model.conditional_sample_with_logp(
            atom_types=atom_types, x_coords=x_coords, x_velocs=x_velocs, adj_list=adj_list, edge_batch_idx=edge_batch_idx,
            masked_elements=masked_elements, num_samples=1, logger=logger)
        y_coords, y_velocs, logp_xy = y_coords.squeeze(0), y_velocs.squeeze(0), logp_xy.squeeze(0)
        logp_yx = model.log_likelihood(
            atom_types=atom_types, x_coords=y_coords, x_velocs=y_velocs if random_velocs else -y_velocs, y_coords=x_coords,
            y_velocs=x_velocs if random_velocs else -x_velocs, adj_list=adj_list, edge_batch_idx=edge_batch_idx,
            masked_elements=masked_elements, logger=logger)  # :452-462
        energy_x, _ = compute_energy(x_coords, x_velocs, batch.names, masked_elements, loss.openmm_provider, random_velocs=random_velocs,
                                     masses=masses, segments=segments)
        energy_y, _ = compute_energy(y_coords, y_velocs, batch.names, masked_elements, loss.openmm_provider, random_velocs=random_velocs,
                                     masses=masses, segments=segments)
        energy_delta = energy_y - energy_x
        neg_log_acceptance = energy_delta + logp_xy - logp_yx  # :497
        per_sample = (torch.clamp(neg_log_acceptance, max=0) if loss.clamp else neg_log_acceptance) + loss.beta * logp_xy  # :513-518
        if loss.high_energy_threshold != -1:  # :522-537
            changes = loss.chirality_checker(batch, y_coords, masked_elements)
            energy_delta = energy_delta + 100000.0 * changes.to(energy_delta.dtype)
            good = energy_delta < loss.high_energy_threshold
            per_sample, num_atoms = per_sample[good.reshape(per_sample.shape)], num_atoms[good]
            if len(num_atoms) == 0:
                per_sample, num_atoms = torch.tensor(10000.0, device=x_coords.device), torch.tensor(1.0, device=x_coords.device)
        total = total + (per_sample / num_atoms).mean()
    return total / loss.num_samples


class NegativeLogLikelihoodLoss:
    """losses.py:306-319."""

    def __init__(self, random_velocs: bool = True):
        self.random_velocs = random_velocs


def nll_loss(loss: NegativeLogLikelihoodLoss, model, batch, device: Optional[Union[str, torch.device]] = None, logger=None) -> Tensor:
    """`get_loss(NegativeLogLikelihoodLoss, ...)`, losses.py:321-356: with `random_velocs` BOTH velocity tensors are fresh
    standard-normal draws (conditioning first, then target -- the order matters for seeded runs), then `model(...)`."""
    to = (lambda t: t.to(device, non_blocking=True)) if device is not None else (lambda t: t)
    x_coords, y_coords = to(batch.atom_coords), to(batch.atom_coord_targets)
    if loss.random_velocs:
        x_velocs = torch.randn_like(x_coords).contiguous()
        y_velocs = torch.randn_like(y_coords).contiguous()
    else:
        x_velocs, y_velocs = to(batch.atom_velocs), to(batch.atom_veloc_targets)
    return model(atom_types=to(batch.atom_types), x_coords=x_coords, x_velocs=x_velocs, y_coords=y_coords, y_velocs=y_velocs,
                 adj_list=to(batch.adj_list), edge_batch_idx=to(batch.edge_batch_idx), masked_elements=to(batch.masked_elements), logger=logger)


def get_loss(loss, model, batch, device: Optional[Union[str, torch.device]] = None, logger=None) -> Tensor:
    """losses.py:215-238: dispatch on the loss type (the reference uses `multimethod`)."""
    for cls, fn in ((NegativeLogLikelihoodLoss, nll_loss), (AcceptanceLoss, acceptance_loss), (EnergyLoss, energy_loss)):
        if isinstance(loss, cls):
            return fn(loss, model, batch, device=device, logger=logger)
    raise TypeError(f"no loss implementation for {type(loss).__name__}")


class LossWrapper(torch.nn.Module):
    """losses.py:241-272: `wrapper(batch, device=..., logger=...)` computes `loss` for `module`; a state dict saved from a
    wrapper ("module.…" keys, what DeepSpeed checkpoints hold) or from the bare module both load."""

    def __init__(self, module: torch.nn.Module, loss=None):
        super().__init__()
        self.module = module
        self.loss = loss

    def load_state_dict(self, state_dict, strict: bool = True):
        if not any(k.split(".", 1)[0] == "module" for k in state_dict):
            print("`state_dict` seems to be meant for `module`; loading this instead")
            return self.module.load_state_dict(state_dict, strict=strict)
        inner = {k.split(".", 1)[1]: v for k, v in state_dict.items() if k.split(".", 1)[0] == "module"}
        out = self.module.load_state_dict(inner, strict=strict)
        loss_sd = {k.split(".", 1)[1]: v for k, v in state_dict.items() if k.split(".", 1)[0] == "loss"}
        if loss_sd:
            if isinstance(self.loss, torch.nn.Module):
                self.loss.load_state_dict(loss_sd, strict=strict)
            elif strict:
                print(f"LossWrapper: loss is not a `torch.nn.Module` but `state_dict` contains the following loss parameters {list(loss_sd)}")
        return out

    def forward(self, *args, **kwargs):
        assert self.loss is not None, "`loss` is not given"
        return get_loss(self.loss, self.module, *args, **kwargs)


def wrap_or_replace_loss(model: torch.nn.Module, loss) -> LossWrapper:
    """losses.py:274-289: wrap `model`, or replace the loss of an (arbitrarily nested) wrapper."""
    return LossWrapper(module=unwrap_loss_wrapper(model), loss=loss)


def unwrap_loss_wrapper(model: torch.nn.Module) -> torch.nn.Module:
    """losses.py:292-303."""
    while isinstance(model, LossWrapper):
        model = model.module
    return model

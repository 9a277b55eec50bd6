"""Batch / wire formats on either side of the hot path: the containers and the dense collate of
the reference's dataloader.py, and a loader for its `<name>-traj-arrays.npz` + `<name>-traj-state0.pdb` trajectory files.

Same class and field names as the reference (dataloader.py:45-198), so `sample_with_model`, `sample_on_batches`,
`evaluate.py`-style drivers and the model's keyword interface take these batches unchanged:

    MolDynDatapoint, DenseMolDynBatch, moldyn_dense_collate_fn, lengths_to_mask      dataloader.py:58-76,109-198,328-413
    load_pdb_trace_data / TrajectoryInformation                                       dataloader.py:45-56,212-276

The reference reads the PDB topology with mdtraj (bonds from residue templates); here a minimal ATOM-record reader
plus the residue templates of `timewarp_b200.peptides` do the same for the residues those templates cover.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

from .peptides import ELEMENT_VOCAB, Peptide, _build_bonds


@dataclass
class TrajectoryInformation:
    name: str
    node_types: np.ndarray  # int32 [V]
    adj_list: np.ndarray  # int32 [E, 2]
    coord_features: List[np.ndarray]  # T x float32 [V, 3]
    veloc_features: List[np.ndarray]
    force_features: List[np.ndarray]
    coord_targets: List[np.ndarray]
    veloc_targets: List[np.ndarray]
    force_targets: List[np.ndarray]


@dataclass
class MolDynDatapoint:
    """One conditioning / target pair of one molecule (dataloader.py:58-76)."""

    name: str
    atom_types: torch.Tensor  # int64 [V]
    adj_list: torch.Tensor  # int64 [E, 2]
    atom_coords: torch.Tensor  # float32 [V, 3]
    atom_velocs: torch.Tensor
    atom_forces: torch.Tensor
    atom_coord_targets: torch.Tensor
    atom_veloc_targets: torch.Tensor
    atom_force_targets: torch.Tensor

    @property
    def num_atoms(self) -> int:
        return int(self.atom_types.shape[0])


_TENSOR_FIELDS = ("atom_types", "adj_list", "edge_batch_idx", "atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets",
                  "atom_veloc_targets", "atom_force_targets", "masked_elements")
_FLOAT_FIELDS = ("atom_coords", "atom_velocs", "atom_forces", "atom_coord_targets", "atom_veloc_targets", "atom_force_targets")


@dataclass
class DenseMolDynBatch:
    """Zero-padded batch (dataloader.py:109-198): [B, max_num_atoms, ...] tensors + `masked_elements` (True = padding)."""

    names: List[str]
    atom_types: torch.Tensor  # int64 [B, V]
    adj_list: torch.Tensor  # int64 [E, 2]
    edge_batch_idx: torch.Tensor  # int64 [E]
    atom_coords: torch.Tensor  # float32 [B, V, 3]
    atom_velocs: torch.Tensor
    atom_forces: torch.Tensor
    atom_coord_targets: torch.Tensor
    atom_veloc_targets: torch.Tensor
    atom_force_targets: torch.Tensor
    masked_elements: torch.Tensor  # bool [B, V]

    def _map(self, fn, fields=_TENSOR_FIELDS) -> "DenseMolDynBatch":
        kw = {f: getattr(self, f) for f in _TENSOR_FIELDS}
        kw.update({f: fn(getattr(self, f)) for f in fields})
        return DenseMolDynBatch(names=list(self.names), **kw)

    def pin_memory(self) -> "DenseMolDynBatch":
        return self._map(lambda t: t.pin_memory())

    def tofp16(self) -> "DenseMolDynBatch":
        return self._map(lambda t: t.half(), _FLOAT_FIELDS)

    def todevice(self, device: torch.device) -> "DenseMolDynBatch":
        return self._map(lambda t: t.to(device, non_blocking=True))

    def model_kwargs(self, device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
        """Keyword tensors of `model.forward` / `log_likelihood` (density_model_base.py:14-25) for this batch."""
        b = self if device is None else self.todevice(device)
        return dict(atom_types=b.atom_types, x_coords=b.atom_coords, x_velocs=b.atom_velocs, y_coords=b.atom_coord_targets,
                    y_velocs=b.atom_veloc_targets, adj_list=b.adj_list, edge_batch_idx=b.edge_batch_idx,
                    masked_elements=b.masked_elements)


def lengths_to_mask(lengths: torch.Tensor) -> torch.Tensor:
    """[B] lengths -> bool [B, max_len], True where the element is PADDING (dataloader.py:402-413)."""
    max_len = int(lengths.max()) if lengths.numel() else 0
    return torch.arange(max_len, device=lengths.device)[None, :] >= lengths[:, None]


def moldyn_dense_collate_fn(datapoints: Sequence[MolDynDatapoint], fp16: bool = False) -> DenseMolDynBatch:
    """dataloader.py:328-399: pad every per-atom tensor to the batch maximum, concatenate the adjacency lists and tag
    every edge with its batch index."""
    num_atoms = torch.tensor([d.num_atoms for d in datapoints], dtype=torch.int64)
    pad = lambda f: pad_sequence([getattr(d, f) for d in datapoints], batch_first=True)  # noqa: E731
    adj = torch.cat(tuple(d.adj_list for d in datapoints), dim=0)
    edge_batch_idx = torch.cat(tuple(i * torch.ones(size=(d.adj_list.shape[0],), dtype=torch.int64) for i, d in enumerate(datapoints)), dim=0)
    batch = DenseMolDynBatch(
        names=[d.name for d in datapoints], atom_types=pad("atom_types"), adj_list=adj, edge_batch_idx=edge_batch_idx,
        atom_coords=pad("atom_coords"), atom_velocs=pad("atom_velocs"), atom_forces=pad("atom_forces"),
        atom_coord_targets=pad("atom_coord_targets"), atom_veloc_targets=pad("atom_veloc_targets"),
        atom_force_targets=pad("atom_force_targets"), masked_elements=lengths_to_mask(num_atoms),
    )  # fmt: skip
    return batch.tofp16() if fp16 else batch


# ------------------------------------------------------------------------------------------ trajectory files
def read_pdb_topology(path: str, name: Optional[str] = None) -> Peptide:
    """ATOM / HETATM records of a PDB file (fixed columns) -> Peptide (coordinates in nm, bonds from the residue templates
    of timewarp_b200.peptides -- what mdtraj's standard-residue bonding gives the reference at dataloader.py:219-223)."""
    names, resn, resi, xyz = [], [], [], []
    with open(path) as f:
        for line in f:
            if line.startswith(("ATOM", "HETATM")):
                names.append(line[12:16].strip())
                resn.append(line[17:20].strip())
                resi.append(int(line[22:26]))
                xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            elif line.startswith("ENDMDL"):
                break
    coords = np.asarray(xyz, dtype=np.float64) / 10.0  # Angstrom -> nm
    return Peptide(name or path, names, resn, resi, coords, _build_bonds(names, resn, resi))


class CoordDeltaTooBig(ValueError):
    """dataloader.py: a conditioning/target pair further apart than any physical step (corrupt trajectory)."""


def load_pdb_trace_data(name: str, state0_file, traj_file: str, step_width: int = 1,
                        equal_data_spacing: bool = False) -> TrajectoryInformation:
    """dataloader.py:212-276: conditioning/target pairs (step, step + step_width) of one `*-traj-arrays.npz` trajectory.
    `state0_file` is the PDB path or an already built Peptide (topology)."""
    topo = state0_file if isinstance(state0_file, Peptide) else read_pdb_topology(state0_file, name)
    traj = np.load(traj_file)
    node_types = np.array([ELEMENT_VOCAB[e] for e in topo.elements], dtype=np.int32)
    adj_list = np.asarray(topo.bonds, dtype=np.int32)
    assert adj_list.min() >= 0 and adj_list.max() < len(node_types)
    by_step: Dict[int, Tuple[np.ndarray, np.ndarray, np.ndarray]] = {}
    for step, pos, vel, frc in zip(traj["step"], traj["positions"], traj["velocities"], traj["forces"]):
        by_step[int(step)] = (pos, vel, frc)
    steps = traj["step"][:100]
    spacing = int((steps[1:] - steps[:-1]).max()) * 10 // 9  # the spacing is always logarithmic (:236-240)
    if spacing <= step_width and not equal_data_spacing:
        warnings.warn(f"The step_width of {step_width} is larger than or equal to the spacing of {spacing} in the data. "
                      "This results in an unequal spacing between conditioning-target pairs.")
    out = TrajectoryInformation(name, node_types, adj_list, [], [], [], [], [], [])
    for step, (pos, vel, frc) in by_step.items():
        if step % spacing != 0 and equal_data_spacing:
            continue
        nxt = by_step.get(step + step_width)
        if nxt is None:
            continue
        delta = float(np.sqrt(np.sum((pos - nxt[0]) ** 2)))
        if delta > 100:
            raise CoordDeltaTooBig(f"{name} trajectory has {delta:g} distance between steps {step} and {step + step_width}")
        out.coord_features.append(pos), out.veloc_features.append(vel), out.force_features.append(frc)
        out.coord_targets.append(nxt[0]), out.veloc_targets.append(nxt[1]), out.force_targets.append(nxt[2])
    return out


def datapoints_from_trajectory(info: TrajectoryInformation) -> List[MolDynDatapoint]:
    """TrajectoryInformation -> MolDynDatapoints (what the reference's dataset classes yield, dataloader.py:416-470)."""
    at = torch.from_numpy(info.node_types.astype(np.int64))
    adj = torch.from_numpy(info.adj_list.astype(np.int64))
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    return [MolDynDatapoint(info.name, at, adj, f(c), f(v), f(fr), f(ct), f(vt), f(ft))
            for c, v, fr, ct, vt, ft in zip(info.coord_features, info.veloc_features, info.force_features, info.coord_targets,
                                            info.veloc_targets, info.force_targets)]

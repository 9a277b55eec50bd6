"""Data augmentation of a dense batch: a random rigid motion (`data_augmentation=True` of
`sample_on_batches` / the training loop).  Restates equivariance/equivariance_transforms.py:15-175 and
equivariance/equivariance_utils.py:5-32: transformations act on coordinates, velocity-like vectors (velocities AND forces),
point-wise features and adjacency lists; `a + b` composes; `transform_batch` applies translation then rotation.
Random draws follow the reference: the translation is ONE `torch.randn(1, 3)`, the rotation comes from
`scipy.spatial.transform.Rotation.random()` (numpy's global generator), so seeded runs draw the same motion."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import Tensor

from .dataloader import DenseMolDynBatch


class BaseDataTransformation:
    """Identity on every kind of attribute (equivariance_transforms.py:15-34)."""

    def transform_pointwise_feature(self, pointwise_feature: Tensor) -> Tensor:
        return pointwise_feature

    def transform_coord(self, coord: Tensor) -> Tensor:
        return coord

    def transform_veloc(self, veloc: Tensor) -> Tensor:
        return veloc

    def transform_adjacency_list(self, adj_list: Tensor) -> Tensor:
        return adj_list

    def __add__(self, other):
        return CompositionTransformation([self]) + other


class CompositionTransformation:
    """Applies its members left to right (equivariance_transforms.py:37-85).  The reference's adjacency-list method feeds every
    member the ORIGINAL list and returns the last member's result (`:70-71`); kept, because it is observable behaviour."""

    def __init__(self, transforms: Sequence[BaseDataTransformation]):
        self.transforms: List[BaseDataTransformation] = list(transforms)

    def _chain(self, method: str, value: Tensor) -> Tensor:
        for t in self.transforms:
            value = getattr(t, method)(value)
        return value

    def transform_pointwise_feature(self, pointwise_feature: Tensor) -> Tensor:
        return self._chain("transform_pointwise_feature", pointwise_feature)

    def transform_coord(self, coord: Tensor) -> Tensor:
        return self._chain("transform_coord", coord)

    def transform_veloc(self, veloc: Tensor) -> Tensor:
        return self._chain("transform_veloc", veloc)

    def transform_adjacency_list(self, adj_list: Tensor) -> Tensor:
        out = adj_list
        for t in self.transforms:
            out = t.transform_adjacency_list(adj_list)
        return out

    def __add__(self, other):
        if isinstance(other, BaseDataTransformation):
            return CompositionTransformation(self.transforms + [other])
        if isinstance(other, CompositionTransformation):
            return CompositionTransformation(self.transforms + other.transforms)
        return NotImplemented


class Permutation(BaseDataTransformation):
    """Re-labels the points of ONE un-batched sample (equivariance_transforms.py:88-118)."""

    def __init__(self, permutation: Tensor):
        self.permutation = permutation
        self.inv_permutation = torch.argsort(permutation)

    def _rows(self, x: Tensor) -> Tensor:
        if x.ndim > 2:
            raise NotImplementedError(f"Permutation transform doesn't work with shape {x.shape}")
        return x[self.inv_permutation]

    transform_pointwise_feature = _rows
    transform_coord = _rows
    transform_veloc = _rows

    def transform_adjacency_list(self, adj_list: Tensor) -> Tensor:
        return self.permutation[adj_list]


class Rotation(BaseDataTransformation):
    """x -> R x for coordinates and for velocity-like vectors (equivariance_transforms.py:121-129)."""

    def __init__(self, rotation_matrix: Tensor):
        self.rotation_matrix = rotation_matrix

    def _rotate(self, x: Tensor) -> Tensor:
        return (self.rotation_matrix.to(x.device) @ x.transpose(-1, -2)).transpose(-1, -2)

    transform_coord = _rotate
    transform_veloc = _rotate


class Translation(BaseDataTransformation):
    """x -> x + a for coordinates only (equivariance_transforms.py:132-137)."""

    def __init__(self, translation_vector: Tensor):
        self.translation_vector = translation_vector

    def transform_coord(self, coord: Tensor) -> Tensor:
        return coord + self.translation_vector.to(coord.device)


def random_rotation_matrix(device=None, dtype=torch.float32) -> Tensor:
    """Haar-uniform rotation (equivariance_utils.py:5-8)."""
    from scipy.spatial.transform import Rotation as R

    return torch.tensor(R.random().as_matrix(), dtype=dtype).to(device)


def random_translation_vector(device=None, dtype=torch.float32) -> Tensor:
    """Standard-normal translation [1, 3] (equivariance_utils.py:16-19)."""
    return torch.randn(1, 3, dtype=dtype).to(device)


def random_permutation(num_points: int, device=None) -> Tensor:
    return torch.randperm(num_points).to(device)


class RandomPermutation(Permutation):
    def __init__(self, num_points: int, device: Optional[str] = None):
        super().__init__(random_permutation(num_points, device=device))


class RandomRotation(Rotation):
    def __init__(self, device: Optional[str] = None, dtype=torch.float32):
        super().__init__(random_rotation_matrix(device=device, dtype=dtype))


class RandomTranslation(Translation):
    def __init__(self, device: Optional[str] = None, dtype=torch.float32):
        super().__init__(random_translation_vector(device=device, dtype=dtype))


def transform_batch(batch: DenseMolDynBatch, transform=None, dtype=torch.float32) -> DenseMolDynBatch:
    """equivariance_transforms.py:153-175: the batch under `transform` (default: one random translation followed by one random
    rotation, shared by every sample of the batch).  Forces transform like velocities; names, edge batch index and the padding
    mask pass through."""
    if transform is None:
        transform = RandomTranslation(dtype=dtype) + RandomRotation(dtype=dtype)
    return DenseMolDynBatch(
        names=batch.names,
        atom_types=transform.transform_pointwise_feature(batch.atom_types),
        adj_list=transform.transform_adjacency_list(batch.adj_list),
        edge_batch_idx=batch.edge_batch_idx,
        atom_coords=transform.transform_coord(batch.atom_coords),
        atom_velocs=transform.transform_veloc(batch.atom_velocs),
        atom_forces=transform.transform_veloc(batch.atom_forces),
        atom_coord_targets=transform.transform_coord(batch.atom_coord_targets),
        atom_veloc_targets=transform.transform_veloc(batch.atom_veloc_targets),
        atom_force_targets=transform.transform_veloc(batch.atom_force_targets),
        masked_elements=batch.masked_elements,
    )

"""Chirality veto (utils/chirality.py:14-80): host-side centre search + device-side sign kernel."""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib


def find_chirality_centers(adj_list: Tensor, atom_types: Tensor, num_h_atoms: int = 2) -> Tensor:
    """Atoms with exactly 4 bonds and more than `num_h_atoms` non-hydrogen neighbours, returned as
    rows [centre, n1, n2, n3] (first three bonded atoms in bond-list order)  -- utils/chirality.py:14-38.
    One-off host-side topology scan.  Like the reference, candidates are positions in the sorted
    unique atom list of `adj_list` (identical to atom indices when every atom has a bond)."""
    adj = np.asarray(adj_list.detach().cpu().numpy())
    types = np.asarray(atom_types.detach().cpu().numpy())[0]
    _, counts = np.unique(adj, return_counts=True)
    centers = []
    for center in np.nonzero(counts == 4)[0]:
        rows, cols = np.nonzero(adj == center)
        bonded = adj[rows, (cols + 1) % 2]
        if np.count_nonzero(types[bonded] - 1) > num_h_atoms:  # element id 1 == hydrogen
            centers.append([int(center), *[int(b) for b in bonded[:3]]])
    return torch.tensor(centers, dtype=adj_list.dtype, device=adj_list.device).reshape(-1, 4)


def _run(coords: Tensor, centers: Tensor, ref_signs, want_signs: bool):
    if coords.device.type != "cuda":
        raise _lib.TimewarpB200Error("chirality kernels run on CUDA only (no CPU fallback)")
    assert coords.dim() == 3
    x = coords.to(torch.float32).contiguous()
    B, V = x.shape[:2]
    c = centers.to(device=x.device, dtype=torch.int64).contiguous()
    C_ = c.shape[0]
    signs = torch.empty(B, C_, dtype=torch.float32, device=x.device) if want_signs else None
    changed = torch.empty(B, dtype=torch.uint8, device=x.device) if ref_signs is not None else None
    ref = ref_signs.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous() if ref_signs is not None else None
    # one reference sign per centre (the reference's [1, C] row, utils/chirality.py:65-80); per-chain rows are not supported
    assert ref is None or ref.numel() == C_, f"reference_signs must hold one sign per centre ({C_}), got shape {tuple(ref_signs.shape)}"
    _lib.check(
        _lib.load().tw_chirality(_lib.ptr(x), _lib.ptr(c) if C_ else None, _lib.ptr(ref) if (ref is not None and C_) else None,
                                 B, V, C_, _lib.ptr(changed), _lib.ptr(signs), torch.cuda.current_stream(x.device).cuda_stream),
        "tw_chirality",
    )
    return signs, changed


def compute_chirality_sign(coords: Tensor, chirality_centers: Tensor) -> Tensor:
    """[B,C] signs of the triple product of the three neighbour directions (utils/chirality.py:41-62)."""
    signs, _ = _run(coords, chirality_centers, None, True)
    return signs


def check_symmetry_change(coords: Tensor, chirality_centers: Tensor, reference_signs: Tensor) -> Tensor:
    """[B] bool, True where any centre's sign differs from the reference (utils/chirality.py:65-80)."""
    if chirality_centers.shape[0] == 0:
        return torch.zeros(coords.shape[0], dtype=torch.bool, device=coords.device)
    _, changed = _run(coords, chirality_centers, reference_signs, False)
    return changed.view(torch.bool)

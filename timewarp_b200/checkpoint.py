"""Checkpoint import / export in the reference's on-disk format: utilities/model_utils.py:12-63.

A reference checkpoint is a `torch.save`d dict whose weights sit under "model_state_dict" (`save_model`) or "module"
(DeepSpeed engines, train_deepspeed.py); the state-dict keys of `custom_transformer_nvp_constructor` models are identical
here, so the weights load verbatim (tests/test_abi_cpu.py::test_reference_checkpoint_loads)."""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Optional

import torch


def save_model(path, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, lr_scheduler=None, **kwargs) -> None:
    """utilities/model_utils.py:12-29."""
    data: Dict[str, Any] = {"model_state_dict": model.state_dict()}
    if optimizer is not None:
        data["optimizer_state_dict"] = optimizer.state_dict()
    if lr_scheduler is not None:
        data["lr_scheduler_state_dict"] = lr_scheduler.state_dict()
    data.update(kwargs)
    torch.save(data, path)


def _find(path, file_name: str) -> str:
    if os.path.isfile(path):
        return str(path)
    hits = [os.path.join(d, file_name) for d, _, files in os.walk(path) if file_name in files]
    assert len(hits) == 1, f"Tried to call unique_item, but {hits} contains {len(hits)} items."  # utilities/common.py:35-39
    return hits[0]


def load_checkpoint_in_subdir(path, file_name: str = "best_model.pt", weights_only: bool = False):
    """utilities/model_utils.py:32-36: `path` is the checkpoint file or a directory with exactly one `file_name` below it.
    weights_only=False unpickles arbitrary objects like the reference does (its checkpoints carry the training config);
    pass True for files from untrusted sources."""
    return torch.load(_find(path, file_name), map_location="cpu", weights_only=weights_only)


def load_model_state_dict(path, file_name: str = "best_model.pt", weights_only: bool = False):
    """utilities/model_utils.py:39-43."""
    data = load_checkpoint_in_subdir(path, file_name, weights_only)
    return data["model_state_dict" if "model_state_dict" in data else "module"]


def load_model(path, model_constructor: Callable, file_name: str = "best_model.pt", weights_only: bool = False) -> torch.nn.Module:
    """utilities/model_utils.py:46-63: `model_constructor(checkpoint_dict) -> nn.Module`, then the weights are loaded."""
    data = load_checkpoint_in_subdir(path, file_name, weights_only)
    model = model_constructor(data)
    model.load_state_dict(data["model_state_dict" if "model_state_dict" in data else "module"])
    return model

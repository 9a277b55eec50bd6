"""Deterministic synthetic weights for benchmarks and profiling runs (no checkpoints are reachable offline).

Every tensor is a pure function of (state-dict key, shape, seed), independent of module construction order, with the
scale of torch's default initialisation: U(-1, 1)/sqrt(fan_in) for Linear weights and biases, LayerNorm gamma 1 + 0.1 U and
beta 0.1 U, embedding N(0, 1), small prior log-scales.  tests/test_abi_cpu.py checks that the test oracle's generator
(oracle/flow_oracle.py::synth_state_dict, used to load the reference model when the golden vectors were made) produces the
same tensors, so a bench run and its CPU baseline see identical parameters."""
from __future__ import annotations

import math
import zlib
from typing import Dict

import torch

from .modules import CHEB_COEFFS_EXPMX


def synth_state_dict(model: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic tensors for every entry of `model.state_dict()` (load with `model.load_state_dict(..., strict=True)`)."""
    ref = model.state_dict()
    out: Dict[str, torch.Tensor] = {}
    for key, cur in ref.items():
        shape = tuple(cur.shape)
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2**31))
        if key.endswith("cheb_coeffs"):  # the reference's initial value + a different perturbation in every layer and head
            base = torch.tensor((CHEB_COEFFS_EXPMX + [0.0] * max(0, shape[1] - len(CHEB_COEFFS_EXPMX)))[: shape[1]])
            t = base[None, :].expand(shape) + 0.02 * (torch.rand(shape, generator=g) * 2 - 1)
        elif key.endswith("log_lengthscales"):  # a different value in every layer
            ls = ref[key[: -len("log_lengthscales")] + "lengthscales"].detach().cpu().float()
            t = torch.log(ls) + 0.3 * (torch.rand(shape, generator=g) * 2 - 1)
        elif key.endswith("lengthscales"):  # configuration, not a weight
            t = cur.detach().cpu().float().clone()
        elif key.endswith("prior_log_scale"):
            t = 0.2 * (torch.rand((), generator=g) - 0.5)
        elif key == "flow.atom_embedder.weight":
            t = torch.randn(shape, generator=g)
        elif ".norm" in key:
            u = torch.rand(shape, generator=g) * 2 - 1
            t = 1 + 0.1 * u if key.endswith("weight") else 0.1 * u
        elif key.endswith(".weight"):
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        else:  # Linear bias: fan_in of the matching weight
            fan_in = ref[key[: -len("bias")] + "weight"].shape[1]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        out[key] = t.to(torch.float32)
    return out

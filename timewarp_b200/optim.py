"""Optimizer step of the NLL / energy training loop as ONE kernel (`tw_adam_step`, csrc/optim.cu).

The reference trains with `torch.optim.Adam(model.parameters(), lr, weight_decay)` (utilities/training_utils.py:356-368).
On this model that is 659 parameter tensors, most of them 128 x 128: torch's fused multi-tensor Adam needs ~50 launches and
1.1 ms per step for 1 GB of traffic.  The hand-written backward already returns every gradient as a slice of one flat buffer
(flow.py `_gradient_table`); `FlatAdam` re-homes the parameters into a buffer with the same slice offsets (the tensors keep
their identity, shapes and state-dict keys -- only their storage moves) and keeps both moment buffers flat, so a step is a
single 28-bytes-per-element stream over the buffers.  Same update rule as `torch.optim.Adam` (bias-corrected moments, L2
weight decay added to the gradient, no amsgrad), checked against it step by step in tests/test_gpu_train.py.

The hyper-parameters live in a small device tensor: a captured CUDA graph of the training step follows a learning-rate
schedule (`param_groups[0]["lr"]` is uploaded by `step()` when it changed; under graph replay call `sync_hyper()`).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    """Adam over the flat gradient buffer of a `ConditionalFlowDensityModel` (CUDA, tensor-core precisions).

    Every trainable parameter of `model` must receive its gradient from the model's own backward (as a view of
    `model._last_flat_grad`); a parameter with a foreign gradient raises -- use `torch.optim.Adam` for such modules."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("FlatAdam: the model has no trainable parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.model = model
        self._flat_p: Optional[Tensor] = None
        self._m: Optional[Tensor] = None
        self._v: Optional[Tensor] = None
        self._hyper: Optional[Tensor] = None  # device: lr, beta1, beta2, eps, weight_decay, step
        self._hyper_host = None
        self._offsets = None

    # ------------------------------------------------------------------ layout
    def _param_offsets(self, flat_g: Tensor, params=None):
        base, end = flat_g.data_ptr(), flat_g.data_ptr() + flat_g.numel() * 4
        offs = []
        for p in (self.param_groups[0]["params"] if params is None else params):
            g = p.grad
            if g is None:
                raise _lib.TimewarpB200Error("FlatAdam: a trainable parameter has no gradient (run the model's backward first)")
            a = g.data_ptr()
            if not (g.dtype == torch.float32 and g.is_contiguous() and base <= a and a + g.numel() * 4 <= end):
                raise _lib.TimewarpB200Error("FlatAdam: a gradient does not live in the model's flat gradient buffer "
                                             "(foreign module or a second autograd node: use torch.optim.Adam)")
            offs.append((a - base) // 4)
        return offs

    def _adopt(self, flat_g: Tensor) -> None:
        """First step: move every parameter into a flat buffer laid out like the gradient buffer."""
        offs = self._param_offsets(flat_g)
        flat_p = torch.zeros_like(flat_g)
        for p, o in zip(self.param_groups[0]["params"], offs):
            view = flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat_p, self._offsets = flat_p, offs
        self._m, self._v = torch.zeros_like(flat_g), torch.zeros_like(flat_g)
        self._hyper = torch.zeros(8, dtype=torch.float32, device=flat_g.device)
        # the model caches raw parameter pointers and the packed bf16 weight images: both refer to the old storage
        self.model._table = None
        self.model._packed = None

    def sync_hyper(self) -> None:
        """Upload lr / betas / eps / weight_decay if they changed on the host (not allowed while a graph is being captured)."""
        g = self.param_groups[0]
        host = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if host != self._hyper_host:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.TimewarpB200Error("FlatAdam: hyper-parameters changed during CUDA graph capture; call sync_hyper() before")
            self._hyper[:5].copy_(torch.tensor(host, dtype=torch.float32), non_blocking=False)
            self._hyper_host = host

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flat_g = getattr(self.model, "_last_flat_grad", None)
        if flat_g is None or not flat_g.is_cuda:
            raise _lib.TimewarpB200Error("FlatAdam: no flat gradient buffer (the model's CUDA training backward has not run)")
        if self._flat_p is None:
            self._adopt(flat_g)
        else:  # cheap per-step checks (first and last parameter): same gradient layout, parameters still inside the flat buffer
            ends = [self.param_groups[0]["params"][0], self.param_groups[0]["params"][-1]]
            if flat_g.numel() != self._flat_p.numel() or self._param_offsets(flat_g, ends) != [self._offsets[0], self._offsets[-1]]:
                raise _lib.TimewarpB200Error("FlatAdam: the gradient layout changed between steps")
            for p, o in zip(ends, (self._offsets[0], self._offsets[-1])):  # `.to()` / `.data = ...` would silently detach a parameter
                if p.data_ptr() != self._flat_p.data_ptr() + 4 * o:
                    raise _lib.TimewarpB200Error("FlatAdam: a parameter left the flat buffer (model.to() after the first step?)")
        self.sync_hyper()
        self._hyper[5:6].add_(1.0)
        stream = torch.cuda.current_stream(flat_g.device).cuda_stream
        _lib.check(_lib.load().tw_adam_step(_lib.ptr(self._flat_p), _lib.ptr(flat_g), _lib.ptr(self._m), _lib.ptr(self._v),
                                            flat_g.numel(), _lib.ptr(self._hyper), stream), "tw_adam_step")
        return loss

    # ------------------------------------------------------------------ checkpointing (utilities/model_utils.py:12-32)
    def state_dict(self):
        g = self.param_groups[0]
        return {"flat": True, "step": None if self._hyper is None else float(self._hyper[5]),
                "exp_avg": self._m, "exp_avg_sq": self._v,
                "hyper": dict(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])}

    def load_state_dict(self, sd) -> None:
        if not sd.get("flat"):
            raise ValueError("FlatAdam.load_state_dict: not a FlatAdam state")
        self.param_groups[0].update(sd["hyper"])
        if sd["exp_avg"] is None:
            return
        if self._flat_p is None:
            raise _lib.TimewarpB200Error("FlatAdam.load_state_dict: run one training step first (the flat layout is fixed by the first backward)")
        self._m.copy_(sd["exp_avg"]), self._v.copy_(sd["exp_avg_sq"])
        self._hyper[5] = sd["step"]


def get_optimizer(model, config) -> torch.optim.Optimizer:
    """utilities/training_utils.py:356-368 (`config.optimizer == "Adam"`, `learning_rate`, `weight_decay`)."""
    assert getattr(config, "optimizer", "Adam") == "Adam"
    return FlatAdam(model, lr=config.learning_rate, weight_decay=getattr(config, "weight_decay", 0.0))

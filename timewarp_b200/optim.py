"""Optimizer step of the NLL / energy training loop as ONE kernel (`tw_adam_step`, csrc/optim.cu).

The reference trains with `torch.optim.Adam(model.parameters(), lr, weight_decay)` (utilities/training_utils.py:356-368).
On this model that is 659 parameter tensors, most of them 128 x 128: torch's fused multi-tensor Adam needs ~50 launches and
1.1 ms per step for 1 GB of traffic.  The hand-written backward already returns every gradient as a slice of one flat buffer
(flow.py `_gradient_table`); `FlatAdam` re-homes the parameters into a buffer with the same slice offsets (the tensors keep
their identity, shapes and state-dict keys -- only their storage moves) and keeps both moment buffers flat, so a step is a
single 28-bytes-per-element stream over the buffers (gradients that are not views of that buffer are gathered first).  Same update rule as `torch.optim.Adam` (bias-corrected moments, L2
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
    """Adam over flat parameter / gradient / moment buffers (CUDA): one launch per step.

    Fast path: every `p.grad` is the view of `model._last_flat_grad` that the model's own backward handed out (autograd's
    AccumulateGrad keeps such a gradient as it is when nothing else references it) -- the kernel reads that buffer directly.
    Otherwise (autograd cloned the gradients -- it does under compute-sanitizer, for instance --, a second autograd node, a
    foreign module, gradients the data-parallel trainer reduced in buckets) the gradients are first gathered into a flat
    buffer of the optimizer's own with one `torch._foreach_copy_`.  A parameter without a gradient keeps zero moments and
    does not move, like in torch.optim.Adam."""

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
        self._gbuf: Optional[Tensor] = None   # gather path: the optimizer's own flat gradient buffer ...
        self._gviews = None                   # ... and its per-parameter views
        self._pending = None                  # a state dict loaded before the first step
        self.fast_path_steps = 0              # steps that read the model's flat gradient buffer directly
        self.gather_steps = 0

    # ------------------------------------------------------------------ layout
    def _view_offsets(self, flat_g: Optional[Tensor]):
        """Offsets of every p.grad inside flat_g, or None if any gradient is missing / lives elsewhere."""
        if flat_g is None or not flat_g.is_cuda or flat_g.dtype != torch.float32:
            return None
        base, end = flat_g.data_ptr(), flat_g.data_ptr() + flat_g.numel() * 4
        offs = []
        for p in self.param_groups[0]["params"]:
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous():
                return None
            a = g.data_ptr()
            if not (base <= a and a + g.numel() * 4 <= end):
                return None
            offs.append((a - base) // 4)
        return offs

    def _adopt(self, flat_g: Optional[Tensor]) -> None:
        """First step: move every parameter into a flat buffer -- laid out like the model's gradient buffer when the gradients
        are views of it (so that later steps can read it in place), else packed in parameter order (16-byte aligned slices)."""
        params = self.param_groups[0]["params"]
        dev = params[0].device
        if not params[0].is_cuda:
            raise _lib.TimewarpB200Error("FlatAdam: no CPU fallback -- the parameters must live on a CUDA device")
        offs = self._view_offsets(flat_g)
        if offs is not None:
            n = flat_g.numel()
        else:
            offs, n = [], 0
            for p in params:
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4
        flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            if p.dtype != torch.float32:
                raise TypeError("FlatAdam: parameters must be float32")
            view = flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat_p, self._offsets = flat_p, offs
        self._m, self._v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
        self._hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        if self._pending is not None:
            self._apply_state(self._pending)
            self._pending = None
        # the model caches raw parameter pointers and the packed bf16 weight images: both refer to the old storage
        if hasattr(self.model, "_table"):
            self.model._table = None
        if hasattr(self.model, "_packed"):
            self.model._packed = None

    def _gather(self) -> Tensor:
        params = self.param_groups[0]["params"]
        if self._gbuf is None:
            self._gbuf = torch.zeros_like(self._flat_p)
            self._gviews = [self._gbuf[o:o + p.numel()].view(p.shape) for p, o in zip(params, self._offsets)]
        have = [(v, p.grad) for v, p in zip(self._gviews, params) if p.grad is not None]
        if len(have) != len(params):
            self._gbuf.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        return self._gbuf

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
        params = self.param_groups[0]["params"]
        if all(p.grad is None for p in params):
            raise _lib.TimewarpB200Error("FlatAdam: no gradient to apply (the backward has not run)")
        flat_g = getattr(self.model, "_last_flat_grad", None)
        if self._flat_p is None:
            self._adopt(flat_g)
        else:
            first, last = params[0], params[-1]  # `.to()` / `.data = ...` would silently detach a parameter from the flat buffer
            if (first.data_ptr() != self._flat_p.data_ptr() + 4 * self._offsets[0]
                    or last.data_ptr() != self._flat_p.data_ptr() + 4 * self._offsets[-1]):
                raise _lib.TimewarpB200Error("FlatAdam: a parameter left the flat buffer (model.to() after the first step?)")
        if flat_g is not None and flat_g.numel() == self._flat_p.numel() and self._view_offsets(flat_g) == self._offsets:
            g_buf = flat_g
            self.fast_path_steps += 1
        else:
            g_buf = self._gather()
            self.gather_steps += 1
        self.sync_hyper()
        self._hyper[5:6].add_(1.0)
        stream = torch.cuda.current_stream(g_buf.device).cuda_stream
        _lib.check(_lib.load().tw_adam_step(_lib.ptr(self._flat_p), _lib.ptr(g_buf), _lib.ptr(self._m), _lib.ptr(self._v),
                                            g_buf.numel(), _lib.ptr(self._hyper), stream), "tw_adam_step")
        return loss

    # ------------------------------------------------------------------ checkpointing (utilities/model_utils.py:12-32)
    @staticmethod
    def _remap(src: Tensor, src_offsets, dst: Tensor, dst_offsets, numels) -> None:
        """Copy per-parameter slices between two flat layouts (a checkpoint may come from a run with the other layout)."""
        for so, do, n in zip(src_offsets, dst_offsets, numels):
            dst[do:do + n].copy_(src[so:so + n])

    def state_dict(self):
        g = self.param_groups[0]
        return {"flat": True, "step": None if self._hyper is None else float(self._hyper[5]),
                "exp_avg": self._m, "exp_avg_sq": self._v, "offsets": None if self._offsets is None else list(self._offsets),
                "numels": [p.numel() for p in g["params"]],
                "hyper": dict(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])}

    def load_state_dict(self, sd) -> None:
        if not sd.get("flat"):
            raise ValueError("FlatAdam.load_state_dict: not a FlatAdam state")
        self.param_groups[0].update(sd["hyper"])
        if sd["exp_avg"] is None:
            return
        if sd["numels"] != [p.numel() for p in self.param_groups[0]["params"]]:
            raise ValueError("FlatAdam.load_state_dict: the checkpoint belongs to a model with different parameters")
        if self._flat_p is None:
            self._pending = sd  # the flat layout is fixed by the first backward: the moments are restored inside the first step()
            return
        self._apply_state(sd)

    def _apply_state(self, sd) -> None:
        numels = sd["numels"]
        self._m.zero_(), self._v.zero_()
        self._remap(sd["exp_avg"].to(self._m.device), sd["offsets"], self._m, self._offsets, numels)
        self._remap(sd["exp_avg_sq"].to(self._v.device), sd["offsets"], self._v, self._offsets, numels)
        self._hyper[5] = float(sd["step"])


def get_optimizer(model, config) -> torch.optim.Optimizer:
    """utilities/training_utils.py:356-368 (`config.optimizer == "Adam"`, `learning_rate`, `weight_decay`)."""
    assert getattr(config, "optimizer", "Adam") == "Adam"
    return FlatAdam(model, lr=config.learning_rate, weight_decay=getattr(config, "weight_decay", 0.0))

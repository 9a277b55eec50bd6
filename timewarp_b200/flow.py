"""ConditionalFlowDensityModel -- the drop-in for the reference's density-model wrapper
(modules/model_wrappers/flow.py:106-336, density_model_base.py:10-88) on top of the CUDA library.

Same method names, keyword names, shapes, RNG consumption and state_dict keys as the reference;
every flow pass is ONE call into libtimewarp_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch import BoolTensor, Tensor

from . import _lib
from .modules import ConditionalSequentialFlow


def _require_cuda(name: str, t: Tensor, dtype=None) -> Tensor:
    if not isinstance(t, Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device.type != "cuda":
        raise _lib.TimewarpB200Error(
            f"{name} is on {t.device}: timewarp_b200 computes on CUDA (sm_100a) only -- there is no CPU fallback"
        )
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def _aligned(buf: Tensor) -> int:
    return (buf.data_ptr() + 1023) // 1024 * 1024


def _gradient_table(model, dev, ls_grad: bool):
    """ONE zero-filled flat buffer for every gradient (and for the scratch of frozen parameters the kernels still accumulate
    into) -- 659 zeros_like fills cost more than the backward GEMMs at small batch sizes -- and the table of pointers into it.
    ls_grad: also request dL/d(lengthscales of the pass) (learnable_kernel)."""
    tensors = model._ordered_params()
    need = [t.requires_grad and t.is_floating_point() for t in tensors]
    wants = [n or (t.is_floating_point() and not _is_optional_grad(model, i)) for i, (t, n) in enumerate(zip(tensors, need))]
    ls_slot = 3 + 2 * (model._cfg.num_mlp_hidden + 1) + 1  # lengthscales of chain[0].scale_transformer.encoder_layers[0]
    if ls_grad:
        wants[ls_slot] = True  # a non-NULL entry asks the backward for dL/d(lengthscales of the pass)
    sizes = [((t.numel() + 3) // 4 * 4 if w else 0) for t, w in zip(tensors, wants)]  # 16-byte aligned slices
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    views, off = [], 0
    for t, w, n in zip(tensors, wants, sizes):
        views.append(flat[off:off + t.numel()].view(t.shape) if w else None)
        off += n
    grads = [v if nd else None for v, nd in zip(views, need)]
    model._last_flat_grad = flat  # data-parallel training all-reduces this buffer directly (distributed.py)
    gtable = (C.c_void_p * len(tensors))(*[(v.data_ptr() if v is not None else None) for v in views])
    return tensors, views, grads, gtable, ls_slot, ls_grad


class _LogLikelihoodFn(torch.autograd.Function):
    """log p(y|x) with a hand-written backward: the forward records a tape in device memory
    (tw_flow_log_likelihood_train), the backward accumulates the gradient of every trainable
    parameter (tw_flow_log_likelihood_backward).  The parameters are passed as inputs so that
    autograd routes the gradients to them; coordinates get no gradient (NLL training, losses.py:321-356)."""

    @staticmethod
    def forward(ctx, model, atom_types, x_coords, x_velocs, y_coords, y_velocs, mask_u8, *params):
        # learnable_kernel: the LAST input is log_lengthscales of the first attention layer (the only one the reference's
        # pass reads, and the only one that receives a gradient); the table itself holds exp() of it
        ctx.n_extra = 1 if model._learnable else 0
        lib = _lib.load()
        dev = x_coords.device
        B, V = x_coords.shape[0], x_coords.shape[1]
        cfg = model._cfg
        tape_b, ws_b = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.tw_flow_train_bytes(C.byref(cfg), B, V, C.byref(tape_b), C.byref(ws_b)), "tw_flow_train_bytes")
        tape = torch.empty(tape_b.value + 1024, dtype=torch.uint8, device=dev)
        out = torch.empty(B, dtype=torch.float32, device=dev)
        table = model._param_table(dev)
        packed = model._packed_weights(dev, force=True)  # the optimizer may have stepped without a version bump
        model._stale_after_train = True  # ... and will probably step again before the next inference-path call
        _lib.check(
            lib.tw_flow_log_likelihood_train(
                C.byref(cfg), table, _lib.ptr(atom_types), _lib.ptr(x_coords), _lib.ptr(x_velocs), _lib.ptr(y_coords),
                _lib.ptr(y_velocs), _lib.ptr(mask_u8), B, V, model._flags(), _lib.ptr(out), packed, _aligned(tape),
                tape_b.value, model._stream(dev),
            ),
            "tw_flow_log_likelihood_train",
        )  # fmt: skip
        ctx.model, ctx.tape, ctx.sizes = model, tape, (B, V, tape_b.value, ws_b.value)
        ctx.packed_key = model._packed[2]
        # learnable_kernel: the table points at ONE buffer that every pass overwrites with its own lengthscales (density: first
        # coupling layer, sampling: last); a loss with both passes in its graph must see each pass's values in its backward
        ctx.ls_eff = model._ls_eff.clone() if model._learnable else None
        ctx.save_for_backward(atom_types, x_velocs, mask_u8)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        model = ctx.model
        atom_types, x_velocs, mask_u8 = ctx.saved_tensors
        dev = x_velocs.device
        B, V, tape_b, ws_b = ctx.sizes
        if ctx.tape is None:
            raise _lib.TimewarpB200Error("backward ran twice over the same graph: the activation tape is released after the first "
                                         "backward (retain_graph=True is not supported)")
        # (the re-pack epoch may differ: an inference call in between re-packs the same parameters into the same buffer)
        if model._packed is None or (model._packed[2][0], model._packed[2][2]) != (ctx.packed_key[0], ctx.packed_key[2]):
            raise _lib.TimewarpB200Error("parameters were modified between the forward and the backward pass")
        if ctx.ls_eff is not None:
            model._ls_eff.copy_(ctx.ls_eff)
        tensors, views, grads, gtable, ls_slot, ls_grad = _gradient_table(model, dev, ctx.n_extra == 1 and ctx.needs_input_grad[-1])
        table = model._param_table(dev)
        ws = torch.empty(ws_b + 1024, dtype=torch.uint8, device=dev)
        g_in = grad_out.to(torch.float32).contiguous()
        # gradients w.r.t. the inputs (x_coords, x_velocs, y_coords, y_velocs = inputs 2..5): AcceptanceLoss conditions the
        # reverse-move density on the proposal (losses.py:452-462)
        want_x = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        want_y = ctx.needs_input_grad[4] or ctx.needs_input_grad[5] or (want_x and model.use_displacement_as_target)
        new3 = lambda: torch.empty(B, V, 3, dtype=torch.float32, device=dev)  # noqa: E731
        dxc, dxv = (new3(), new3()) if want_x else (None, None)
        dz0c, dz0v = (new3(), new3()) if want_y else (None, None)
        _lib.check(
            lib.tw_flow_log_likelihood_backward_inputs(
                C.byref(model._cfg), table, gtable, _lib.ptr(atom_types), _lib.ptr(x_velocs), _lib.ptr(mask_u8), B, V,
                _lib.ptr(g_in), model._packed[1], _aligned(ctx.tape), tape_b, _aligned(ws), ws_b, _lib.ptr(dxc), _lib.ptr(dxv),
                _lib.ptr(dz0c), _lib.ptr(dz0v), model._stream(dev),
            ),
            "tw_flow_log_likelihood_backward_inputs",
        )  # fmt: skip
        ctx.tape = None
        g_x = g_xv = g_y = g_yv = None
        if want_x:
            keep = (mask_u8 == 0)[:, :, None].to(torch.float32)
            # x_c = x - mean over the unmasked atoms (molecule_utils.py:15-29): every atom's centred coordinate depends on them
            g_x = dxc - keep * (dxc.sum(1, keepdim=True) / keep.sum(1, keepdim=True).clamp_min(1.0))
            if model.use_displacement_as_target:  # the flow input is y - x (flow.py:148-149)
                g_x = g_x - dz0c
            g_xv = None if model.ignore_conditional_velocity else dxv
        if ctx.needs_input_grad[4]:
            g_y = dz0c
        if ctx.needs_input_grad[5]:
            g_yv = dz0v
        head = (None, None, g_x if ctx.needs_input_grad[2] else None, g_xv if ctx.needs_input_grad[3] else None, g_y, g_yv, None)
        if ctx.n_extra:
            g_log_ls = None
            if ls_grad:  # d/d(log l) = l * d/dl, in place inside the flat buffer (data-parallel training all-reduces it)
                g_log_ls = views[ls_slot].mul_(tensors[ls_slot])
            return head + tuple(grads) + (g_log_ls,)
        return head + tuple(grads)


class _SampleFn(torch.autograd.Function):
    """conditional_sample_with_logp under autograd (the energy-based losses, losses.py:396-664): y and delta = log p(y|x) -
    prior(z) from the taped sampling pass (tw_flow_sample_train); the backward (tw_flow_sample_backward) returns the gradient
    of every trainable parameter and of the latent draws z.  The prior term and the scaling of the draws by exp(log_scale)
    stay in torch, so the prior log-scales get their gradient from autograd."""

    @staticmethod
    def forward(ctx, model, atom_types, x_coords, x_velocs, mask_u8, z_coords, z_velocs, *params):
        ctx.n_extra = 1 if model._learnable else 0
        lib = _lib.load()
        dev = x_coords.device
        B, V = x_coords.shape[0], x_coords.shape[1]
        cfg = model._cfg
        tape_b, ws_b = C.c_size_t(0), C.c_size_t(0)
        _lib.check(lib.tw_flow_train_bytes(C.byref(cfg), B, V, C.byref(tape_b), C.byref(ws_b)), "tw_flow_train_bytes")
        tape = torch.empty(tape_b.value + 1024, dtype=torch.uint8, device=dev)
        y_coords, y_velocs = torch.empty_like(x_coords), torch.empty_like(x_coords)
        delta = torch.empty(B, dtype=torch.float32, device=dev)
        table = model._param_table(dev)
        packed = model._packed_weights(dev, force=True)
        model._stale_after_train = True
        zc, zv = z_coords.detach().contiguous(), z_velocs.detach().contiguous()
        _lib.check(
            lib.tw_flow_sample_train(
                C.byref(cfg), table, _lib.ptr(atom_types), _lib.ptr(x_coords), _lib.ptr(x_velocs), _lib.ptr(mask_u8), B, V,
                model._flags(), _lib.ptr(zc), _lib.ptr(zv), _lib.ptr(y_coords), _lib.ptr(y_velocs), _lib.ptr(delta), packed,
                _aligned(tape), tape_b.value, model._stream(dev),
            ),
            "tw_flow_sample_train",
        )  # fmt: skip
        ctx.model, ctx.tape, ctx.sizes = model, tape, (B, V, tape_b.value, ws_b.value)
        ctx.packed_key = model._packed[2]
        ctx.ls_eff = model._ls_eff.clone() if model._learnable else None  # (see _LogLikelihoodFn.forward)
        ctx.save_for_backward(atom_types, x_velocs, mask_u8)
        return y_coords, y_velocs, delta

    @staticmethod
    def backward(ctx, g_yc, g_yv, g_delta):
        lib = _lib.load()
        model = ctx.model
        atom_types, x_velocs, mask_u8 = ctx.saved_tensors
        dev = x_velocs.device
        B, V, tape_b, ws_b = ctx.sizes
        if ctx.tape is None:
            raise _lib.TimewarpB200Error("backward ran twice over the same graph: the activation tape is released after the first "
                                         "backward (retain_graph=True is not supported)")
        if model._packed is None or (model._packed[2][0], model._packed[2][2]) != (ctx.packed_key[0], ctx.packed_key[2]):
            raise _lib.TimewarpB200Error("parameters were modified between the forward and the backward pass")
        if ctx.ls_eff is not None:
            model._ls_eff.copy_(ctx.ls_eff)
        tensors, views, grads, gtable, ls_slot, ls_grad = _gradient_table(model, dev, ctx.n_extra == 1 and ctx.needs_input_grad[-1])
        table = model._param_table(dev)
        ws = torch.empty(ws_b + 1024, dtype=torch.uint8, device=dev)

        def dense(g, shape):
            if g is None:
                return torch.zeros(shape, dtype=torch.float32, device=dev)
            return g.to(torch.float32).contiguous()

        g_yc, g_yv, g_delta = dense(g_yc, (B, V, 3)), dense(g_yv, (B, V, 3)), dense(g_delta, (B,))
        dzc, dzv = torch.empty(B, V, 3, dtype=torch.float32, device=dev), torch.empty(B, V, 3, dtype=torch.float32, device=dev)
        _lib.check(
            lib.tw_flow_sample_backward(
                C.byref(model._cfg), table, gtable, _lib.ptr(atom_types), _lib.ptr(x_velocs), _lib.ptr(mask_u8), B, V,
                _lib.ptr(g_yc), _lib.ptr(g_yv), _lib.ptr(g_delta), model._packed[1], _aligned(ctx.tape), tape_b, _aligned(ws), ws_b,
                _lib.ptr(dzc), _lib.ptr(dzv), model._stream(dev),
            ),
            "tw_flow_sample_backward",
        )  # fmt: skip
        ctx.tape = None
        out = (None,) * 5 + (dzc, dzv) + tuple(grads)
        if ctx.n_extra:
            out += ((views[ls_slot].mul_(tensors[ls_slot]) if ls_grad else None),)
        return out


def _is_optional_grad(model, i: int) -> bool:
    """Entries of the gradient table that may be NULL: prior log-scales and the lengthscale buffers."""
    if i in (1, 2):
        return True
    if i < 3:
        return False
    cfg = model._cfg
    per_mlp = 2 * (cfg.num_mlp_hidden + 1)
    per_net = 2 * per_mlp + 11 * cfg.num_transformer_layers
    if i >= 3 + cfg.num_coupling_layers * 2 * per_net:
        return True  # trailing cheb_coeffs section
    j = (i - 3) % per_net
    return per_mlp <= j < per_mlp + 11 * cfg.num_transformer_layers and (j - per_mlp) % 11 == 1


class ConditionalFlowDensityModel(nn.Module):
    def __init__(
        self,
        flow: ConditionalSequentialFlow,
        flow_config: _lib.FlowConfig,
        use_displacement_as_target: bool = True,
        scale_requires_grad: bool = True,
        ignore_conditional_velocity: bool = False,
    ):
        super().__init__()
        self.flow = flow
        self.coords_prior_log_scale = nn.Parameter(torch.tensor(0.0), requires_grad=scale_requires_grad)
        self.velocs_prior_log_scale = nn.Parameter(torch.tensor(0.0), requires_grad=scale_requires_grad)
        self.ignore_conditional_velocity = ignore_conditional_velocity
        self.use_displacement_as_target = use_displacement_as_target
        self._cfg = flow_config
        # `learnable_kernel` attention: exp(log_lengthscales) of the first attention layer executed in the pass, written
        # here before every pass; every lengthscale entry of the parameter table points at this buffer (the reference's
        # cache key maps `lengthscales` to 0, so one layer's scores serve the whole pass -- SURVEY.md quirk B)
        self._local = flow_config.attention_type == _lib.TW_ATTENTION_LOCAL
        self._learnable = not self._local and hasattr(
            flow.chain[0].scale_transformer.encoder_layers[0].self_attn.attention, "log_lengthscales")
        self._ls_eff: Optional[Tensor] = None
        self._table = None  # (ctypes array, keep-alive list)
        self._workspace: Optional[Tensor] = None
        self._packed = None  # (buffer, aligned ptr, key): bf16 operand images of the weights (tensor-core precisions)
        self._pack_epoch = 0
        self._last_flat_grad: Optional[Tensor] = None
        self._stale_after_train = False  # a taped forward ran since the last pack: re-pack once on the next pass

    # ---------------------------------------------------------------- plumbing
    @property
    def precision(self) -> str:
        return {v: k for k, v in _lib.PRECISION.items()}[self._cfg.precision]

    def set_precision(self, precision: str) -> "ConditionalFlowDensityModel":
        self._cfg.precision = _lib.PRECISION[precision]
        self._workspace = None
        return self

    def _apply(self, fn, *a, **kw):  # .to()/.cuda()/.float() move the parameters: rebuild the pointer table
        self._table = None
        self._ls_eff = None
        self._workspace = None
        self._packed = None
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._table = None
        return super().load_state_dict(*a, **kw)

    def _ordered_params(self):
        """Tensors in the order of the C-ABI parameter table (include/timewarp_b200.h)."""
        out = [self.flow.atom_embedder.weight, self.coords_prior_log_scale, self.velocs_prior_log_scale]
        if self._learnable:
            first = self.flow.chain[0].scale_transformer.encoder_layers[0].self_attn.attention.log_lengthscales
            if self._ls_eff is None or self._ls_eff.device != first.device:
                self._ls_eff = torch.exp(first.detach()).contiguous()
        for layer in self.flow.chain:
            for block in (layer.scale_transformer, layer.shift_transformer):
                for lin in block.in_mlp.linears():
                    out += [lin.weight, lin.bias]
                for enc in block.encoder_layers:
                    if self._local:  # slots [wv, lengthscales, wo] = [qkv_proj.weight, (ignored), output_proj.weight]
                        out += [enc.self_attn.qkv_proj.weight, enc.norm1.bias, enc.self_attn.output_proj.weight, enc.linear1.weight,
                                enc.linear1.bias, enc.linear2.weight, enc.linear2.bias, enc.norm1.weight, enc.norm1.bias,
                                enc.norm2.weight, enc.norm2.bias]  # fmt: skip
                        continue
                    out += [
                        enc.self_attn.values_proj.weight, self._ls_eff if self._learnable else enc.self_attn.attention.lengthscales,
                        enc.self_attn.attention._out_projection.weight, enc.linear1.weight, enc.linear1.bias,
                        enc.linear2.weight, enc.linear2.bias, enc.norm1.weight, enc.norm1.bias, enc.norm2.weight,
                        enc.norm2.bias,
                    ]  # fmt: skip
                for lin in block.out_mlp.linears():
                    out += [lin.weight, lin.bias]
        if self._cfg.attention_type == _lib.TW_ATTENTION_CHEBYSHEV:  # trailing section of the table (include/timewarp_b200.h)
            for layer in self.flow.chain:
                for block in (layer.scale_transformer, layer.shift_transformer):
                    out += [enc.self_attn.attention.cheb_coeffs for enc in block.encoder_layers]
        return out

    def _param_table(self, device):
        if self._table is None or self._table[2] != device:
            tensors = self._ordered_params()
            lib = _lib.load()
            n = lib.tw_flow_num_params(C.byref(self._cfg))
            if n != len(tensors):
                raise _lib.TimewarpB200Error(f"parameter table mismatch: library expects {n}, module has {len(tensors)}")
            for t in tensors:
                if t.device != device:
                    raise _lib.TimewarpB200Error(
                        f"model parameters are on {t.device} but inputs are on {device}; call model.to(device) first"
                    )
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise TypeError("model parameters must be contiguous float32")
            arr = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
            self._table = (arr, tensors, device)
        return self._table[0]

    def invalidate_packed_weights(self) -> None:
        """Force a re-pack of the bf16 weight images at the next pass.  Needed only after parameter updates that do not
        bump the tensors' version counters (e.g. writes through raw pointers); in-place torch ops are detected, and
        every taped (training) forward re-packs unconditionally."""
        self._pack_epoch += 1

    def _packed_weights(self, device, force: bool = False) -> Optional[int]:
        """Device pointer of the packed bf16 weight images (None for fp32).  Re-packed whenever a parameter was modified
        in place (version counters), the precision changed, or `force` (taped training forward: fused optimizers such as
        torch.optim.Adam(fused=True) update the parameters WITHOUT bumping their version counters)."""
        if self._cfg.precision == _lib.PRECISION["fp32"]:
            return None
        table = self._param_table(device)
        if force or self._stale_after_train:
            self._pack_epoch += 1
            self._stale_after_train = False
        key = (self._cfg.precision, self._pack_epoch, tuple(t._version for t in self._table[1] if t is not self._ls_eff))
        if self._packed is None or self._packed[2] != key or self._packed[0].device != device:
            lib = _lib.load()
            need = C.c_size_t(0)
            _lib.check(lib.tw_flow_packed_bytes(C.byref(self._cfg), C.byref(need)), "tw_flow_packed_bytes")
            buf = self._packed[0] if (self._packed is not None and self._packed[0].numel() >= need.value + 1024
                                      and self._packed[0].device == device) else torch.empty(need.value + 1024, dtype=torch.uint8, device=device)
            aligned = (buf.data_ptr() + 1023) // 1024 * 1024
            _lib.check(lib.tw_flow_pack_weights(C.byref(self._cfg), table, aligned, need.value, self._stream(device)), "tw_flow_pack_weights")
            self._packed = (buf, aligned, key)
        return self._packed[1]

    def _get_workspace(self, n: int, n_cond: int, V: int, device) -> Tuple[Tensor, int]:
        lib = _lib.load()
        need = C.c_size_t(0)
        _lib.check(lib.tw_flow_workspace_bytes(C.byref(self._cfg), n, n_cond, V, C.byref(need)), "tw_flow_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != device:
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=device)
        return self._workspace, self._workspace.numel()

    def _train_supported(self, B: int, V: int) -> bool:
        """True when the library has backward kernels for this configuration (tensor-core precision and the
        flagship layer sizes): tw_flow_train_bytes answers TW_ERR_UNSUPPORTED otherwise."""
        if self._cfg.precision == _lib.PRECISION["fp32"]:
            return False
        a, b = C.c_size_t(0), C.c_size_t(0)
        return _lib.load().tw_flow_train_bytes(C.byref(self._cfg), B, V, C.byref(a), C.byref(b)) == _lib.TW_OK

    def _set_pass_lengthscales(self, reverse: bool) -> None:
        """learnable_kernel: lengthscales of the first attention layer the reference executes in this direction."""
        if not self._learnable:
            return
        layer = self.flow.chain[len(self.flow.chain) - 1 if reverse else 0]
        log_ls = layer.scale_transformer.encoder_layers[0].self_attn.attention.log_lengthscales
        self._ls_eff.copy_(torch.exp(log_ls.detach()))

    @staticmethod
    def _stream(device) -> int:
        return torch.cuda.current_stream(device).cuda_stream

    def _flags(self) -> int:
        return _lib.TW_FLOW_DISPLACEMENT_TARGET if self.use_displacement_as_target else 0

    # ---------------------------------------------------------------- reference API
    def forward(
        self,
        atom_types: Tensor,  # [B, V] int64
        x_coords: Tensor,  # [B, V, 3]
        x_velocs: Tensor,  # [B, V, 3]
        y_coords: Tensor,  # [B, V, 3]
        y_velocs: Tensor,  # [B, V, 3]
        adj_list: Tensor,  # [E, 2] int64 (unused by this model, kept for the interface)
        edge_batch_idx: Tensor,  # [E] int64 (unused)
        masked_elements: BoolTensor,  # [B, V] True = padding
        logger=None,
    ) -> Tensor:
        """Average negative log-likelihood per atom (density_model_base.py:14-47)."""
        num_atoms = (~masked_elements).sum(dim=1)
        log_likelihood = self.log_likelihood(
            atom_types=atom_types, adj_list=adj_list, x_coords=x_coords, x_velocs=x_velocs, y_coords=y_coords,
            y_velocs=y_velocs, edge_batch_idx=edge_batch_idx, masked_elements=masked_elements, logger=logger,
        )  # fmt: skip
        loss = -(log_likelihood / num_atoms).mean()
        if logger is not None:
            logger.log_scalar_async("nll_loss", loss)
        return loss

    def log_likelihood(
        self, atom_types: Tensor, x_coords: Tensor, x_velocs: Tensor, y_coords: Tensor, y_velocs: Tensor,
        adj_list: Tensor, edge_batch_idx: Tensor, masked_elements: BoolTensor, logger=None,
    ) -> Tensor:  # fmt: skip
        """log p(y | x) for each batch element (flow.py:131-215)."""
        ll, _, _ = self._log_likelihood_impl(atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements, False)
        if logger is not None:  # flow.py:206-214 logs means of the pieces; only the total is materialised here
            logger.log_scalar_async("log_prob_y", ll.mean())
            logger.log_scalar_async("coord_std", torch.exp(self.coords_prior_log_scale.detach()))
            logger.log_scalar_async("veloc_std", torch.exp(self.velocs_prior_log_scale.detach()))
        return ll

    def _log_likelihood_impl(self, atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements, want_latent):
        x_coords = _require_cuda("x_coords", x_coords, torch.float32)
        dev = x_coords.device
        B, V = x_coords.shape[0], x_coords.shape[1]
        if x_coords.dim() != 3 or x_coords.shape[2] != 3:
            raise ValueError(f"x_coords must be [B, V, 3], got {tuple(x_coords.shape)}")
        x_velocs = _require_cuda("x_velocs", x_velocs, torch.float32)
        y_coords = _require_cuda("y_coords", y_coords, torch.float32)
        y_velocs = _require_cuda("y_velocs", y_velocs, torch.float32)
        atom_types = _require_cuda("atom_types", atom_types, torch.int64)
        mask = _require_cuda("masked_elements", masked_elements, torch.bool)
        for name, t in (("x_velocs", x_velocs), ("y_coords", y_coords), ("y_velocs", y_velocs)):
            if t.shape != x_coords.shape:
                raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(x_coords.shape)}")
        if atom_types.shape != (B, V) or mask.shape != (B, V):
            raise ValueError("atom_types / masked_elements must be [B, V]")
        if self.ignore_conditional_velocity:  # flow.py:144-145
            x_velocs = torch.zeros_like(x_velocs)
        lib = _lib.load()
        mask_u8 = mask.view(torch.uint8)
        self._param_table(dev)
        self._set_pass_lengthscales(reverse=False)
        trainable = torch.is_grad_enabled() and not want_latent and any(p.requires_grad for p in self.parameters())
        if trainable and (self.training or self._train_supported(B, V)):
            # hand-written backward (tensor-core precisions, flagship layer sizes).  Configurations without backward
            # kernels raise TW_ERR_UNSUPPORTED in .train() mode; in .eval() mode they take the inference path and the
            # result carries no grad_fn (so .backward() on it fails in autograd).
            extra = ()
            if self._learnable:  # the log_lengthscales the pass reads (flow.chain[0]...: the reference's cache quirk)
                extra = (self.flow.chain[0].scale_transformer.encoder_layers[0].self_attn.attention.log_lengthscales,)
            out = _LogLikelihoodFn.apply(self, atom_types, x_coords, x_velocs, y_coords, y_velocs, mask_u8, *self._ordered_params(), *extra)
            return out, None, None
        table = self._param_table(dev)
        ws, ws_bytes = self._get_workspace(B, B, V, dev)
        out = torch.empty(B, dtype=torch.float32, device=dev)
        zc = torch.empty_like(x_coords) if want_latent else None
        zv = torch.empty_like(x_coords) if want_latent else None
        _lib.check(
            lib.tw_flow_log_likelihood(
                C.byref(self._cfg), table, _lib.ptr(atom_types), _lib.ptr(x_coords), _lib.ptr(x_velocs), _lib.ptr(y_coords),
                _lib.ptr(y_velocs), _lib.ptr(mask_u8), B, V, self._flags(), _lib.ptr(out), _lib.ptr(zc), _lib.ptr(zv),
                self._packed_weights(dev), _lib.ptr(ws), ws_bytes, self._stream(dev),
            ),
            "tw_flow_log_likelihood",
        )  # fmt: skip
        return out, zc, zv

    def conditional_sample(
        self, atom_types: Tensor, x_coords: Tensor, x_velocs: Tensor, adj_list: Tensor, edge_batch_idx: Tensor,
        masked_elements: BoolTensor, num_samples: int, logger=None,
    ) -> Tuple[Tensor, Tensor]:  # fmt: skip
        """Conditional samples y ~ p(.|x): ([S,B,V,3], [S,B,V,3])  (flow.py:217-240)."""
        y_coords, y_velocs, _ = self._sample_impl(atom_types, x_coords, x_velocs, masked_elements, num_samples, None, None, False)
        return y_coords, y_velocs

    def conditional_sample_with_logp(
        self, atom_types: Tensor, x_coords: Tensor, x_velocs: Tensor, adj_list: Tensor, edge_batch_idx: Tensor,
        masked_elements: BoolTensor, num_samples: int, logger=None,
    ) -> Tuple[Tensor, Tensor, Tensor]:  # fmt: skip
        """Conditional samples and their log-density: (..., ..., [S,B])  (flow.py:242-336)."""
        return self._sample_impl(atom_types, x_coords, x_velocs, masked_elements, num_samples, None, None, True)

    def sample_from_latents(self, atom_types, x_coords, x_velocs, masked_elements, z_coords: Tensor, z_velocs: Tensor):
        """Same as conditional_sample_with_logp but with the prior draws given ([S,B,V,3], already
        scaled by exp(log_scale)) -- used by parity tests and by drivers that own the RNG."""
        return self._sample_impl(atom_types, x_coords, x_velocs, masked_elements, z_coords.shape[0], z_coords, z_velocs, True)

    def _sample_impl(self, atom_types, x_coords, x_velocs, masked_elements, num_samples, z_coords, z_velocs, want_logp):
        x_coords = _require_cuda("x_coords", x_coords, torch.float32)
        dev = x_coords.device
        if x_coords.dim() != 3 or x_coords.shape[2] != 3:
            raise ValueError(f"x_coords must be [B, V, 3], got {tuple(x_coords.shape)}")
        B, V = x_coords.shape[0], x_coords.shape[1]
        S = int(num_samples)
        x_velocs = _require_cuda("x_velocs", x_velocs, torch.float32)
        atom_types = _require_cuda("atom_types", atom_types, torch.int64)
        mask = _require_cuda("masked_elements", masked_elements, torch.bool)
        if x_velocs.shape != x_coords.shape or atom_types.shape != (B, V) or mask.shape != (B, V):
            raise ValueError("inconsistent input shapes")
        if want_logp and S > 1 and B > 1:
            # the reference's prior-mask broadcast (flow.py:326-331) only works for S == 1 or B == 1
            raise ValueError("conditional_sample_with_logp requires num_samples == 1 or batch size == 1 (flow.py:326-331)")
        if self.ignore_conditional_velocity:
            x_velocs = torch.zeros_like(x_velocs)
        trainable = want_logp and S == 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if trainable and (self.training or self._train_supported(B, V)):
            return self._sample_taped(atom_types, x_coords, x_velocs, mask, z_coords, z_velocs)
        if self.training and want_logp and S > 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # no silent loss of the graph: the differentiable sampler covers num_samples == 1 (how the losses call it)
            raise NotImplementedError("conditional_sample_with_logp under autograd supports num_samples == 1 only; "
                                      "wrap inference calls in torch.no_grad()")
        if z_coords is None:
            # RNG contract (flow.py:274-275): two normal_() draws [S,B,V,3] from the device's default
            # generator, coords first, each scaled by exp(log_scale).
            z_coords = torch.empty(S, B, V, 3, dtype=torch.float32, device=dev).normal_() * torch.exp(self.coords_prior_log_scale.detach())
            z_velocs = torch.empty(S, B, V, 3, dtype=torch.float32, device=dev).normal_() * torch.exp(self.velocs_prior_log_scale.detach())
        else:
            z_coords = _require_cuda("z_coords", z_coords, torch.float32)
            z_velocs = _require_cuda("z_velocs", z_velocs, torch.float32)
            if z_coords.shape != (S, B, V, 3) or z_velocs.shape != (S, B, V, 3):
                raise ValueError("latents must be [S, B, V, 3]")
        lib = _lib.load()
        table = self._param_table(dev)
        self._set_pass_lengthscales(reverse=True)
        ws, ws_bytes = self._get_workspace(S * B, B, V, dev)
        y_coords = torch.empty(S, B, V, 3, dtype=torch.float32, device=dev)
        y_velocs = torch.empty(S, B, V, 3, dtype=torch.float32, device=dev)
        logp = torch.empty(S, B, dtype=torch.float32, device=dev) if want_logp else None
        mask_u8 = mask.view(torch.uint8)
        _lib.check(
            lib.tw_flow_sample(
                C.byref(self._cfg), table, _lib.ptr(atom_types), _lib.ptr(x_coords), _lib.ptr(x_velocs), _lib.ptr(mask_u8),
                B, V, S, self._flags(), _lib.ptr(z_coords), _lib.ptr(z_velocs), _lib.ptr(y_coords), _lib.ptr(y_velocs),
                _lib.ptr(logp), self._packed_weights(dev), _lib.ptr(ws), ws_bytes, self._stream(dev),
            ),
            "tw_flow_sample",
        )  # fmt: skip
        return y_coords, y_velocs, logp

    def _sample_taped(self, atom_types, x_coords, x_velocs, mask, z_coords, z_velocs):
        """S = 1 sampling with a grad_fn (flow.py:242-336 under autograd): parameters, prior log-scales and -- through them --
        the samples and their log-density are differentiable; the conditioning state is not."""
        dev = x_coords.device
        B, V = x_coords.shape[:2]
        if z_coords is None:  # same RNG contract as the inference path; the scaling stays on the autograd tape
            z_coords = torch.empty(1, B, V, 3, dtype=torch.float32, device=dev).normal_() * torch.exp(self.coords_prior_log_scale)
            z_velocs = torch.empty(1, B, V, 3, dtype=torch.float32, device=dev).normal_() * torch.exp(self.velocs_prior_log_scale)
        elif z_coords.shape != (1, B, V, 3) or z_velocs.shape != (1, B, V, 3):
            raise ValueError("latents must be [1, B, V, 3]")
        self._param_table(dev)
        self._set_pass_lengthscales(reverse=True)
        extra = ()
        if self._learnable:  # the log_lengthscales the sampling pass reads: first attention layer of the LAST coupling layer
            extra = (self.flow.chain[len(self.flow.chain) - 1].scale_transformer.encoder_layers[0].self_attn.attention.log_lengthscales,)
        zc, zv = z_coords[0], z_velocs[0]
        y_coords, y_velocs, delta = _SampleFn.apply(self, atom_types, x_coords, x_velocs, mask.view(torch.uint8), zc, zv,
                                                    *self._ordered_params(), *extra)
        keep = (~mask)[:, :, None].to(torch.float32)

        def prior(z, log_scale):  # Normal(0, exp(log_scale)).log_prob summed over the unmasked atoms (flow.py:322-334)
            return ((-0.5 * (z * torch.exp(-log_scale)) ** 2 - log_scale - 0.5 * math.log(2.0 * math.pi)) * keep).sum((-1, -2))

        logp = prior(zc, self.coords_prior_log_scale) + prior(zv, self.velocs_prior_log_scale) + delta
        return y_coords[None], y_velocs[None], logp[None]

    # ---------------------------------------------------------------- test / debug hooks
    def attention_scores(self, x_coords_centred: Tensor, masked_elements: Tensor) -> Tensor:
        """compute_kernel_attention_scores (kernel_attention.py:69-121) -> [B,H,V,V]."""
        if self._local:
            raise TypeError("`local` attention has no position-only scores (dot-product attention, local_self_attention.py:99-100)")
        x = _require_cuda("x_coords", x_coords_centred, torch.float32)
        mask = _require_cuda("masked_elements", masked_elements, torch.bool).view(torch.uint8)
        att = self.flow.chain[0].scale_transformer.encoder_layers[0].self_attn.attention
        ls = torch.exp(att.log_lengthscales.detach()).contiguous() if self._learnable else att.lengthscales
        B, V = x.shape[:2]
        out = torch.empty(B, ls.numel(), V, V, dtype=torch.float32, device=x.device)
        _lib.check(
            _lib.load().tw_attn_scores(_lib.ptr(x), _lib.ptr(mask), _lib.ptr(ls), B, V, ls.numel(), _lib.ptr(out), self._stream(x.device)),
            "tw_attn_scores",
        )
        return out

    def scale_and_shift(self, layer_idx, atom_types, z_coords, z_velocs, x_coords_centred, x_velocs, masked_elements):
        """NVPCouplingLayer._get_scale_and_shift of one coupling layer (custom_transformer_nvp.py:44-93)."""
        x = _require_cuda("x_coords", x_coords_centred, torch.float32)
        dev = x.device
        B, V = x.shape[:2]
        args = [_require_cuda(n, t, torch.float32) for n, t in (("x_velocs", x_velocs), ("z_coords", z_coords), ("z_velocs", z_velocs))]
        at = _require_cuda("atom_types", atom_types, torch.int64)
        mask = _require_cuda("masked_elements", masked_elements, torch.bool).view(torch.uint8)
        ws, ws_bytes = self._get_workspace(B, B, V, dev)
        self._param_table(dev)
        self._set_pass_lengthscales(reverse=False)
        scale, shift = torch.empty_like(x), torch.empty_like(x)
        _lib.check(
            _lib.load().tw_flow_scale_shift(
                C.byref(self._cfg), self._param_table(dev), int(layer_idx), _lib.ptr(at), _lib.ptr(x), _lib.ptr(args[0]),
                _lib.ptr(args[1]), _lib.ptr(args[2]), _lib.ptr(mask), B, V, _lib.ptr(scale), _lib.ptr(shift),
                self._packed_weights(dev), _lib.ptr(ws), ws_bytes, self._stream(dev),
            ),
            "tw_flow_scale_shift",
        )  # fmt: skip
        return scale, shift

"""nn.Module tree of the kernel-attention RealNVP flow.

The sub-modules are PARAMETER CONTAINERS laid out exactly like the reference's modules so that
`state_dict()` keys, `.parameters()` order and seeded initialisation match the reference
(SURVEY.md section 2.1): a reference checkpoint loads verbatim and vice versa.  They do not
implement `forward` -- all arithmetic runs in the CUDA library through the whole-pass entry
points of `ConditionalFlowDensityModel` (timewarp_b200/flow.py).
"""
from __future__ import annotations

from typing import List, Literal, Sequence

import torch
import torch.nn as nn

ELEMENT_VOCAB_SIZE = 5  # C,H,N,O,S (dataloader.py:24-25)


class _Container(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover
        raise NotImplementedError(
            f"{type(self).__name__} is a parameter container; compute goes through "
            "ConditionalFlowDensityModel.{forward,log_likelihood,conditional_sample*} (fused CUDA path, no per-module fallback)"
        )


class MLP(_Container):
    """modules/layers/mlp.py:6-26 -- Linear -> SiLU per hidden dim, final Linear."""

    def __init__(self, input_dim: int, out_dim: int, hidden_layer_dims: Sequence[int]):
        super().__init__()
        layers: List[nn.Module] = []
        cur = input_dim
        for h in hidden_layer_dims:
            layers.append(nn.Linear(cur, h))
            layers.append(nn.SiLU())
            cur = h
        layers.append(nn.Linear(cur, out_dim))
        self._layers = nn.Sequential(*layers)

    def linears(self) -> List[nn.Linear]:
        return [m for m in self._layers if isinstance(m, nn.Linear)]


class KernelAttention(_Container):
    """modules/layers/kernel_attention.py:159-214 (persistent `lengthscales` buffer + bias-free out projection)."""

    def __init__(self, *, value_dim: int, output_dim: int, lengthscales: Sequence[float], normalise_kernel_values: bool):
        super().__init__()
        self.register_buffer("lengthscales", torch.tensor(lengthscales, dtype=torch.float32), persistent=True)
        # Quirk A (kernel_attention.py:176 vs :197-206): the flag is stored but never forwarded -> scores are always normalised.
        self.normalise_kernel_values = normalise_kernel_values
        self._out_projection = nn.Linear(value_dim * len(lengthscales), output_dim, bias=False)


class LearnableLengthscaleKernelAttention(KernelAttention):
    """modules/layers/kernel_attention.py:217-253: lengthscales = exp(log_lengthscales), a Parameter per layer.  Same
    construction order as the reference (buffer + projection in the base class, the Parameter, a SECOND projection that
    replaces the first) so that seeded initialisation consumes the RNG identically."""

    def __init__(self, *, value_dim: int, output_dim: int, lengthscales: Sequence[float], normalise_kernel_values: bool):
        super().__init__(value_dim=value_dim, output_dim=output_dim, lengthscales=lengthscales,
                         normalise_kernel_values=normalise_kernel_values)
        self.log_lengthscales = nn.Parameter(torch.log(torch.tensor(lengthscales, dtype=torch.float32)), requires_grad=True)
        self._out_projection = nn.Linear(value_dim * len(lengthscales), output_dim, bias=False)


# Chebyshev-rational coefficients of exp(-s): the initial value of `cheb_coeffs` (kernel_attention.py:291-327; numerical
# constants of the reference, data not code)
CHEB_COEFFS_EXPMX = [
    4.275836e-01, -5.464240e-01, 7.106222e-02, 5.473271e-02, 5.744192e-03, -7.926410e-03, -5.392865e-03, -1.210823e-03,
    6.996851e-04, 8.686655e-04, 4.459163e-04, 7.084817e-05, -9.620444e-05, -1.110469e-04, -6.551055e-05, -1.875292e-05,
    7.930955e-06, 1.553729e-05, 1.246072e-05, 6.282442e-06, 1.216243e-06, -1.468327e-06, -2.141963e-06, -1.694741e-06,
    -9.063254e-07, -2.337215e-07, 1.609271e-07, 2.978384e-07, 2.700519e-07, 1.730454e-07, 7.272222e-08, 1.192814e-09,
]  # fmt: skip


class LearnableChebyshevKernelAttention(KernelAttention):
    """modules/layers/kernel_attention.py:255-339: the basis function is a learnable Chebyshev-rational expansion
    `sum_c cheb_coeffs[h, c] R_c((d / l_h)^2)`.  Same state-dict key and shape as the reference ([H, cheb_order]); the
    reference builds the Parameter as an EXPANDED tensor whose heads share storage -- here every head owns its row."""

    def __init__(self, *, value_dim: int, output_dim: int, lengthscales: Sequence[float], cheb_order: int,
                 normalise_kernel_values: bool, force_asymptotic_zero: bool):
        assert cheb_order >= 1
        super().__init__(value_dim=value_dim, output_dim=output_dim, lengthscales=lengthscales,
                         normalise_kernel_values=normalise_kernel_values)
        take = min(len(CHEB_COEFFS_EXPMX), cheb_order)
        coeffs = torch.tensor(CHEB_COEFFS_EXPMX[:take] + [0.0] * max(0, cheb_order - len(CHEB_COEFFS_EXPMX)), dtype=torch.float32)
        self.cheb_coeffs = nn.Parameter(coeffs[None, :].repeat(len(lengthscales), 1), requires_grad=True)
        self.cheb_order, self.force_asymptotic_zero = cheb_order, force_asymptotic_zero
        self._out_projection = nn.Linear(value_dim * len(lengthscales), output_dim, bias=False)


class KernelSelfAttention(_Container):
    """modules/layers/kernel_self_attention.py:12-48."""

    def __init__(self, *, input_dim: int, num_heads: int, value_dim: int, attention: KernelAttention):
        super().__init__()
        self.num_heads, self.input_dim, self.value_dim = num_heads, input_dim, value_dim
        self.values_proj = nn.Linear(input_dim, num_heads * value_dim, bias=False)
        self.attention = attention


class LocalSelfAttention(_Container):
    """modules/layers/local_self_attention.py:14-119: dot-product attention over the atoms within `max_radius` of each atom
    (bias-free fused q|k|v projection per head, bias-free output projection; same construction order as the reference)."""

    def __init__(self, *, input_dim: int, output_dim: int, num_heads: int, value_dim: int, key_query_dim: int, max_radius: float):
        super().__init__()
        self.num_heads, self.value_dim, self.key_query_dim, self.max_radius = num_heads, value_dim, key_query_dim, max_radius
        self.qkv_proj = nn.Linear(input_dim, num_heads * (value_dim + 2 * key_query_dim), bias=False)
        self.output_proj = nn.Linear(num_heads * value_dim, output_dim, bias=False)


class CustomTransformerEncoderLayer(_Container):
    """modules/layers/custom_attention_encoder.py:24-114 (post-LN, ReLU FFN, dropout must be 0)."""

    def __init__(self, *, d_model: int, self_attention: KernelSelfAttention, dim_feedforward: int, layer_norm_eps: float = 1e-5):
        super().__init__()
        self.d_model, self.dim_feedforward = d_model, dim_feedforward
        self.self_attn = self_attention
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)


def custom_attention_transformer_encoder_constructor(config) -> CustomTransformerEncoderLayer:
    """modules/layers/custom_attention_encoder.py:140-219: `local`, `kernel`, `learnable_kernel` and `chebyshev_kernel`."""
    if config.attention_type not in ("local", "kernel", "learnable_kernel", "chebyshev_kernel"):
        raise RuntimeError(f"Unknown attention type: {config.attention_type}.")  # custom_attention_encoder.py:211-212
    if float(config.dropout) != 0.0:
        raise NotImplementedError("dropout must be 0 (configs/kernel_transformer_nvp.yaml:27); no dropout kernel exists")
    if config.attention_type == "local":
        assert config.max_radius is not None
        local = LocalSelfAttention(input_dim=config.d_model, output_dim=config.d_model, num_heads=config.num_heads,
                                   value_dim=config.d_model, key_query_dim=config.d_model, max_radius=config.max_radius)
        return CustomTransformerEncoderLayer(d_model=config.d_model, self_attention=local, dim_feedforward=config.dim_feedforward)
    assert config.lengthscales is not None
    assert len(config.lengthscales) > 0
    assert config.normalise_kernel_values is not None
    # construction order == reference (attention first) so that seeded init is identical
    if config.attention_type == "chebyshev_kernel":
        assert config.cheb_order is not None and config.cheb_order >= 1
        assert config.force_asymptotic_zero is not None
        attention = LearnableChebyshevKernelAttention(
            value_dim=config.d_model, output_dim=config.d_model, lengthscales=list(config.lengthscales),
            cheb_order=config.cheb_order, normalise_kernel_values=config.normalise_kernel_values,
            force_asymptotic_zero=config.force_asymptotic_zero)
    else:
        attention_cls = {"kernel": KernelAttention, "learnable_kernel": LearnableLengthscaleKernelAttention}[config.attention_type]
        attention = attention_cls(
            value_dim=config.d_model,
            output_dim=config.d_model,
            lengthscales=list(config.lengthscales),
            normalise_kernel_values=config.normalise_kernel_values,
        )
    self_attention = KernelSelfAttention(
        input_dim=config.d_model, num_heads=len(config.lengthscales), value_dim=config.d_model, attention=attention
    )
    return CustomTransformerEncoderLayer(
        d_model=config.d_model, self_attention=self_attention, dim_feedforward=config.dim_feedforward
    )


class CustomAttentionTransformerBlock(_Container):
    """modules/layers/custom_transformer_block.py:15-82."""

    def __init__(self, input_dim: int, output_dim: int, mlp_hidden_layer_dims: List[int],
                 transformer_encoder_layers: Sequence[CustomTransformerEncoderLayer]):
        super().__init__()
        self.in_mlp = MLP(input_dim=input_dim, hidden_layer_dims=mlp_hidden_layer_dims, out_dim=transformer_encoder_layers[0].d_model)
        self.encoder_layers = nn.ModuleList(transformer_encoder_layers)
        self.out_mlp = MLP(input_dim=transformer_encoder_layers[-1].d_model, hidden_layer_dims=mlp_hidden_layer_dims, out_dim=output_dim)


class CustomAttentionTransformerCouplingLayer(_Container):
    """modules/custom_transformer_nvp.py:14-93 + modules/layers/nvp.py:13-205."""

    def __init__(self, atom_embedding_dim: int, mlp_hidden_layer_dims: List[int],
                 transformed_vars: Literal["positions", "velocities"],
                 scale_transformer_encoder_layers: List[CustomTransformerEncoderLayer],
                 shift_transformer_encoder_layers: List[CustomTransformerEncoderLayer]):
        super().__init__()
        self.transformed_vars = transformed_vars
        self.scale_transformer = CustomAttentionTransformerBlock(
            input_dim=atom_embedding_dim + 9, output_dim=3, mlp_hidden_layer_dims=mlp_hidden_layer_dims,
            transformer_encoder_layers=scale_transformer_encoder_layers)
        self.shift_transformer = CustomAttentionTransformerBlock(
            input_dim=atom_embedding_dim + 9, output_dim=3, mlp_hidden_layer_dims=mlp_hidden_layer_dims,
            transformer_encoder_layers=shift_transformer_encoder_layers)


class ConditionalSequentialFlow(_Container):
    """modules/model_wrappers/flow.py:44-103."""

    def __init__(self, layers: Sequence[nn.Module], atom_embedder: nn.Module):
        super().__init__()
        self.atom_embedder = atom_embedder
        self.chain = nn.ModuleList(layers)

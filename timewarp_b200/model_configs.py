"""Config dataclasses of the hot-path model.  Field names are the reference's (the drop-in's
config contract): model_configs.py:61-69, modules/layers/custom_attention_encoder.py:126-137,
modules/model_wrappers/flow.py:339-347.  `precision` is the only addition."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class ConditionalFlowDensityConfig:
    scale_requires_grad: bool = True
    ignore_conditional_velocity: bool = False
    use_displacement_as_target: bool = True


@dataclass
class CustomAttentionEncoderLayerConfig:
    d_model: int  # Dimension of values in self attention
    dim_feedforward: int  # Dimension of hidden layer in the pointwise MLP in transformer block
    dropout: float  # Dropout rate in transformer block (must be 0: configs/kernel_transformer_nvp.yaml:27)
    num_heads: int  # Number of heads in multihead attention.
    attention_type: str  # "kernel" | "learnable_kernel" | "chebyshev_kernel" | "local"
    lengthscales: Optional[List[float]] = None
    max_radius: Optional[float] = None
    normalise_kernel_values: Optional[bool] = None
    cheb_order: Optional[int] = None
    force_asymptotic_zero: Optional[bool] = None


@dataclass
class CustomAttentionTransformerNVPConfig:
    atom_embedding_dim: int
    latent_mlp_hidden_dims: List[int]  # MLP that maps from physical space to latent space and back
    num_coupling_layers: int  # Number of coupling layers in RealNVP
    num_transformer_layers: int  # Number of transformer encoder layers per coupling layer
    encoder_layer_config: CustomAttentionEncoderLayerConfig
    position_layer_index_mod_2: int = 0
    conditional_flow_density: ConditionalFlowDensityConfig = field(default_factory=ConditionalFlowDensityConfig)
    # B200 addition: arithmetic of the token-wise GEMMs ("fp32" | "bf16x3" | "bf16"), see DESIGN.md
    precision: str = "bf16x3"


def kernel_transformer_nvp_config(precision: str = "bf16x3") -> CustomAttentionTransformerNVPConfig:
    """configs/kernel_transformer_nvp.yaml:19-30 -- the config every BASELINE workload uses."""
    return CustomAttentionTransformerNVPConfig(
        atom_embedding_dim=32,
        latent_mlp_hidden_dims=[256],
        num_coupling_layers=8,
        num_transformer_layers=3,
        encoder_layer_config=CustomAttentionEncoderLayerConfig(
            d_model=128,
            dim_feedforward=2048,
            dropout=0.0,
            num_heads=6,
            attention_type="kernel",
            lengthscales=[0.1, 0.2, 0.5, 0.7, 1.0, 1.2],
            normalise_kernel_values=True,
        ),
        precision=precision,
    )

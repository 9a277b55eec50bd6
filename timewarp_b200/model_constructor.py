"""custom_transformer_nvp_constructor -- drop-in for model_constructor.py:153-197."""
from __future__ import annotations

import torch.nn as nn

from . import _lib
from .flow import ConditionalFlowDensityModel
from .model_configs import ConditionalFlowDensityConfig, CustomAttentionTransformerNVPConfig
from .modules import (
    ELEMENT_VOCAB_SIZE,
    ConditionalSequentialFlow,
    CustomAttentionTransformerCouplingLayer,
    custom_attention_transformer_encoder_constructor,
)


def flow_config_from(config, precision=None) -> _lib.FlowConfig:
    enc = config.encoder_layer_config
    hidden = list(config.latent_mlp_hidden_dims)
    if len(hidden) > _lib.TW_MAX_MLP_HIDDEN:
        raise ValueError(f"at most {_lib.TW_MAX_MLP_HIDDEN} hidden MLP layers are supported")
    local = enc.attention_type == "local"
    num_heads = enc.num_heads if local else len(enc.lengthscales)
    if num_heads > _lib.TW_MAX_HEADS:
        raise ValueError(f"at most {_lib.TW_MAX_HEADS} heads are supported")
    if not local and enc.num_heads != len(enc.lengthscales):
        import warnings

        warnings.warn(  # custom_attention_encoder.py:158-161: the lengthscales win
            f"Number of lengthscales ({len(enc.lengthscales)}) not equal number of heads of the transformer "
            f"({enc.num_heads}). Using {len(enc.lengthscales)} heads instead."
        )
    prec = precision or getattr(config, "precision", "bf16x3")
    c = _lib.FlowConfig()
    c.atom_embedding_dim = config.atom_embedding_dim
    c.num_mlp_hidden = len(hidden)
    for i, h in enumerate(hidden):
        c.mlp_hidden_dims[i] = h
    c.num_coupling_layers = config.num_coupling_layers
    c.num_transformer_layers = config.num_transformer_layers
    c.d_model = enc.d_model
    c.dim_feedforward = enc.dim_feedforward
    c.num_heads = num_heads
    c.position_layer_index_mod_2 = config.position_layer_index_mod_2
    c.num_atom_types = ELEMENT_VOCAB_SIZE
    c.layer_norm_eps = 1e-5
    c.precision = _lib.PRECISION[prec]
    if local:
        assert enc.max_radius is not None and enc.max_radius > 0
        c.attention_type = _lib.TW_ATTENTION_LOCAL
        c.max_radius = float(enc.max_radius)
    elif enc.attention_type == "chebyshev_kernel":
        assert enc.cheb_order is not None and enc.cheb_order >= 1 and enc.force_asymptotic_zero is not None
        c.attention_type = _lib.TW_ATTENTION_CHEBYSHEV
        c.cheb_order = int(enc.cheb_order)
        c.force_asymptotic_zero = int(bool(enc.force_asymptotic_zero))
    else:
        c.attention_type = _lib.TW_ATTENTION_KERNEL
    return c


def custom_transformer_nvp_constructor(config: CustomAttentionTransformerNVPConfig, precision=None) -> ConditionalFlowDensityModel:
    assert config.num_coupling_layers % 2 == 0, "Real NVP should have an even number of coupling layers"
    position_mod_index = config.position_layer_index_mod_2
    assert position_mod_index == 0 or position_mod_index == 1, "positions_layer_index can only be 0 or 1"

    # same construction order as the reference => identical parameters for identical torch seeds
    coupling_layers = [
        CustomAttentionTransformerCouplingLayer(
            atom_embedding_dim=config.atom_embedding_dim,
            mlp_hidden_layer_dims=list(config.latent_mlp_hidden_dims),
            transformed_vars="positions" if layer_idx % 2 == position_mod_index else "velocities",
            scale_transformer_encoder_layers=[
                custom_attention_transformer_encoder_constructor(config.encoder_layer_config)
                for _ in range(config.num_transformer_layers)
            ],
            shift_transformer_encoder_layers=[
                custom_attention_transformer_encoder_constructor(config.encoder_layer_config)
                for _ in range(config.num_transformer_layers)
            ],
        )
        for layer_idx in range(config.num_coupling_layers)
    ]
    atom_embedder = nn.Embedding(num_embeddings=ELEMENT_VOCAB_SIZE, embedding_dim=config.atom_embedding_dim)
    flow = ConditionalSequentialFlow(layers=coupling_layers, atom_embedder=atom_embedder)
    cfd = getattr(config, "conditional_flow_density", None) or ConditionalFlowDensityConfig()
    return ConditionalFlowDensityModel(
        flow=flow,
        flow_config=flow_config_from(config, precision),
        use_displacement_as_target=cfd.use_displacement_as_target,
        scale_requires_grad=cfd.scale_requires_grad,
        ignore_conditional_velocity=cfd.ignore_conditional_velocity,
    )

"""timewarp_b200 -- B200-native conditional-sampling hot path of microsoft/timewarp.

Public surface (mirrors the reference's):
    custom_transformer_nvp_constructor(config) -> ConditionalFlowDensityModel   (model_constructor.py:153-197)
    ConditionalFlowDensityModel.{forward, log_likelihood, conditional_sample, conditional_sample_with_logp}
    PeptidePotentialEnergy  (OpenmmPotentialEnergyTorch-shaped energy callable, utils/openmm/openmm_bridge.py:252-307)
    sample_with_model / explore  (utils/evaluation_utils.py:468-745, exploration.py:229-250)
"""
from .model_configs import (  # noqa: F401
    ConditionalFlowDensityConfig,
    CustomAttentionEncoderLayerConfig,
    CustomAttentionTransformerNVPConfig,
    kernel_transformer_nvp_config,
)
from .model_constructor import custom_transformer_nvp_constructor  # noqa: F401
from .flow import ConditionalFlowDensityModel  # noqa: F401

__all__ = [
    "ConditionalFlowDensityConfig",
    "CustomAttentionEncoderLayerConfig",
    "CustomAttentionTransformerNVPConfig",
    "kernel_transformer_nvp_config",
    "custom_transformer_nvp_constructor",
    "ConditionalFlowDensityModel",
]

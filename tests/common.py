"""Helpers shared by the test files."""
import os

import numpy as np
import torch

from oracle import flow_oracle as fo
import timewarp_b200 as tw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TINY_O = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4,
                         num_transformer_layers=2, d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0])
FULL_O = fo.OracleConfig()
# `learnable_kernel` attention (per-layer log_lengthscales)
TINY_L = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                         d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="learnable_kernel")
FULL_L = fo.OracleConfig(attention_type="learnable_kernel")
# `chebyshev_kernel` attention (a Chebyshev-rational basis per attention layer, no score sharing)
TINY_C = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                         d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="chebyshev_kernel", cheb_order=6,
                         force_asymptotic_zero=True)
FULL_C = fo.OracleConfig(attention_type="chebyshev_kernel", cheb_order=12, force_asymptotic_zero=False)
# `local` attention: dot-product attention over the atoms within max_radius (modules/layers/local_self_attention.py)
TINY_LOC = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                           d_model=16, dim_feedforward=32, lengthscales=[], attention_type="local", max_radius=0.3, num_heads=3)
FULL_LOC = fo.OracleConfig(lengthscales=[], attention_type="local", max_radius=0.45, num_heads=6)


def model_config(o: fo.OracleConfig, precision: str):
    if getattr(o, "attention_type", "kernel") == "local":
        enc = tw.CustomAttentionEncoderLayerConfig(d_model=o.d_model, dim_feedforward=o.dim_feedforward, dropout=0.0,
                                                   num_heads=o.num_heads, attention_type="local", max_radius=o.max_radius)
        return tw.CustomAttentionTransformerNVPConfig(
            atom_embedding_dim=o.atom_embedding_dim, latent_mlp_hidden_dims=list(o.latent_mlp_hidden_dims),
            num_coupling_layers=o.num_coupling_layers, num_transformer_layers=o.num_transformer_layers, encoder_layer_config=enc,
            position_layer_index_mod_2=o.position_layer_index_mod_2, precision=precision)
    return tw.CustomAttentionTransformerNVPConfig(
        atom_embedding_dim=o.atom_embedding_dim,
        latent_mlp_hidden_dims=list(o.latent_mlp_hidden_dims),
        num_coupling_layers=o.num_coupling_layers,
        num_transformer_layers=o.num_transformer_layers,
        encoder_layer_config=tw.CustomAttentionEncoderLayerConfig(
            d_model=o.d_model, dim_feedforward=o.dim_feedforward, dropout=0.0, num_heads=len(o.lengthscales),
            attention_type=getattr(o, "attention_type", "kernel"), lengthscales=list(o.lengthscales), normalise_kernel_values=True,
            cheb_order=(o.cheb_order if getattr(o, "attention_type", "kernel") == "chebyshev_kernel" else None),
            force_asymptotic_zero=(o.force_asymptotic_zero if getattr(o, "attention_type", "kernel") == "chebyshev_kernel" else None)),
        position_layer_index_mod_2=o.position_layer_index_mod_2,
        precision=precision,
    )


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def build_model(o: fo.OracleConfig, precision: str, weight_seed: int = 0, device="cuda"):
    m = tw.custom_transformer_nvp_constructor(model_config(o, precision))
    sd = fo.synth_state_dict(o, weight_seed)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval(), sd


EMPTY_ADJ = torch.zeros(0, 2, dtype=torch.long)
EMPTY_EBI = torch.zeros(0, dtype=torch.long)


def proposal_weights(sd, n_hidden: int = 1):
    """The bench's "proposal" weight set (bench.py::bench_state_dict): the same synthetic tensors with the last layer of every
    out_mlp scaled by 1e-5 and the prior scales (5e-4 nm, 1) -- a near-identity flow whose proposals are local moves, so both
    branches of the MH rule are exercised."""
    sd = dict(sd)
    last = 2 * n_hidden
    for k in list(sd):
        if f".out_mlp._layers.{last}." in k:
            sd[k] = sd[k] * 1e-5
    sd["coords_prior_log_scale"] = torch.tensor(float(np.log(5e-4)))
    sd["velocs_prior_log_scale"] = torch.tensor(0.0)
    return sd


class CudaReplayDraws:
    """`draws` object for oracle/mh_oracle.py that takes every draw from the CUDA default generator (same calls, same
    shapes, same order as the product code), so that after `torch.manual_seed(s)` the oracle replays the stream the product
    consumed after the same seed."""

    def randn(self, shape):
        return torch.randn(tuple(shape), device="cuda").cpu()

    def rand(self, n):
        return torch.rand(n, device="cuda").cpu()

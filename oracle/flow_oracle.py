"""CPU oracle for the Timewarp kernel-attention RealNVP flow.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain CPU torch, of the reference's conditional
flow (`custom_transformer_nvp`).  It exists so that the CUDA path can be checked on a box
that does not have `/root/reference`.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it; the product package
`timewarp_b200` never does.

Pinning: `tests/golden/make_golden.py` imports the *unmodified* reference from
`/root/reference`, runs it on seeded inputs and commits inputs + outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this file against those vectors.

Every function cites the reference lines it restates (paths relative to the reference root).
The state-dict key layout is the reference's own.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor


@dataclass
class OracleConfig:
    """Sizes of the flow; mirrors CustomAttentionTransformerNVPConfig (model_configs.py:61-69)
    + CustomAttentionEncoderLayerConfig (modules/layers/custom_attention_encoder.py:126-137)."""

    atom_embedding_dim: int = 32
    latent_mlp_hidden_dims: List[int] = field(default_factory=lambda: [256])
    num_coupling_layers: int = 8
    num_transformer_layers: int = 3
    d_model: int = 128
    dim_feedforward: int = 2048
    lengthscales: List[float] = field(default_factory=lambda: [0.1, 0.2, 0.5, 0.7, 1.0, 1.2])
    position_layer_index_mod_2: int = 0
    layer_norm_eps: float = 1e-5
    # "kernel" | "learnable_kernel" (LearnableLengthscaleKernelAttention, kernel_attention.py:217-253)
    # | "chebyshev_kernel" (LearnableChebyshevKernelAttention, :255-339)
    attention_type: str = "kernel"
    cheb_order: int = 0
    force_asymptotic_zero: bool = False
    # "local" (LocalSelfAttention, modules/layers/local_self_attention.py): dot-product attention over the atoms within
    # max_radius; num_heads heads (no lengthscales)
    max_radius: float = 0.0
    num_heads: int = 0


StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# utils/molecule_utils.py:15-29
def centre_of_mass(coords: Tensor, masked_elements: Tensor) -> Tensor:
    inv_mask = ~masked_elements
    c = inv_mask.unsqueeze(-1) * coords
    num_points = inv_mask.sum(dim=-1, keepdim=True).unsqueeze(-1)
    return c.sum(dim=-2, keepdim=True) / num_points


# Chebyshev-rational coefficients of exp(-s), the initial value of every layer's `cheb_coeffs`
# (kernel_attention.py:291-327; numerical-quadrature constants of the reference, data not code).
CHEB_COEFFS_EXPMX = [
    4.275836e-01, -5.464240e-01, 7.106222e-02, 5.473271e-02, 5.744192e-03, -7.926410e-03, -5.392865e-03, -1.210823e-03,
    6.996851e-04, 8.686655e-04, 4.459163e-04, 7.084817e-05, -9.620444e-05, -1.110469e-04, -6.551055e-05, -1.875292e-05,
    7.930955e-06, 1.553729e-05, 1.246072e-05, 6.282442e-06, 1.216243e-06, -1.468327e-06, -2.141963e-06, -1.694741e-06,
    -9.063254e-07, -2.337215e-07, 1.609271e-07, 2.978384e-07, 2.700519e-07, 1.730454e-07, 7.272222e-08, 1.192814e-09,
]  # fmt: skip


# kernel_attention.py:13-66: F(x) = sum_c coeff[h, c] R_c(x^2), R_n(y) = T_n((y - 1) / (y + 1)) by the three-term recursion
def chebyshev_basis(scaled: Tensor, coeffs: Tensor, force_asymptotic_zero: bool) -> Tensor:
    if force_asymptotic_zero:
        coeffs = coeffs - coeffs.mean(dim=1, keepdim=True)  # :27-28
    y = scaled**2
    order = coeffs.shape[1]
    rprev = torch.ones_like(y)
    rfactor = (y - 1.0) / (y + 1.0)
    rcur = rfactor
    terms = [rprev] + ([rcur] if order >= 2 else [])
    for _ in range(2, order):
        rnext = 2.0 * rfactor * rcur - rprev
        terms.append(rnext)
        rcur, rprev = rnext, rcur
    cheb = torch.stack(terms, dim=2)  # [B, H, order, Q, M]
    return torch.einsum("bhcqm,hc->bhqm", cheb, coeffs.to(y.dtype))  # :34


# modules/layers/kernel_attention.py:9-10,69-121
def kernel_attention_scores(
    positions: Tensor,  # [B, V, 3]
    masked_elements: Tensor,  # [B, V] bool, True = padding
    lengthscales: Tensor,  # [H]
    distance_mode: str = "cdist",
    cheb_coeffs: Optional[Tensor] = None,  # [H, order]: Chebyshev basis instead of the Gaussian
    force_asymptotic_zero: bool = False,
) -> Tensor:  # [B, H, V, V]
    if distance_mode == "cdist":
        # literal reference call (kernel_attention.py:98-102)
        d = torch.cdist(positions, positions, compute_mode="use_mm_for_euclid_dist_if_necessary")
    elif distance_mode in ("direct", "direct_sq"):
        diff = positions[:, :, None, :] - positions[:, None, :, :]
        d = torch.sqrt((diff * diff).sum(-1))
    else:
        raise ValueError(distance_mode)
    scaled = d.unsqueeze(-3) / lengthscales[None, :, None, None]  # :105-110
    if cheb_coeffs is not None:
        w = chebyshev_basis(scaled, cheb_coeffs, force_asymptotic_zero)
    elif distance_mode == "direct_sq":
        # same value without the square root: torch's sqrt has an infinite derivative at the zero self-distances, which turns
        # the gradient w.r.t. the positions into NaN; used by the tests that differentiate w.r.t. the conditioning state
        w = torch.exp(-((diff * diff).sum(-1).unsqueeze(-3) / lengthscales[None, :, None, None] ** 2))
    else:
        w = torch.exp(-(scaled**2))  # gaussian_basis_function :9-10
    w = w.masked_fill(masked_elements[:, None, None, :], 0.0)  # :114
    w = w / (torch.abs(w).sum(dim=-1, keepdim=True) + 1e-5)  # :116-119
    return w


# modules/layers/mlp.py:6-26  (Linear -> SiLU per hidden dim, final Linear)
def mlp(sd: StateDict, prefix: str, n_hidden: int, x: Tensor) -> Tensor:
    for i in range(n_hidden):
        x = torch.nn.functional.linear(
            x, sd[f"{prefix}._layers.{2 * i}.weight"], sd[f"{prefix}._layers.{2 * i}.bias"]
        )
        x = torch.nn.functional.silu(x)
    return torch.nn.functional.linear(
        x, sd[f"{prefix}._layers.{2 * n_hidden}.weight"], sd[f"{prefix}._layers.{2 * n_hidden}.bias"]
    )


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps)


# modules/layers/custom_attention_encoder.py:82-114 + kernel_self_attention.py:29-48
# + kernel_attention.py:124-156,185-214
# modules/layers/local_self_attention.py:46-119.  The reference gathers the (at most K) atoms within max_radius with topk and
# softmaxes over them; restated as a dense masked softmax over all atoms (identical weights: every gathered atom beyond the
# radius is masked to -inf there, and atoms not gathered are beyond the radius by construction of K).
def local_self_attention(sd: StateDict, prefix: str, cfg: OracleConfig, src: Tensor, positions: Tensor, masked_elements: Tensor) -> Tensor:
    B, V, D = src.shape
    H = cfg.num_heads
    qkv = torch.nn.functional.linear(src, sd[f"{prefix}.self_attn.qkv_proj.weight"]).reshape(B, V, H, 3 * D)
    q, k, v = torch.split(qkv, [D, D, D], dim=-1)  # [B,V,H,D] each (key_query_dim = value_dim = d_model, custom_attention_encoder.py:146-153)
    diff = positions[:, :, None, :] - positions[:, None, :, :]
    dist = torch.sqrt((diff * diff).sum(-1))  # cdist(..., "donot_use_mm_for_euclid_dist") :62-64
    pad = masked_elements[:, None, :] | masked_elements[:, :, None]
    dist = dist.masked_fill(pad, math.inf)  # :67-70
    outside = dist > cfg.max_radius  # :87-89 (neighbor_mask)
    scores = torch.einsum("bihd,bjhd->bijh", q, k) / math.sqrt(D)  # :99-100
    scores = scores.masked_fill(outside[..., None], -math.inf)
    w = torch.softmax(scores, dim=-2)
    w = torch.nan_to_num(w, nan=0.0).masked_fill(outside[..., None], 0.0)  # :104-109 (rows with no neighbour: zeros)
    out = torch.einsum("bijh,bjhd->bihd", w, v).reshape(B, V, H * D)
    return torch.nn.functional.linear(out, sd[f"{prefix}.self_attn.output_proj.weight"])


def encoder_layer(sd: StateDict, prefix: str, cfg: OracleConfig, src: Tensor, scores) -> Tensor:
    B, V, D = src.shape
    if isinstance(scores, tuple):  # ("local", positions, masked_elements)
        src2 = local_self_attention(sd, prefix, cfg, src, scores[1], scores[2])
        src = layer_norm(src + src2, sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], cfg.layer_norm_eps)
        h = torch.relu(torch.nn.functional.linear(src, sd[f"{prefix}.linear1.weight"], sd[f"{prefix}.linear1.bias"]))
        src2 = torch.nn.functional.linear(h, sd[f"{prefix}.linear2.weight"], sd[f"{prefix}.linear2.bias"])
        return layer_norm(src + src2, sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], cfg.layer_norm_eps)
    if callable(scores):  # chebyshev_kernel: every attention layer has its own basis function -> its own scores
        scores = scores(prefix)
    H = scores.shape[1]
    values = torch.nn.functional.linear(src, sd[f"{prefix}.self_attn.values_proj.weight"])  # [B,V,H*Dv]
    Dv = values.shape[-1] // H
    values = values.reshape(B, V, H, Dv).transpose(1, 2)  # [B,H,V,Dv]
    attended = scores @ values  # attend(), kernel_attention.py:139
    flat = attended.transpose(-2, -3).reshape(B, V, H * Dv)  # flatten_multihead :142-156
    src2 = torch.nn.functional.linear(flat, sd[f"{prefix}.self_attn.attention._out_projection.weight"])
    src = layer_norm(src + src2, sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], cfg.layer_norm_eps)
    h = torch.relu(torch.nn.functional.linear(src, sd[f"{prefix}.linear1.weight"], sd[f"{prefix}.linear1.bias"]))
    src2 = torch.nn.functional.linear(h, sd[f"{prefix}.linear2.weight"], sd[f"{prefix}.linear2.bias"])
    src = layer_norm(src + src2, sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], cfg.layer_norm_eps)
    return src


# modules/layers/custom_transformer_block.py:46-82
def transformer_block(sd: StateDict, prefix: str, cfg: OracleConfig, inp: Tensor, scores: Tensor) -> Tensor:
    nh = len(cfg.latent_mlp_hidden_dims)
    h = mlp(sd, f"{prefix}.in_mlp", nh, inp)
    for t in range(cfg.num_transformer_layers):
        h = encoder_layer(sd, f"{prefix}.encoder_layers.{t}", cfg, h, scores)
    return mlp(sd, f"{prefix}.out_mlp", nh, h)


def transforms_positions(cfg: OracleConfig, layer_idx: int) -> bool:
    # model_constructor.py:169
    return layer_idx % 2 == cfg.position_layer_index_mod_2


# modules/custom_transformer_nvp.py:44-93
def scale_and_shift(
    sd: StateDict, cfg: OracleConfig, k: int, z_coords, z_velocs, x_features, x_coords, x_velocs, scores
) -> Tuple[Tensor, Tensor]:
    z_other = z_velocs if transforms_positions(cfg, k) else z_coords
    u = torch.cat((x_features, x_coords, x_velocs, z_other), dim=-1)
    s = transformer_block(sd, f"flow.chain.{k}.scale_transformer", cfg, u, scores)
    scale = torch.exp(s)
    shift = transformer_block(sd, f"flow.chain.{k}.shift_transformer", cfg, u, scores)
    return scale, shift


# modules/layers/nvp.py:22-183 + modules/model_wrappers/flow.py:51-103
def sequential_flow(
    sd: StateDict,
    cfg: OracleConfig,
    z_coords: Tensor,
    z_velocs: Tensor,
    x_features: Tensor,
    x_coords: Tensor,
    x_velocs: Tensor,
    masked_elements: Tensor,
    delta_logp: Tensor,
    reverse: bool,
    distance_mode: str = "cdist",
    trace: Optional[list] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    # The reference evaluates the scores once per pass through its Cache
    # (model_constructor.py:189-196; kernel_attention.py:197-206): the cache key maps `lengthscales` to 0, so the scores
    # of the FIRST attention layer executed in the pass are reused by every layer.  For `kernel` all layers hold the same
    # buffer; for `learnable_kernel` (lengthscales = exp(log_lengthscales), a Parameter per layer) that is the scale
    # network's first encoder layer of coupling layer 0 in the density direction and of the LAST coupling layer when sampling.
    first = cfg.num_coupling_layers - 1 if reverse else 0
    att = f"flow.chain.{first}.scale_transformer.encoder_layers.0.self_attn.attention"
    if cfg.attention_type == "local":
        ls = None
    elif cfg.attention_type == "learnable_kernel":
        ls = torch.exp(sd[f"{att}.log_lengthscales"]).to(x_coords.dtype)
    else:
        ls = sd[f"{att}.lengthscales"].to(x_coords.dtype)
    if cfg.attention_type == "local":
        scores = ("local", x_coords, masked_elements)
    elif cfg.attention_type == "chebyshev_kernel":
        # the basis function (a lambda per module, kernel_attention.py:333-335) is part of the cache key: no sharing
        def scores(prefix):
            return kernel_attention_scores(x_coords, masked_elements, ls, distance_mode,
                                           cheb_coeffs=sd[f"{prefix}.self_attn.attention.cheb_coeffs"].to(x_coords.dtype),
                                           force_asymptotic_zero=cfg.force_asymptotic_zero)
    else:
        scores = kernel_attention_scores(x_coords, masked_elements, ls, distance_mode)
    idxs = range(cfg.num_coupling_layers)
    idxs = idxs[::-1] if reverse else idxs
    keep = ~masked_elements[:, :, None]
    for k in idxs:
        scale, shift = scale_and_shift(sd, cfg, k, z_coords, z_velocs, x_features, x_coords, x_velocs, scores)
        log_scales = torch.log(scale) * keep  # nvp.py:127 / :175 (log(exp(s)) literally)
        if reverse:
            logdet = -torch.sum(log_scales, dim=(-1, -2))  # :176
            if transforms_positions(cfg, k):
                z_coords = (z_coords - shift) / scale  # :179
            else:
                z_velocs = (z_velocs - shift) / scale  # :181
        else:
            logdet = torch.sum(log_scales, dim=(-1, -2))  # :128
            if transforms_positions(cfg, k):
                z_coords = z_coords * scale + shift  # :131
            else:
                z_velocs = z_velocs * scale + shift  # :133
        delta_logp = delta_logp - logdet  # :86
        if trace is not None:
            trace.append(dict(layer=k, scale=scale, shift=shift, z_coords=z_coords, z_velocs=z_velocs))
    return z_coords, z_velocs, delta_logp


def _normal_log_prob(value: Tensor, log_scale: Tensor) -> Tensor:
    # torch.distributions.Normal.log_prob with loc=0, scale=exp(log_scale)  (flow.py:159-166,191-192)
    scale = torch.exp(log_scale)
    var = scale**2
    return -(value**2) / (2 * var) - torch.log(scale) - math.log(math.sqrt(2 * math.pi))


def prior_log_prob(sd: StateDict, z_coords, z_velocs, masked_elements) -> Tensor:
    # flow.py:191-203 / :322-333
    keep = ~masked_elements[:, :, None]
    lp_c = (keep * _normal_log_prob(z_coords, sd["coords_prior_log_scale"].to(z_coords.dtype))).sum(dim=(-1, -2))
    lp_v = (keep * _normal_log_prob(z_velocs, sd["velocs_prior_log_scale"].to(z_coords.dtype))).sum(dim=(-1, -2))
    return lp_c + lp_v


# modules/model_wrappers/flow.py:131-215
def log_likelihood(
    sd: StateDict,
    cfg: OracleConfig,
    atom_types: Tensor,
    x_coords: Tensor,
    x_velocs: Tensor,
    y_coords: Tensor,
    y_velocs: Tensor,
    masked_elements: Tensor,
    distance_mode: str = "cdist",
    return_latent: bool = False,
    trace: Optional[list] = None,
):
    y_res = y_coords - x_coords  # :148-149 (un-centred x)
    com = centre_of_mass(x_coords, masked_elements)  # :156
    xc = x_coords - com  # :157
    delta = torch.zeros(x_coords.shape[0], dtype=x_coords.dtype)
    feats = torch.nn.functional.embedding(atom_types, sd["flow.atom_embedder.weight"])  # :172
    z_c, z_v, delta = sequential_flow(
        sd, cfg, y_res, y_velocs, feats, xc, x_velocs, masked_elements, delta, False, distance_mode, trace
    )
    lp = prior_log_prob(sd, z_c, z_v, masked_elements)
    ll = lp - delta  # :203
    if return_latent:
        return ll, z_c, z_v
    return ll


# modules/model_wrappers/density_model_base.py:14-47
def nll_loss(sd, cfg, atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements, distance_mode="cdist"):
    num_atoms = (~masked_elements).sum(dim=1)
    ll = log_likelihood(sd, cfg, atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements, distance_mode)
    return -(ll / num_atoms).mean()


def draw_latents(sd: StateDict, x_coords: Tensor, x_velocs: Tensor, num_samples: int) -> Tuple[Tensor, Tensor]:
    """The reference's RNG consumption (flow.py:264-277): Normal(0, exp(log_scale)).rsample((S,))
    for coords and then velocs == two `normal_()` draws of shape [S,B,V,3] from the default
    generator, each scaled by exp(log_scale)."""
    zc = torch.distributions.Normal(
        loc=torch.zeros_like(x_coords), scale=torch.exp(sd["coords_prior_log_scale"])
    ).rsample((num_samples,))
    zv = torch.distributions.Normal(
        loc=torch.zeros_like(x_velocs), scale=torch.exp(sd["velocs_prior_log_scale"])
    ).rsample((num_samples,))
    return zc, zv


# modules/model_wrappers/flow.py:242-336
def conditional_sample_with_logp(
    sd: StateDict,
    cfg: OracleConfig,
    atom_types: Tensor,
    x_coords: Tensor,
    x_velocs: Tensor,
    masked_elements: Tensor,
    num_samples: int,
    z_coords: Optional[Tensor] = None,  # [S,B,V,3] latent draws (already scaled); drawn if None
    z_velocs: Optional[Tensor] = None,
    distance_mode: str = "cdist",
):
    B = x_coords.shape[0]
    S = num_samples
    com = centre_of_mass(x_coords, masked_elements)
    xc = x_coords - com
    if z_coords is None:
        z_coords, z_velocs = draw_latents(sd, xc, x_velocs, S)
    zc = z_coords.reshape(-1, *z_coords.shape[-2:])
    zv = z_velocs.reshape(-1, *z_velocs.shape[-2:])
    feats = torch.nn.functional.embedding(atom_types, sd["flow.atom_embedder.weight"])
    delta = torch.zeros(B * S, dtype=x_coords.dtype)
    mask_rep = masked_elements.repeat(S, 1)
    yc_res, yv, delta = sequential_flow(
        sd,
        cfg,
        zc,
        zv,
        feats.repeat(S, 1, 1),
        xc.repeat(S, 1, 1),
        x_velocs.repeat(S, 1, 1),
        mask_rep,
        delta,
        True,
        distance_mode,
    )
    x_un = (xc + com).repeat(S, 1, 1)  # :303-304
    yc = x_un + yc_res  # :308
    # NB the reference broadcasts the *un-repeated* mask here (flow.py:326-331) which only
    # works for S==1 or B==1; with the repeated mask the value is identical in those cases.
    lp = prior_log_prob(sd, zc, zv, mask_rep)
    logp = lp + delta  # :334
    V = x_coords.shape[1]
    return yc.reshape(S, B, V, 3), yv.reshape(S, B, V, 3), logp.reshape(S, B)


# --------------------------------------------------------------------------------------
# Helpers shared by tests / bench (not reference restatements).
def state_dict_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference model's state_dict for `cfg`."""
    E, D, F = cfg.atom_embedding_dim, cfg.d_model, cfg.dim_feedforward
    H = cfg.num_heads if cfg.attention_type == "local" else len(cfg.lengthscales)
    hid = list(cfg.latent_mlp_hidden_dims)
    out: Dict[str, Tuple[int, ...]] = {
        "coords_prior_log_scale": (),
        "velocs_prior_log_scale": (),
        "flow.atom_embedder.weight": (5, E),
    }

    def add_mlp(prefix, din, dout):
        dims = [din] + hid + [dout]
        for i in range(len(dims) - 1):
            out[f"{prefix}._layers.{2 * i}.weight"] = (dims[i + 1], dims[i])
            out[f"{prefix}._layers.{2 * i}.bias"] = (dims[i + 1],)

    for k in range(cfg.num_coupling_layers):
        for net in ("scale", "shift"):
            p = f"flow.chain.{k}.{net}_transformer"
            add_mlp(f"{p}.in_mlp", E + 9, D)
            for t in range(cfg.num_transformer_layers):
                q = f"{p}.encoder_layers.{t}"
                if cfg.attention_type == "local":
                    out[f"{q}.self_attn.qkv_proj.weight"] = (H * 3 * D, D)
                    out[f"{q}.self_attn.output_proj.weight"] = (D, H * D)
                else:
                    out[f"{q}.self_attn.values_proj.weight"] = (H * D, D)
                    out[f"{q}.self_attn.attention.lengthscales"] = (H,)
                if cfg.attention_type == "learnable_kernel":
                    out[f"{q}.self_attn.attention.log_lengthscales"] = (H,)
                if cfg.attention_type == "chebyshev_kernel":
                    out[f"{q}.self_attn.attention.cheb_coeffs"] = (H, cfg.cheb_order)
                if cfg.attention_type != "local":
                    out[f"{q}.self_attn.attention._out_projection.weight"] = (D, H * D)
                out[f"{q}.linear1.weight"] = (F, D)
                out[f"{q}.linear1.bias"] = (F,)
                out[f"{q}.linear2.weight"] = (D, F)
                out[f"{q}.linear2.bias"] = (D,)
                for n in ("norm1", "norm2"):
                    out[f"{q}.{n}.weight"] = (D,)
                    out[f"{q}.{n}.bias"] = (D,)
            add_mlp(f"{p}.out_mlp", D, 3)
    return out


def synth_state_dict(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> StateDict:
    """Deterministic weights that depend only on (key, shape, seed) -- independent of module
    construction order, so the reference model, this oracle and the CUDA module can all be
    loaded with bit-identical parameters without shipping 144 MB of weights.
    Distributions match torch defaults in scale: U(-1,1)/sqrt(fan_in) for Linear weights and
    biases; LayerNorm gamma 1+0.1 U, beta 0.1 U; embedding N(0,1); prior log-scales small."""
    import zlib

    sd: StateDict = {}
    for key, shape in state_dict_shapes(cfg).items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2**31))
        if key.endswith("cheb_coeffs"):  # the reference's initial value + a different perturbation in every layer and head
            base = torch.tensor((CHEB_COEFFS_EXPMX + [0.0] * max(0, shape[1] - len(CHEB_COEFFS_EXPMX)))[: shape[1]])
            t = base[None, :].expand(shape) + 0.02 * (torch.rand(shape, generator=g) * 2 - 1)
        elif key.endswith("log_lengthscales"):  # a different value in every layer (the reference uses only the first executed one)
            t = torch.log(torch.tensor(cfg.lengthscales, dtype=torch.float32)) + 0.3 * (torch.rand(shape, generator=g) * 2 - 1)
        elif key.endswith("lengthscales"):
            t = torch.tensor(cfg.lengthscales, dtype=torch.float32)
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
            wshape = state_dict_shapes_cache(cfg)[key[: -len("bias")] + "weight"]
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(wshape[1])
        sd[key] = t.to(dtype)
    return sd


_shape_cache: Dict[int, Dict[str, Tuple[int, ...]]] = {}


def state_dict_shapes_cache(cfg: OracleConfig):
    k = id(cfg)
    if k not in _shape_cache:
        _shape_cache[k] = state_dict_shapes(cfg)
    return _shape_cache[k]


def to_dtype(sd: StateDict, dtype) -> StateDict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def nll_loss_and_grads(sd: StateDict, cfg: OracleConfig, atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements,
                       distance_mode: str = "cdist"):
    """Loss of density_model_base.py:27-42 and its torch-autograd gradient w.r.t. every floating-point
    entry of the state dict that the reference registers as a Parameter (the `lengthscales` entries are
    buffers, kernel_attention.py:169-171) -- what `loss.backward()` leaves in `.grad` in train.py."""
    leaves = {k: v.detach().clone().requires_grad_(not k.endswith(".lengthscales")) for k, v in sd.items()}
    loss = nll_loss(leaves, cfg, atom_types, x_coords, x_velocs, y_coords, y_velocs, masked_elements, distance_mode)
    names = [k for k, v in leaves.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}

"""Pin the CPU oracle (oracle/flow_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo

TINY = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4,
                       num_transformer_layers=2, d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0])
FULL = fo.OracleConfig()
CASES = [("tiny_ad_ragged", TINY), ("tiny_ad", TINY), ("full_ad22", FULL), ("full_ad22_ragged", FULL), ("full_2olx65", FULL)]
# `learnable_kernel` attention: per-layer log_lengthscales, only the first executed layer's are used (cache-key quirk)
TINY_L = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                         d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="learnable_kernel")
FULL_L = fo.OracleConfig(attention_type="learnable_kernel")
LEARNABLE = [("tiny_ad_learnable", TINY_L), ("full_ad22_learnable", FULL_L)]
# `chebyshev_kernel` attention: a Chebyshev-rational basis function per attention layer, no score sharing between layers
TINY_C = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                         d_model=16, dim_feedforward=32, lengthscales=[0.3, 1.0], attention_type="chebyshev_kernel", cheb_order=6,
                         force_asymptotic_zero=True)
FULL_C = fo.OracleConfig(attention_type="chebyshev_kernel", cheb_order=12, force_asymptotic_zero=False)
CHEBYSHEV = [("tiny_ad_chebyshev", TINY_C), ("full_ad22_chebyshev", FULL_C)]
# `local` attention: dot-product attention over the atoms within max_radius (modules/layers/local_self_attention.py)
TINY_LOC = fo.OracleConfig(atom_embedding_dim=8, latent_mlp_hidden_dims=[24], num_coupling_layers=4, num_transformer_layers=2,
                           d_model=16, dim_feedforward=32, lengthscales=[], attention_type="local", max_radius=0.3, num_heads=3)
FULL_LOC = fo.OracleConfig(lengthscales=[], attention_type="local", max_radius=0.45, num_heads=6)
LOCAL = [("tiny_ad_local", TINY_LOC), ("full_ad22_local", FULL_LOC)]


def load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


@pytest.mark.parametrize("name,cfg", CASES + LEARNABLE + CHEBYSHEV + LOCAL)
def test_log_likelihood_and_loss(golden_dir, name, cfg):
    g = load(golden_dir, name)
    sd = fo.synth_state_dict(cfg, int(g["weight_seed"]))
    args = (g["atom_types"], g["x_coords"], g["x_velocs"], g["y_coords"], g["y_velocs"], g["masked_elements"])
    ll = fo.log_likelihood(sd, cfg, *args)
    # same torch ops in the same order as the reference: agreement is to fp32 round-off
    torch.testing.assert_close(ll, g["log_likelihood"], rtol=2e-6, atol=2e-5)
    loss = fo.nll_loss(sd, cfg, *args)
    torch.testing.assert_close(loss, g["loss"], rtol=2e-6, atol=2e-6)
    # fp64 oracle with direct distances == reference fp64 run (cdist in fp64)
    sd64 = fo.to_dtype(sd, torch.float64)
    ll64 = fo.log_likelihood(sd64, cfg, args[0], *[a.double() for a in args[1:5]], args[5], distance_mode="direct")
    torch.testing.assert_close(ll64, g["log_likelihood_f64"], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("name,cfg", CASES)
def test_scores_and_layer0(golden_dir, name, cfg):
    g = load(golden_dir, name)
    sd = fo.synth_state_dict(cfg, int(g["weight_seed"]))
    mask = g["masked_elements"]
    xc = g["x_coords"] - fo.centre_of_mass(g["x_coords"], mask)
    ls = torch.tensor(cfg.lengthscales)
    sc = fo.kernel_attention_scores(xc, mask, ls)
    torch.testing.assert_close(sc, g["scores"], rtol=1e-6, atol=1e-7)
    # rows over un-masked keys sum to ~1 (reference tests/test_kernel_attention.py:19-46)
    assert torch.allclose(sc.sum(-1), torch.ones_like(sc.sum(-1)), atol=1e-3)
    # direct-difference distances stay within 2e-3 of the mm-based cdist scores
    sc_d = fo.kernel_attention_scores(xc, mask, ls, distance_mode="direct")
    assert (sc_d - sc).abs().max() < 2e-3
    feats = torch.nn.functional.embedding(g["atom_types"], sd["flow.atom_embedder.weight"])
    scale, shift = fo.scale_and_shift(sd, cfg, 0, g["y_coords"] - g["x_coords"], g["y_velocs"], feats, xc, g["x_velocs"], sc)
    torch.testing.assert_close(scale, g["layer0_scale"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(shift, g["layer0_shift"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,cfg", CASES + LEARNABLE + CHEBYSHEV + LOCAL)
def test_sampling(golden_dir, name, cfg):
    g = load(golden_dir, name)
    sd = fo.synth_state_dict(cfg, int(g["weight_seed"]))
    at, x, xv, mask = g["atom_types"], g["x_coords"], g["x_velocs"], g["masked_elements"]
    yc, yv, lp = fo.conditional_sample_with_logp(sd, cfg, at, x, xv, mask, 1, g["s1_z_coords"], g["s1_z_velocs"])
    torch.testing.assert_close(yc, g["s1_y_coords"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(yv, g["s1_y_velocs"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(lp, g["s1_logp"], rtol=2e-6, atol=5e-5)
    S = g["sS_z_coords"].shape[0]
    yc, yv, lp = fo.conditional_sample_with_logp(sd, cfg, at[:1], x[:1], xv[:1], mask[:1], S, g["sS_z_coords"], g["sS_z_velocs"])
    torch.testing.assert_close(yc, g["sS_y_coords"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(lp, g["sS_logp"], rtol=2e-6, atol=5e-5)
    p_yx = fo.log_likelihood(sd, cfg, at[:1].repeat(S, 1), yc.squeeze(1), yv.squeeze(1), x[:1].repeat(S, 1, 1),
                             xv[:1].repeat(S, 1, 1), mask[:1].repeat(S, 1))
    torch.testing.assert_close(p_yx, g["sS_p_yx"], rtol=5e-6, atol=2e-4)


def test_rng_contract(golden_dir):
    """Latent draws == two normal_() draws [S,B,V,3], coords first (flow.py:274-275)."""
    g = load(golden_dir, "tiny_ad")
    sd = fo.synth_state_dict(TINY, 0)
    torch.manual_seed(4321 + 12)
    zc, zv = fo.draw_latents(sd, g["x_coords"][:1], g["x_velocs"][:1], 4)
    torch.testing.assert_close(zc, g["sS_z_coords"], rtol=0, atol=0)
    torch.testing.assert_close(zv, g["sS_z_velocs"], rtol=0, atol=0)
    torch.manual_seed(4321 + 12)
    e1 = torch.empty(4, 1, 22, 3).normal_() * torch.exp(sd["coords_prior_log_scale"])
    torch.testing.assert_close(e1, zc, rtol=0, atol=0)


def test_sample_then_density_roundtrip(golden_dir):
    """Size-independent property: log p from sampling == log_likelihood of the sample."""
    g = load(golden_dir, "full_ad22")
    sd = fo.synth_state_dict(FULL, 0)
    at, x, xv, mask = g["atom_types"], g["x_coords"], g["x_velocs"], g["masked_elements"]
    yc, yv, lp = fo.conditional_sample_with_logp(sd, FULL, at, x, xv, mask, 1, g["s1_z_coords"], g["s1_z_velocs"])
    ll = fo.log_likelihood(sd, FULL, at, x, xv, yc[0], yv[0], mask)
    torch.testing.assert_close(ll, lp[0], rtol=1e-5, atol=2e-3)


@pytest.mark.parametrize("name,cfg", LEARNABLE)
def test_learnable_scores_use_first_executed_layer(golden_dir, name, cfg):
    g = load(golden_dir, name)
    sd = fo.synth_state_dict(cfg, int(g["weight_seed"]))
    mask = g["masked_elements"]
    xc = g["x_coords"] - fo.centre_of_mass(g["x_coords"], mask)
    ls0 = torch.exp(sd["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"])
    torch.testing.assert_close(fo.kernel_attention_scores(xc, mask, ls0), g["scores"], rtol=1e-6, atol=1e-7)
    # the layers really differ (otherwise the quirk would be untested)
    ls_last = torch.exp(sd[f"flow.chain.{cfg.num_coupling_layers - 1}.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"])
    assert (ls0 - ls_last).abs().max() > 1e-2


@pytest.mark.parametrize("name,cfg", CHEBYSHEV)
def test_chebyshev_scores(golden_dir, name, cfg):
    """Scores of the first attention layer vs the reference's chebyshev_basis_function; the known answer of the reference's
    tests/test_kernel_attention.py:163-208 -- the initial coefficients reproduce exp(-s) -- holds for the restated basis."""
    g = load(golden_dir, name)
    sd = fo.synth_state_dict(cfg, int(g["weight_seed"]))
    mask = g["masked_elements"]
    xc = g["x_coords"] - fo.centre_of_mass(g["x_coords"], mask)
    cc = sd["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.cheb_coeffs"]
    sc = fo.kernel_attention_scores(xc, mask, torch.tensor(cfg.lengthscales), cheb_coeffs=cc, force_asymptotic_zero=cfg.force_asymptotic_zero)
    torch.testing.assert_close(sc, g["scores"], rtol=1e-5, atol=1e-6)
    s_ = torch.linspace(0.0, 3.0, 50)[None, None, None, :]
    full = torch.tensor(fo.CHEB_COEFFS_EXPMX)[None, :]
    approx = fo.chebyshev_basis(s_, full, False)
    assert (approx - torch.exp(-(s_**2))).abs().max() < 1e-5


def test_chebyshev_known_answers():
    """The reference's own known answers for the Chebyshev-rational basis (tests/test_kernel_attention.py:163-208): five
    expansion values from the Julia implementation, exp(-s^2) reproduced to 1e-2 on [0, 10] by the first six coefficients,
    and the asymptotic-zero option."""
    eye = torch.eye(5)
    s = torch.full((1, 1, 1, 1), 0.7).sqrt()  # the basis squares its argument; the reference expands 0.7 directly
    vals = torch.stack([fo.chebyshev_basis(s, eye[c][None], False).flatten()[0] for c in range(5)])
    ref = torch.tensor([1.0, -0.17647058823529416, -0.9377162629757785, 0.507429269285569, 0.7586235796985188])
    assert torch.allclose(vals, ref)
    coeffs = torch.tensor([0.42758357, -0.54642403, 0.07106222, 0.05473271, 0.00574419, -0.00792641])[None].expand(3, -1)
    torch.manual_seed(0)
    d = 10.0 * torch.rand((2, 3, 64, 128))
    assert torch.allclose(torch.exp(-(d**2.0)), fo.chebyshev_basis(d, coeffs, False), atol=1.0e-2, rtol=0.0)
    far = fo.chebyshev_basis(1000.0 * torch.ones((1, 3, 1, 1)), coeffs, True).flatten()
    assert torch.allclose(torch.zeros_like(far), far, atol=1.0e-6, rtol=0.0)

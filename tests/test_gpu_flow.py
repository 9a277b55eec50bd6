"""Parity of the CUDA flow (through the C ABI, via the reference-shaped module) against
(a) golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded inputs.
Tolerance: BASELINE.json north_star -- log-prob and samples within 1e-4 relative (fp32)."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from tests.common import (EMPTY_ADJ, EMPTY_EBI, FULL_C, FULL_L, FULL_LOC, FULL_O, TINY_C, TINY_L, TINY_LOC, TINY_O, build_model,
                          load_golden)

pytestmark = pytest.mark.gpu
REL = 1e-4  # north_star tolerance

CASES = [("tiny_ad_ragged", TINY_O, "fp32"), ("tiny_ad", TINY_O, "fp32"), ("full_ad22", FULL_O, "fp32"),
         ("full_ad22_ragged", FULL_O, "fp32"), ("full_2olx65", FULL_O, "fp32")]
# learnable_kernel attention: the golden files carry no layer-0 trace; the lengthscales differ per layer and direction
LEARNABLE = [("tiny_ad_learnable", TINY_L, "fp32"), ("full_ad22_learnable", FULL_L, "fp32"), ("full_ad22_learnable", FULL_L, "bf16x3"),
             ("tiny_ad_chebyshev", TINY_C, "fp32"), ("full_ad22_chebyshev", FULL_C, "fp32"), ("full_ad22_chebyshev", FULL_C, "bf16x3"),
             # local (dot-product) attention: CUDA-core attention kernels; with bf16x3 the MLPs and the FFN run on the tensor cores
             ("tiny_ad_local", TINY_LOC, "fp32"), ("full_ad22_local", FULL_LOC, "fp32"), ("full_ad22_local", FULL_LOC, "bf16x3")]


def _kw(g, dev="cuda", rows=slice(None)):
    return dict(atom_types=g["atom_types"][rows].to(dev), x_coords=g["x_coords"][rows].to(dev), x_velocs=g["x_velocs"][rows].to(dev),
                adj_list=EMPTY_ADJ.to(dev), edge_batch_idx=EMPTY_EBI.to(dev), masked_elements=g["masked_elements"][rows].to(dev))


def assert_rel(actual, expected, rel=REL, what=""):
    actual, expected = actual.detach().cpu().double(), expected.detach().cpu().double()
    scale = expected.abs().max().clamp_min(1e-30)
    err = (actual - expected).abs().max() / scale
    assert torch.isfinite(actual).all(), what
    assert err < rel, f"{what}: max rel err {err:.3e} >= {rel}"
    return float(err)


@pytest.mark.parametrize("name,cfg,prec", CASES + LEARNABLE)
@torch.no_grad()
def test_golden_log_likelihood_and_loss(name, cfg, prec):
    g = load_golden(name)
    m, _ = build_model(cfg, prec, int(g["weight_seed"]))
    ll = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert ll.shape == g["log_likelihood"].shape and ll.dtype == torch.float32 and ll.is_cuda
    assert_rel(ll, g["log_likelihood"], what="log_likelihood vs reference fp32")
    assert_rel(ll, g["log_likelihood_f64"], what="log_likelihood vs reference fp64")
    loss = m(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert loss.dim() == 0
    assert_rel(loss, g["loss"], what="loss")


@pytest.mark.parametrize("name,cfg,prec", CASES)
def test_golden_scores_and_layer0(name, cfg, prec):
    g = load_golden(name)
    m, sd = build_model(cfg, prec, int(g["weight_seed"]))
    mask = g["masked_elements"]
    xc = g["x_coords"] - fo.centre_of_mass(g["x_coords"], mask)
    sc = m.attention_scores(xc.cuda(), mask.cuda())
    # the reference uses the mm-based cdist (diagonal distances up to ~3e-4 nm instead of 0); the kernel uses direct differences
    assert (sc.cpu() - g["scores"]).abs().max() < 2e-3
    ref_direct = fo.kernel_attention_scores(xc, mask, torch.tensor(cfg.lengthscales), distance_mode="direct")
    torch.testing.assert_close(sc.cpu(), ref_direct, rtol=2e-5, atol=1e-7)
    # rows sum to 1 (reference tests/test_kernel_attention.py:19-46, atol 1e-3)
    assert torch.allclose(sc.sum(-1), torch.ones_like(sc.sum(-1)), atol=1e-3)
    scale, shift = m.scale_and_shift(0, g["atom_types"].cuda(), (g["y_coords"] - g["x_coords"]).cuda(), g["y_velocs"].cuda(), xc.cuda(),
                                     g["x_velocs"].cuda(), mask.cuda())
    keep = (~mask)[:, :, None].expand_as(g["layer0_scale"])
    assert_rel(scale.cpu()[keep], g["layer0_scale"][keep], what="layer0 scale")
    assert_rel(shift.cpu()[keep], g["layer0_shift"][keep], what="layer0 shift")


@pytest.mark.parametrize("name,cfg,prec", CASES + LEARNABLE)
@torch.no_grad()
def test_golden_sampling(name, cfg, prec):
    g = load_golden(name)
    m, _ = build_model(cfg, prec, int(g["weight_seed"]))
    mask = g["masked_elements"]
    keep1 = (~mask)[None, :, :, None].expand_as(g["s1_y_coords"])
    yc, yv, lp = m.sample_from_latents(g["atom_types"].cuda(), g["x_coords"].cuda(), g["x_velocs"].cuda(), mask.cuda(),
                                       g["s1_z_coords"].cuda(), g["s1_z_velocs"].cuda())
    assert yc.shape == g["s1_y_coords"].shape and lp.shape == g["s1_logp"].shape
    assert_rel(yc.cpu()[keep1], g["s1_y_coords"][keep1], what="S=1 y_coords")
    assert_rel(yv.cpu()[keep1], g["s1_y_velocs"][keep1], what="S=1 y_velocs")
    assert_rel(lp, g["s1_logp"], what="S=1 logp")
    # S proposals from one state + reverse-move density (the MH iteration of evaluation_utils.py:609-657)
    S = g["sS_z_coords"].shape[0]
    yc, yv, lp = m.sample_from_latents(g["atom_types"][:1].cuda(), g["x_coords"][:1].cuda(), g["x_velocs"][:1].cuda(), mask[:1].cuda(),
                                       g["sS_z_coords"].cuda(), g["sS_z_velocs"].cuda())
    keepS = (~mask[:1])[None, :, :, None].expand_as(g["sS_y_coords"])
    assert_rel(yc.cpu()[keepS], g["sS_y_coords"][keepS], what="S y_coords")
    assert_rel(lp, g["sS_logp"], what="S logp")
    p_yx = m.log_likelihood(atom_types=g["atom_types"][:1].repeat(S, 1).cuda(), y_coords=g["x_coords"][:1].repeat(S, 1, 1).cuda(),
                            y_velocs=g["x_velocs"][:1].repeat(S, 1, 1).cuda(), x_coords=yc.squeeze(1), x_velocs=yv.squeeze(1),
                            adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask[:1].repeat(S, 1).cuda())
    assert_rel(p_yx, g["sS_p_yx"], what="p_yx")


def test_oracle_parity_seeded_random_batch():
    """Same seeded inputs through the CUDA path and the CPU oracle, ragged batch with padding."""
    torch.manual_seed(3)
    B, V = 6, 19
    lengths = [19, 19, 11, 7, 3, 1]
    m, sd = build_model(TINY_O, "fp32", weight_seed=5)
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    at = torch.randint(0, 5, (B, V)) * (~mask)
    x, xv, y, yv = (torch.randn(B, V, 3) * s * keep for s in (0.3, 1.0, 0.3, 1.0))
    ll_ref = fo.log_likelihood(sd, TINY_O, at, x, xv, y, yv, mask, distance_mode="direct")
    ll = m.log_likelihood(atom_types=at.cuda(), x_coords=x.cuda(), x_velocs=xv.cuda(), y_coords=y.cuda(), y_velocs=yv.cuda(),
                          adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask.cuda())
    assert_rel(ll, ll_ref, what="ragged ll vs oracle")


def test_batch_equals_loop():
    """Reference tests/test_batching.py:132-177: batched log_likelihood == per-sample loop (rtol/atol 1e-4)."""
    g = load_golden("tiny_ad_ragged")
    m, _ = build_model(TINY_O, "fp32", 0)
    ll = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    lengths = (~g["masked_elements"]).sum(1).tolist()
    for b, n in enumerate(lengths):
        rows = slice(b, b + 1)
        one = m.log_likelihood(atom_types=g["atom_types"][rows, :n].cuda(), x_coords=g["x_coords"][rows, :n].cuda(),
                               x_velocs=g["x_velocs"][rows, :n].cuda(), y_coords=g["y_coords"][rows, :n].cuda(),
                               y_velocs=g["y_velocs"][rows, :n].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(),
                               masked_elements=g["masked_elements"][rows, :n].cuda())
        torch.testing.assert_close(one[0], ll[b], rtol=1e-4, atol=1e-4)


def test_rng_contract_and_roundtrip():
    """conditional_sample_with_logp consumes the device generator exactly like the reference
    (two normal_() draws [S,B,V,3], coords first) and sample -> log_likelihood round-trips."""
    g = load_golden("full_ad22")
    m, _ = build_model(FULL_O, "fp32", 0)
    kw = _kw(g)
    torch.manual_seed(99)
    yc, yv, lp = m.conditional_sample_with_logp(num_samples=1, **kw)
    torch.manual_seed(99)
    B, V = g["x_coords"].shape[:2]
    zc = torch.empty(1, B, V, 3, device="cuda").normal_() * torch.exp(m.coords_prior_log_scale.detach())
    zv = torch.empty(1, B, V, 3, device="cuda").normal_() * torch.exp(m.velocs_prior_log_scale.detach())
    yc2, yv2, lp2 = m.sample_from_latents(kw["atom_types"], kw["x_coords"], kw["x_velocs"], kw["masked_elements"], zc, zv)
    assert torch.equal(yc, yc2) and torch.equal(yv, yv2) and torch.equal(lp, lp2)
    ll = m.log_likelihood(y_coords=yc[0], y_velocs=yv[0], **kw)
    assert_rel(ll, lp[0], rel=1e-4, what="sample->density round trip")
    # S>1 with B>1 is rejected like the reference (flow.py:326-331 broadcast)
    with pytest.raises(ValueError):
        m.conditional_sample_with_logp(num_samples=3, **kw)
    y2, v2 = m.conditional_sample(num_samples=3, **kw)  # without logp any (S,B) works
    assert y2.shape == (3, B, V, 3)


def test_conditioning_broadcast_matches_repeat():
    """S proposals from B=1 == the same call with the conditioning explicitly repeated (flow.py:285-296)."""
    g = load_golden("tiny_ad")
    m, _ = build_model(TINY_O, "fp32", 0)
    S = g["sS_z_coords"].shape[0]
    a = m.sample_from_latents(g["atom_types"][:1].cuda(), g["x_coords"][:1].cuda(), g["x_velocs"][:1].cuda(),
                              g["masked_elements"][:1].cuda(), g["sS_z_coords"].cuda(), g["sS_z_velocs"].cuda())
    b = m.sample_from_latents(g["atom_types"][:1].repeat(S, 1).cuda(), g["x_coords"][:1].repeat(S, 1, 1).cuda(),
                              g["x_velocs"][:1].repeat(S, 1, 1).cuda(), g["masked_elements"][:1].repeat(S, 1).cuda(),
                              g["sS_z_coords"].transpose(0, 1).contiguous().cuda(), g["sS_z_velocs"].transpose(0, 1).contiguous().cuda())
    assert torch.equal(a[0].squeeze(1), b[0].squeeze(0)) and torch.equal(a[2].squeeze(1), b[2].squeeze(0))


def test_empty_and_errors():
    m, _ = build_model(TINY_O, "fp32", 0)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")  # noqa: E731
    ll = m.log_likelihood(atom_types=z(0, 5, dt=torch.long), x_coords=z(0, 5, 3), x_velocs=z(0, 5, 3), y_coords=z(0, 5, 3),
                          y_velocs=z(0, 5, 3), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=z(0, 5, dt=torch.bool))
    assert ll.shape == (0,)
    with pytest.raises(ValueError):
        m.log_likelihood(atom_types=z(2, 5, dt=torch.long), x_coords=z(2, 5, 3), x_velocs=z(2, 4, 3), y_coords=z(2, 5, 3),
                         y_velocs=z(2, 5, 3), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=z(2, 5, dt=torch.bool))
    with pytest.raises(TypeError):
        m.log_likelihood(atom_types=z(2, 5, dt=torch.long), x_coords=z(2, 5, 3).double(), x_velocs=z(2, 5, 3), y_coords=z(2, 5, 3),
                         y_velocs=z(2, 5, 3), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=z(2, 5, dt=torch.bool))


def test_state_dict_roundtrip_changes_output():
    """Loading a checkpoint (in-place copy into the same storage) is picked up by the next call."""
    g = load_golden("tiny_ad")
    m, _ = build_model(TINY_O, "fp32", 0)
    kw = dict(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    a = m.log_likelihood(**kw)
    m.load_state_dict(fo.synth_state_dict(TINY_O, 1))
    b = m.log_likelihood(**kw)
    assert not torch.allclose(a, b)
    m.load_state_dict(fo.synth_state_dict(TINY_O, 0))
    assert torch.equal(m.log_likelihood(**kw), a)


def test_learnable_kernel_module_surface():
    """State-dict keys of the reference's LearnableLengthscaleKernelAttention, scores from layer 0's exp(log_lengthscales),
    and `.train()` on a configuration without backward kernels raising instead of falling back."""
    g = load_golden("tiny_ad_learnable")
    m, sd = build_model(TINY_L, "fp32", int(g["weight_seed"]))
    assert set(m.state_dict().keys()) == set(sd.keys())
    assert "flow.chain.1.shift_transformer.encoder_layers.1.self_attn.attention.log_lengthscales" in dict(m.named_parameters())
    mask = g["masked_elements"]
    xc = g["x_coords"] - fo.centre_of_mass(g["x_coords"], mask)
    assert (m.attention_scores(xc.cuda(), mask.cuda()).cpu() - g["scores"]).abs().max() < 2e-3
    m.train()  # (tiny layer sizes / fp32: no backward kernels -- the full-size learnable training path is in test_gpu_train.py)
    with pytest.raises(Exception):
        m(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))


def test_local_attention_module_surface():
    """State-dict keys of the reference's LocalSelfAttention; oracle parity on a seeded ragged batch whose radius leaves some
    atoms without any neighbour but themselves; shared conditioning (S proposals from one state); no training path."""
    g = load_golden("tiny_ad_local")
    m, sd = build_model(TINY_LOC, "fp32", int(g["weight_seed"]))
    assert set(m.state_dict().keys()) == set(sd.keys())
    assert m.state_dict()["flow.chain.0.scale_transformer.encoder_layers.0.self_attn.qkv_proj.weight"].shape == (3 * 3 * 16, 16)
    with pytest.raises(TypeError):
        m.attention_scores(g["x_coords"].cuda(), g["masked_elements"].cuda())
    gen = torch.Generator().manual_seed(5)
    B, V = 5, 13
    lengths = torch.tensor([13, 7, 1, 10, 4])
    mask = torch.arange(V)[None, :] >= lengths[:, None]
    at = torch.randint(0, 5, (B, V), generator=gen)
    xc = torch.randn(B, V, 3, generator=gen) * 0.25  # many pairs beyond max_radius = 0.3
    xv, yv = torch.randn(B, V, 3, generator=gen), torch.randn(B, V, 3, generator=gen)
    yc = xc + 0.02 * torch.randn(B, V, 3, generator=gen)
    want = fo.log_likelihood(sd, TINY_LOC, at, xc, xv, yc, yv, mask)
    with torch.no_grad():
        got = m.log_likelihood(atom_types=at.cuda(), x_coords=xc.cuda(), x_velocs=xv.cuda(), y_coords=yc.cuda(), y_velocs=yv.cuda(),
                               adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask.cuda())
    assert_rel(got, want, what="local attention vs oracle (ragged, sparse neighbourhoods)")
    m.train()
    with pytest.raises(Exception):
        m(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))


def test_chebyshev_kernel_module_surface():
    """State-dict keys / shapes of the reference's LearnableChebyshevKernelAttention; the initial coefficients are the
    exp(-s) expansion, so a freshly built chebyshev model scores like the Gaussian kernel (tests/test_kernel_attention.py:163-208)."""
    g = load_golden("tiny_ad_chebyshev")
    m, sd = build_model(TINY_C, "fp32", int(g["weight_seed"]))
    assert set(m.state_dict().keys()) == set(sd.keys())
    assert m.state_dict()["flow.chain.2.shift_transformer.encoder_layers.1.self_attn.attention.cheb_coeffs"].shape == (2, 6)
    import timewarp_b200 as tw
    from tests.common import model_config
    import dataclasses
    full = dataclasses.replace(TINY_C, cheb_order=32, force_asymptotic_zero=False)
    fresh = tw.custom_transformer_nvp_constructor(model_config(full, "fp32")).cuda().eval()
    gauss = tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32")).cuda().eval()
    gauss.load_state_dict({k: v for k, v in fresh.state_dict().items() if not k.endswith("cheb_coeffs")})
    kw = _kw(g)
    with torch.no_grad():
        a = fresh.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **kw)
        b = gauss.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **kw)
    assert_rel(a, b, rel=1e-4, what="initial Chebyshev coefficients == Gaussian kernel")
    m.train()  # (tiny layer sizes / fp32: no backward kernels -- chebyshev training at full size is in test_gpu_train.py)
    with pytest.raises(Exception):
        m(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **kw)


@torch.no_grad()
def test_reference_checkpoint_reproduces_reference_log_likelihood():
    """End to end drop-in: a checkpoint written by the reference's save_model (tests/golden/ckpt) loaded through
    checkpoint.load_model gives the log-likelihoods the reference computed with those weights (tiny_ad golden)."""
    import os
    import timewarp_b200 as tw
    from tests.common import GOLDEN, model_config
    from timewarp_b200 import checkpoint as ck

    m = ck.load_model(os.path.join(GOLDEN, "ckpt"), lambda data: tw.custom_transformer_nvp_constructor(model_config(TINY_O, "fp32")),
                      weights_only=True).cuda().eval()
    g = load_golden("tiny_ad")
    ll = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert_rel(ll, g["log_likelihood"], what="log_likelihood from a reference checkpoint")


@pytest.mark.parametrize("attention_type", ["local", "kernel", "learnable_kernel"])
def test_batching_reference_models(attention_type):
    """The reference's own batching test (tests/test_batching.py:50-122,132-177) with its model configurations -- d_model 4,
    dim_feedforward 8, two hidden MLP layers [8, 8], 2 coupling x 2 transformer layers; local: 2 heads, max_radius 0.5;
    kernel / learnable_kernel: lengthscales [0.1, 0.2, 0.5, 1.0] -- on a ragged batch of two peptides (22 and 65 atoms):
    batched log_likelihood == per-sample loop at rtol / atol 1e-4, and == the oracle."""
    import timewarp_b200 as tw
    from timewarp_b200.peptides import alanine_dipeptide, tetrapeptide_2olx

    local = attention_type == "local"
    enc = tw.CustomAttentionEncoderLayerConfig(
        d_model=4, dim_feedforward=8, dropout=0.0, num_heads=2 if local else 4, attention_type=attention_type,
        lengthscales=None if local else [0.1, 0.2, 0.5, 1.0], max_radius=0.5 if local else None,
        normalise_kernel_values=None if local else False)
    cfg = tw.CustomAttentionTransformerNVPConfig(atom_embedding_dim=4, latent_mlp_hidden_dims=[8, 8], num_coupling_layers=2,
                                                 num_transformer_layers=2, encoder_layer_config=enc, precision="fp32")
    torch.manual_seed(1)
    m = tw.custom_transformer_nvp_constructor(cfg).cuda().eval()
    peps = [alanine_dipeptide(), tetrapeptide_2olx()]
    V = max(p.num_atoms for p in peps)
    gen = torch.Generator().manual_seed(2)
    at = torch.zeros(2, V, dtype=torch.long)
    xc, mask = torch.zeros(2, V, 3), torch.ones(2, V, dtype=torch.bool)
    for b, p in enumerate(peps):
        n = p.num_atoms
        at[b, :n], xc[b, :n], mask[b, :n] = torch.tensor(p.atom_types), torch.tensor(p.coords_nm, dtype=torch.float32), False
    xv, yv = torch.randn(2, V, 3, generator=gen), torch.randn(2, V, 3, generator=gen)
    yc = xc + 0.02 * torch.randn(2, V, 3, generator=gen)

    def ll_of(rows, n):
        return m.log_likelihood(atom_types=at[rows, :n].cuda(), x_coords=xc[rows, :n].cuda(), x_velocs=xv[rows, :n].cuda(),
                                y_coords=yc[rows, :n].cuda(), y_velocs=yv[rows, :n].cuda(), adj_list=EMPTY_ADJ.cuda(),
                                edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask[rows, :n].cuda())

    with torch.no_grad():
        batched = ll_of(slice(0, 2), V)
        for b, p in enumerate(peps):
            torch.testing.assert_close(ll_of(slice(b, b + 1), p.num_atoms)[0], batched[b], rtol=1e-4, atol=1e-4)
    ocfg = fo.OracleConfig(atom_embedding_dim=4, latent_mlp_hidden_dims=[8, 8], num_coupling_layers=2, num_transformer_layers=2,
                           d_model=4, dim_feedforward=8, lengthscales=[] if local else [0.1, 0.2, 0.5, 1.0],
                           attention_type=attention_type, max_radius=0.5 if local else 0.0, num_heads=2 if local else 0)
    want = fo.log_likelihood({k: v.detach().cpu() for k, v in m.state_dict().items()}, ocfg, at, xc, xv, yc, yv, mask)
    assert_rel(batched, want, what=f"{attention_type}: reference test configuration vs oracle")

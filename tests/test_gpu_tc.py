"""Tensor-core (tcgen05) path: descriptor conventions on the real chip, and parity of the bf16x3 /
bf16 conditioner kernels against the fp32 CUDA-core path, the golden vectors and the CPU oracle."""
import pytest
import torch

from oracle import flow_oracle as fo
from tests.common import EMPTY_ADJ, EMPTY_EBI, FULL_O, build_model, load_golden
from tests.test_gpu_flow import _kw, assert_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mode,b_mode,N,K,d_col", [(0, 0, 128, 128, 0), (0, 0, 256, 128, 128), (3, 0, 128, 128, 384), (0, 1, 80, 80, 8),
                                                     (1, 1, 48, 48, 0), (2, 2, 128, 64, 0), (3, 1, 64, 256, 64)])
def test_umma_probe(a_mode, b_mode, N, K, d_col):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from umma_probe import run
    err, status, mx = run(N, K, a_mode, b_mode, d_col)
    assert status == 0 and err < 1e-3 * mx


@pytest.mark.parametrize("name", ["full_ad22", "full_ad22_ragged", "full_2olx65"])
def test_bf16x3_matches_golden(name):
    g = load_golden(name)
    m, _ = build_model(FULL_O, "bf16x3", int(g["weight_seed"]))
    ll = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    err64 = assert_rel(ll, g["log_likelihood_f64"], what="bf16x3 ll vs reference fp64")
    assert_rel(ll, g["log_likelihood"], what="bf16x3 ll vs reference fp32")
    assert ll.requires_grad  # grad mode: the taped forward (two-kernel attention, tape written)
    with torch.no_grad():  # inference path: fused attention kernel, nothing taped
        ll_inf = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert not ll_inf.requires_grad
    assert_rel(ll_inf, g["log_likelihood_f64"], what="bf16x3 inference-path ll vs reference fp64")
    assert_rel(ll_inf, ll.detach(), rel=1e-5, what="inference path vs taped path")
    m32, _ = build_model(FULL_O, "fp32", int(g["weight_seed"]))
    ll32 = m32.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert_rel(ll, ll32, rel=2e-5, what="bf16x3 vs fp32 CUDA path")
    print(name, "bf16x3 rel err vs fp64:", err64)
    yc, yv, lp = m.sample_from_latents(g["atom_types"].cuda(), g["x_coords"].cuda(), g["x_velocs"].cuda(), g["masked_elements"].cuda(),
                                       g["s1_z_coords"].cuda(), g["s1_z_velocs"].cuda())
    keep = (~g["masked_elements"])[None, :, :, None].expand_as(g["s1_y_coords"])
    assert_rel(yc.cpu()[keep], g["s1_y_coords"][keep], what="bf16x3 samples")
    assert_rel(lp, g["s1_logp"], what="bf16x3 sample logp")


def test_bf16x3_many_tiles_and_tail():
    """Several 128-token tiles per CTA + a ragged tail tile, against the fp32 CUDA-core path."""
    torch.manual_seed(0)
    B, V = 300, 22  # 6600 tokens = 51 full tiles + a 72-row tail
    m, sd = build_model(FULL_O, "bf16x3", 1)
    m32, _ = build_model(FULL_O, "fp32", 1)
    at = torch.randint(0, 5, (B, V), device="cuda")
    x, xv, y, yv = (torch.randn(B, V, 3, device="cuda") * s for s in (0.3, 1.0, 0.3, 1.0))
    mask = torch.zeros(B, V, dtype=torch.bool, device="cuda")
    kw = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(),
              masked_elements=mask)
    a, b = m.log_likelihood(**kw), m32.log_likelihood(**kw)
    assert_rel(a, b, rel=2e-5, what="bf16x3 vs fp32, 52 tiles")
    with torch.no_grad():
        assert_rel(m.log_likelihood(**kw), b, rel=2e-5, what="bf16x3 inference path vs fp32, 52 tiles")
    # weights modified in place are re-packed
    with torch.no_grad():
        m.flow.chain[0].scale_transformer.encoder_layers[0].linear1.weight.mul_(1.5)
        m32.flow.chain[0].scale_transformer.encoder_layers[0].linear1.weight.mul_(1.5)
    assert_rel(m.log_likelihood(**kw), m32.log_likelihood(**kw), rel=2e-5, what="after in-place weight update")


@pytest.mark.parametrize("V", [17, 31, 33, 47, 49, 63, 64, 66, 79, 81, 97, 127])
@torch.no_grad()
def test_attention_kernels_sweep_vs_fp32_path(V):
    """Every group size / buffer plan of the two feature-major attention kernels (k_attn_fm3 for V <= 80: G = 80 // VP samples per
    group; k_attn_fm above), with sample counts that leave a partial last group, an odd number of groups (a CTA without work in a
    pair), more groups than CTAs, ragged masks -- and proposals from ONE conditioning state (shared score images) -- against the
    fp32 CUDA-core path of the same model."""
    torch.manual_seed(V)
    m, _ = build_model(FULL_O, "bf16x3", 5)
    m32, _ = build_model(FULL_O, "fp32", 5)
    for B in (1, 2, 7, 151 if V <= 33 else 11):
        lengths = torch.randint(max(1, V // 2), V + 1, (B,))
        lengths[0] = V
        mask = (torch.arange(V)[None, :] >= lengths[:, None]).cuda()
        keep = (~mask)[:, :, None]
        at = torch.randint(0, 5, (B, V), device="cuda") * (~mask)
        x, xv, y, yv = (torch.randn(B, V, 3, device="cuda") * s * keep for s in (0.3, 1.0, 0.3, 1.0))
        kw = dict(atom_types=at, x_coords=x, x_velocs=xv, adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask)
        assert_rel(m.log_likelihood(y_coords=y, y_velocs=yv, **kw), m32.log_likelihood(y_coords=y, y_velocs=yv, **kw), rel=2e-5,
                   what=f"V={V} B={B}: bf16x3 vs fp32 path")
    # S proposals from one state: the conditioning (and its score images) is shared by all S samples
    S = 5
    zc, zv = torch.randn(S, 1, V, 3, device="cuda") * 0.01, torch.randn(S, 1, V, 3, device="cuda")
    kw1 = dict(atom_types=at[:1], x_coords=x[:1], x_velocs=xv[:1], masked_elements=mask[:1], z_coords=zc, z_velocs=zv)
    a = m.sample_from_latents(**kw1)
    b = m32.sample_from_latents(**kw1)
    assert_rel(a[2], b[2], rel=2e-5, what=f"V={V}: S proposals, log p")
    keep1 = (~mask[:1])[None, :, :, None].expand_as(a[0])
    assert_rel(a[0][keep1], b[0][keep1], rel=1e-4, what=f"V={V}: S proposals, y_coords")  # (max-norm relative, like the golden tests)


def test_bf16_plain_is_close():
    g = load_golden("full_ad22")
    m, _ = build_model(FULL_O, "bf16", 0)
    ll = m.log_likelihood(y_coords=g["y_coords"].cuda(), y_velocs=g["y_velocs"].cuda(), **_kw(g))
    assert_rel(ll, g["log_likelihood"], rel=2e-2, what="plain bf16 (training precision)")


@pytest.mark.parametrize("B,V,lengths", [(37, 65, None), (9, 65, [65, 64, 33, 65, 1, 17, 65, 48, 65]), (3, 80, [80, 79, 66]), (130, 22, None), (150, 65, None),
                                         # the feature-major attention kernel at every group size: G = 10 (16 atoms), 3 (48), 1 (100, 128);
                                         # batch sizes that leave a partial last group
                                         (11, 16, [16, 15, 1, 16, 9, 16, 16, 3, 16, 16, 2]), (7, 48, None), (5, 100, [100, 99, 81, 100, 97]),
                                         (3, 128, [128, 127, 113]), (1, 65, None),
                                         # more than 128 atoms (1hgv has 691): attention on the CUDA-core kernels, MLPs and FFN on tcgen05
                                         (2, 150, [150, 131]), (1, 691, None)])
def test_bf16x3_inference_path_odd_sizes_vs_oracle(B, V, lengths):
    """The inference kernels (CTA-pair FFN with a partial last 256-token tile, fused attention with ragged / masked
    samples, samples straddling tile boundaries; 150 x 65 tokens = 39 pair tiles over 37 pairs: two leftover tiles split
    along the hidden dimension) against the CPU oracle: log_likelihood, and sample -> density round trip."""
    torch.manual_seed(B * 1000 + V)
    mask = torch.zeros(B, V, dtype=torch.bool)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    m, sd = build_model(FULL_O, "bf16x3", 3)
    ll_ref = fo.log_likelihood(sd, FULL_O, at, x, xv, y, yv, mask, distance_mode="direct")
    kw = dict(atom_types=at.cuda(), x_coords=x.cuda(), x_velocs=xv.cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(),
              masked_elements=mask.cuda())
    with torch.no_grad():
        ll = m.log_likelihood(y_coords=y.cuda(), y_velocs=yv.cuda(), **kw)
        assert_rel(ll, ll_ref, what="inference path vs oracle")
        yc, yvel, lp = m.conditional_sample_with_logp(num_samples=1, **kw)
        ll2 = m.log_likelihood(y_coords=yc[0], y_velocs=yvel[0], **kw)
    torch.testing.assert_close(ll2, lp[0], rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("name", ["full_ad22", "full_2olx65"])
@torch.no_grad()
def test_bf16x3_shared_conditioning_proposals(name):
    """S proposals from ONE conditioning state (sample_with_model shape: n_cond = 1, every sample shares the score images)
    and the reverse-move density of those proposals (distinct conditioning per sample), inference kernels vs the reference."""
    g = load_golden(name)
    m, _ = build_model(FULL_O, "bf16x3", int(g["weight_seed"]))
    mask = g["masked_elements"]
    S = g["sS_z_coords"].shape[0]
    yc, yv, lp = m.sample_from_latents(g["atom_types"][:1].cuda(), g["x_coords"][:1].cuda(), g["x_velocs"][:1].cuda(), mask[:1].cuda(),
                                       g["sS_z_coords"].cuda(), g["sS_z_velocs"].cuda())
    keepS = (~mask[:1])[None, :, :, None].expand_as(g["sS_y_coords"])
    assert_rel(yc.cpu()[keepS], g["sS_y_coords"][keepS], what="S proposals: y_coords")
    assert_rel(lp, g["sS_logp"], what="S proposals: logp")
    p_yx = m.log_likelihood(atom_types=g["atom_types"][:1].repeat(S, 1).cuda(), y_coords=g["x_coords"][:1].repeat(S, 1, 1).cuda(),
                            y_velocs=g["x_velocs"][:1].repeat(S, 1, 1).cuda(), x_coords=yc.squeeze(1), x_velocs=yv.squeeze(1),
                            adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask[:1].repeat(S, 1).cuda())
    assert_rel(p_yx, g["sS_p_yx"], what="reverse-move density")
    # a larger S: 300 proposals from one state == the same proposals evaluated as 300 independent samples
    torch.manual_seed(5)
    S2 = 300
    V = g["x_coords"].shape[1]
    zc, zv = 0.05 * torch.randn(S2, 1, V, 3, device="cuda"), torch.randn(S2, 1, V, 3, device="cuda")
    a = m.sample_from_latents(g["atom_types"][:1].cuda(), g["x_coords"][:1].cuda(), g["x_velocs"][:1].cuda(), mask[:1].cuda(), zc, zv)
    b = m.sample_from_latents(g["atom_types"][:1].repeat(S2, 1).cuda(), g["x_coords"][:1].repeat(S2, 1, 1).cuda(),
                              g["x_velocs"][:1].repeat(S2, 1, 1).cuda(), mask[:1].repeat(S2, 1).cuda(),
                              zc.transpose(0, 1).contiguous(), zv.transpose(0, 1).contiguous())
    torch.testing.assert_close(a[0].squeeze(1), b[0].squeeze(0), rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(a[2].squeeze(1), b[2].squeeze(0), rtol=1e-5, atol=1e-3)


@torch.no_grad()
def test_bench_size_properties():
    """BASELINE.json configs[2] at FULL size (1024 chains x 65 atoms, bf16x3 inference kernels) through size-independent
    properties: proposal -> density round trip, invariance of every row to how the batch is cut, the MH step's reverse-move
    density == per-row evaluation, and the CPU oracle on a handful of rows."""
    from timewarp_b200.peptides import tetrapeptide_2olx

    pep = tetrapeptide_2olx()
    B, V = 1024, pep.num_atoms
    m, sd = build_model(FULL_O, "bf16x3", 0)
    gen = torch.Generator().manual_seed(11)
    x = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, V, 3, generator=gen)
    xv = torch.randn(B, V, 3, generator=gen)
    at = torch.tensor(pep.atom_types)[None].repeat(B, 1)
    mask = torch.zeros(B, V, dtype=torch.bool)
    zc, zv = 0.05 * torch.randn(1, B, V, 3, generator=gen), torch.randn(1, B, V, 3, generator=gen)
    kw = dict(adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda())
    yc, yv, lp = m.sample_from_latents(at.cuda(), x.cuda(), xv.cuda(), mask.cuda(), zc.cuda(), zv.cuda())
    assert yc.shape == (1, B, V, 3) and lp.shape == (1, B) and torch.isfinite(lp).all()
    # (1) proposal -> density round trip on all 1024 chains
    ll = m.log_likelihood(atom_types=at.cuda(), x_coords=x.cuda(), x_velocs=xv.cuda(), y_coords=yc[0], y_velocs=yv[0],
                          masked_elements=mask.cuda(), **kw)
    assert_rel(ll, lp[0], what="1024 x 65: sample -> density round trip")
    # (2) a row does not depend on how the batch is cut (tiles / CTA assignment / tail handling differ)
    for rows in (slice(0, 1), slice(100, 357), slice(1000, 1024)):
        part = m.log_likelihood(atom_types=at[rows].cuda(), x_coords=x[rows].cuda(), x_velocs=xv[rows].cuda(), y_coords=yc[0, rows],
                                y_velocs=yv[0, rows], masked_elements=mask[rows].cuda(), **kw)
        torch.testing.assert_close(part, ll[rows], rtol=1e-6, atol=2e-4)  # |ll| ~ 1e3: a few fp32 ulps
    # (3) the reverse-move density of the MH rule (conditioning = the proposals) is finite and row-wise reproducible
    p_yx = m.log_likelihood(atom_types=at.cuda(), x_coords=yc[0], x_velocs=yv[0], y_coords=x.cuda(), y_velocs=xv.cuda(),
                            masked_elements=mask.cuda(), **kw)
    one = m.log_likelihood(atom_types=at[7:8].cuda(), x_coords=yc[0, 7:8], x_velocs=yv[0, 7:8], y_coords=x[7:8].cuda(),
                           y_velocs=xv[7:8].cuda(), masked_elements=mask[7:8].cuda(), **kw)
    assert torch.isfinite(p_yx).all()
    torch.testing.assert_close(one, p_yx[7:8], rtol=1e-6, atol=2e-4)
    # (4) the CPU oracle on five rows spread over the batch
    pick = torch.tensor([0, 255, 511, 768, 1023])
    want = fo.log_likelihood(sd, FULL_O, at[pick], x[pick], xv[pick], yc[0].cpu()[pick], yv[0].cpu()[pick], mask[pick], distance_mode="direct")
    assert_rel(ll.cpu()[pick], want, what="1024 x 65: oracle on five rows")

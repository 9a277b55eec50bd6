"""Training path on the GPU: the generic operand-image GEMM in every mode, and the hand-written backward of
log_likelihood / the NLL loss against (a) gradients from `loss.backward()` on the unmodified reference
(tests/golden/grads_*.npz) and (b) the CPU oracle's autograd on seeded ragged batches."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from tests.common import EMPTY_ADJ, EMPTY_EBI, FULL_C, FULL_L, FULL_LOC, FULL_O, GOLDEN, build_model
from timewarp_b200 import _lib

pytestmark = pytest.mark.gpu

# Gradient tolerance.  Forward values agree with the reference to ~1e-6 (bf16x3 = 16-17 significand bits per operand,
# fp32 accumulation); through the ~150 chained contractions of the backward pass the per-tensor error is 3e-5..1e-4
# (median over tensors, measured against the oracle's fp64 autograd with tools/grad_diag.py; the reference's own
# fp32 autograd is at 3e-7).  The largest per-tensor errors are 1e-3 (in_mlp) and 5e-3 (FFN linear1 rows whose ReLU
# input is within rounding of zero, so the unit is on in one arithmetic and off in the other).
GRAD_RTOL = 2e-3
GRAD_RTOL_RELU = 1e-2  # FFN linear1.{weight,bias}: gradients behind the ReLU boundary
GRAD_MEDIAN_RTOL = 2e-4


def _tol(name):
    return GRAD_RTOL_RELU if ".linear1." in name else GRAD_RTOL


def _gemm(mode, A, B, c_shape, bn=128, splits=1, precision="bf16x3", init=None):
    lib = _lib.load()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    Cm = torch.zeros(c_shape, device="cuda") if init is None else init.clone()
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    _lib.check(lib.tw_debug_gemm(mode, _lib.PRECISION[precision], A.data_ptr(), A.shape[0], A.shape[1], B.data_ptr(), B.shape[0], B.shape[1],
                                 Cm.data_ptr(), bn, splits, base, ws.numel() - 1024, torch.cuda.current_stream().cuda_stream), "tw_debug_gemm")
    torch.cuda.synchronize()
    return Cm


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (128, 128, 64), (1000, 2048, 128), (77, 128, 768)])
def test_gemm_nt(M, N, K):
    torch.manual_seed(0)
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    assert _rel(_gemm(0, A, B, (M, N)), A.double() @ B.double().T) < 2e-5


@pytest.mark.parametrize("M,K,N,bn", [(300, 128, 2048, 128), (513, 2048, 128, 128), (200, 256, 64, 64), (90, 256, 128, 128)])
def test_gemm_nn(M, K, N, bn):
    torch.manual_seed(1)
    A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    assert _rel(_gemm(1, A, B, (M, N), bn=bn), A.double() @ B.double()) < 2e-5


def test_gemm_nn_headed():
    torch.manual_seed(2)
    H, M = 6, 333
    A, B = torch.randn(M, H * 128, device="cuda"), torch.randn(128, H * 128, device="cuda")
    ref = sum(A[:, h * 128:(h + 1) * 128].double() @ B[:, h * 128:(h + 1) * 128].double() for h in range(H))
    assert _rel(_gemm(2, A, B, (M, 128)), ref) < 2e-5


@pytest.mark.parametrize("M,I,J,bn,splits", [(300, 128, 2048, 128, 4), (1111, 2048, 128, 128, 3), (500, 256, 64, 64, 5), (64, 128, 768, 128, 1)])
def test_gemm_tn_accumulates(M, I, J, bn, splits):
    torch.manual_seed(3)
    A, B = torch.randn(M, I, device="cuda"), torch.randn(M, J, device="cuda")
    init = torch.randn(I, J, device="cuda")
    out = _gemm(3, A, B, (I, J), bn=bn, splits=splits, init=init)
    assert _rel(out, init.double() + A.double().T @ B.double()) < 2e-5


def test_gemm_plain_bf16():
    torch.manual_seed(4)
    A, B = torch.randn(256, 128, device="cuda"), torch.randn(128, 128, device="cuda")
    ref = A.bfloat16().double() @ B.bfloat16().double().T
    assert _rel(_gemm(0, A, B, (256, 128), precision="bf16"), ref) < 1e-5


def _loss_and_grads(model, g):
    model.train()
    model.zero_grad(set_to_none=True)
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    loss = model(**kw)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}


def _load(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (torch.from_numpy(d[k]) if d[k].dtype.kind != "U" else d[k]) for k in d.files}


@pytest.mark.parametrize("name", ["grads_full_ad22", "grads_full_ad22_ragged", "grads_full_ad22_learnable", "grads_full_ad22_chebyshev",
                                  "grads_full_ad22_local"])
def test_backward_matches_reference_gradients(name):
    g = _load(name)
    learnable = name.endswith("learnable")
    m, _ = build_model(FULL_L if learnable else (FULL_C if name.endswith("chebyshev") else (FULL_LOC if name.endswith("local") else FULL_O)),
                       "bf16x3", int(g["weight_seed"]))
    loss, grads = _loss_and_grads(m, g)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    names = [str(n) for n in g["grad_names"]]
    if learnable:
        # the reference leaves .grad = None on every log_lengthscales but the first executed layer's (recorded norm 0); so do we
        unused = {n for n, v in zip(names, g["grad_norms"].tolist()) if n.endswith("log_lengthscales") and v == 0.0}
        assert len(unused) == 47 and not (unused & set(grads))
        names = [n for n in names if n not in unused]
        g["grad_norms"] = torch.tensor([v for n, v in zip([str(n) for n in g["grad_names"]], g["grad_norms"].tolist()) if n not in unused])
    assert set(names) == set(grads)
    worst = 0.0
    for n, ref_norm in zip(names, g["grad_norms"].tolist()):
        got = float(grads[n].double().norm())
        worst = max(worst, abs(got - ref_norm) / max(ref_norm, 1e-6))
        assert abs(got - ref_norm) <= _tol(n) * ref_norm + 1e-6, (n, got, ref_norm)
    for k in g:
        if k.startswith("grad::"):
            got = grads[k[6:]]
            got = got[:8] if got.numel() > 20000 else got
            err = _rel(got, g[k]) if float(g[k].norm()) > 0 else float(got.norm())
            assert err < _tol(k), (k, err)
    print(name, "worst gradient-norm rel err", worst)


def test_backward_matches_oracle_autograd_every_tensor():
    """Every parameter gradient, tensor by tensor, on a ragged batch that is not a golden case."""
    torch.manual_seed(7)
    B, V = 5, 30
    lengths = [30, 22, 17, 30, 9]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    g = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, masked_elements=mask)
    m, sd = build_model(FULL_O, "bf16x3", 2)
    loss, grads = _loss_and_grads(m, g)
    # fp64 autograd of the oracle = the true gradient (an fp32 oracle run has ReLU-boundary flips of its own)
    loss_ref, grads_ref = fo.nll_loss_and_grads(fo.to_dtype(sd, torch.float64), FULL_O, at, x.double(), xv.double(), y.double(),
                                                yv.double(), mask, distance_mode="direct")
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    total = float(torch.sqrt(sum(v.double().norm() ** 2 for v in grads_ref.values())))
    worst, errs = ("", 0.0), []
    for k, ref in grads_ref.items():
        err = float((grads[k].double() - ref.double()).norm())
        scale = max(float(ref.double().norm()), 1e-4 * total)
        errs.append(err / scale)
        if err / scale > worst[1]:
            worst = (k, err / scale)
        assert err <= _tol(k) * scale, (k, err, float(ref.norm()))
    assert float(np.median(errs)) < GRAD_MEDIAN_RTOL
    print("worst per-tensor gradient rel err", worst, "median", float(np.median(errs)))


def test_large_batch_gradient_is_the_mean_of_chunk_gradients():
    """Size-independent property at a size the oracle would take minutes for: the mean-NLL gradient of 96 ragged samples
    (2112 tokens = 17 token tiles: every CTA of the fused FFN-backward kernel walks over several token tiles, the weight-gradient
    GEMMs split their contraction) equals the mean of the gradients of its 8 chunks of 12 samples (3 token tiles each)."""
    torch.manual_seed(11)
    B, V, n_chunks = 96, 22, 8
    lengths = torch.randint(9, V + 1, (B,))
    lengths[::12] = V
    mask = torch.arange(V)[None, :] >= lengths[:, None]
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    g = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, masked_elements=mask)
    m, _ = build_model(FULL_O, "bf16x3", 2)
    loss, grads = _loss_and_grads(m, g)
    acc, loss_acc = None, 0.0
    per = B // n_chunks
    for c in range(n_chunks):
        gc = {k: v[c * per:(c + 1) * per] for k, v in g.items()}
        lc, gr = _loss_and_grads(m, gc)
        loss_acc += float(lc) / n_chunks
        acc = {k: v.double() / n_chunks for k, v in gr.items()} if acc is None else {k: acc[k] + gr[k].double() / n_chunks for k in acc}
    assert abs(float(loss) - loss_acc) < 1e-5 * abs(loss_acc)
    total = float(torch.sqrt(sum(v.norm() ** 2 for v in acc.values())))
    errs = []
    for k, ref in acc.items():
        err = float((grads[k].double() - ref).norm())
        scale = max(float(ref.norm()), 1e-4 * total)
        errs.append(err / scale)
        assert err <= 2e-4 * scale, (k, err / scale)
    assert float(np.median(errs)) < 2e-5, float(np.median(errs))


def test_local_attention_gradients_match_oracle_autograd():
    """`local` attention trains (local_self_attention.py:46-119; the reference's tests/test_batching.py:132-177 runs train-capable
    models of all three attention types): every parameter gradient -- qkv_proj / output_proj included -- against the oracle's
    fp64 autograd on a ragged batch with sparse neighbourhoods (max_radius 0.45 nm, positions spread over ~1 nm)."""
    torch.manual_seed(17)
    B, V = 4, 26
    lengths = [26, 19, 26, 7]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    g = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, masked_elements=mask)
    m, sd = build_model(FULL_LOC, "bf16x3", 4)
    loss, grads = _loss_and_grads(m, g)
    loss_ref, grads_ref = fo.nll_loss_and_grads(fo.to_dtype(sd, torch.float64), FULL_LOC, at, x.double(), xv.double(), y.double(),
                                                yv.double(), mask, distance_mode="direct")
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    assert any(k.endswith("qkv_proj.weight") for k in grads_ref) and set(grads_ref) == set(grads)
    total = float(torch.sqrt(sum(v.double().norm() ** 2 for v in grads_ref.values())))
    worst, errs = ("", 0.0), []
    for k, ref in grads_ref.items():
        err = float((grads[k].double() - ref.double()).norm())
        scale = max(float(ref.double().norm()), 1e-4 * total)
        errs.append(err / scale)
        if err / scale > worst[1]:
            worst = (k, err / scale)
        assert err <= _tol(k) * scale, (k, err, float(ref.norm()))
    assert float(np.median(errs)) < GRAD_MEDIAN_RTOL
    print("local attention: worst per-tensor gradient rel err", worst, "median", float(np.median(errs)))


def test_learnable_lengthscale_gradient_matches_oracle_autograd():
    """learnable_kernel: d(loss)/d(log_lengthscales) of the first executed attention layer on a ragged batch (padding in the
    key mask, several tiles), against the oracle's fp64 autograd; frozen lengthscales skip the extra kernels."""
    torch.manual_seed(9)
    B, V = 6, 40
    lengths = [40, 22, 31, 40, 9, 17]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.25 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    g = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, masked_elements=mask)
    m, sd = build_model(FULL_L, "bf16x3", 4)
    loss, grads = _loss_and_grads(m, g)
    loss_ref, grads_ref = fo.nll_loss_and_grads(fo.to_dtype(sd, torch.float64), FULL_L, at, x.double(), xv.double(), y.double(),
                                                yv.double(), mask, distance_mode="direct")
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    key = "flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"
    ref = grads_ref[key].double()
    assert float(ref.norm()) > 0
    err = float((grads[key].double() - ref).norm() / ref.norm())
    assert err < GRAD_RTOL, (err, grads[key], ref)
    assert [k for k in grads if k.endswith("log_lengthscales")] == [key]
    # the other gradients are unaffected by the extra kernels
    k2 = "flow.chain.3.shift_transformer.encoder_layers.1.self_attn.values_proj.weight"
    assert _rel(grads[k2], grads_ref[k2]) < GRAD_RTOL
    # frozen lengthscales: no gradient, same loss
    for n, p in m.named_parameters():
        if n.endswith("log_lengthscales"):
            p.requires_grad_(False)
    loss2, grads2 = _loss_and_grads(m, g)
    assert float(loss2) == float(loss) and not any(k.endswith("log_lengthscales") for k in grads2)
    assert _rel(grads2[k2], grads[k2]) < 1e-5  # (weight gradients accumulate with fp32 atomics: not bit-reproducible)


def test_chebyshev_coefficient_gradients_match_oracle_autograd():
    """chebyshev_kernel training with the expansion forced to vanish at infinity (coefficient-mean subtraction in the chain
    rule) on a ragged batch: every cheb_coeffs tensor and a few others against the oracle's fp64 autograd."""
    import dataclasses
    torch.manual_seed(19)
    B, V = 5, 28
    lengths = [28, 20, 13, 28, 7]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.25 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    cfg = dataclasses.replace(FULL_C, cheb_order=9, force_asymptotic_zero=True)
    g = dict(atom_types=at, x_coords=x, x_velocs=xv, y_coords=y, y_velocs=yv, masked_elements=mask)
    m, sd = build_model(cfg, "bf16x3", 8)
    loss, grads = _loss_and_grads(m, g)
    loss_ref, grads_ref = fo.nll_loss_and_grads(fo.to_dtype(sd, torch.float64), cfg, at, x.double(), xv.double(), y.double(), yv.double(),
                                                mask, distance_mode="direct")
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    cheb = [k for k in grads_ref if k.endswith("cheb_coeffs")]
    assert len(cheb) == 48
    total = float(torch.sqrt(sum(grads_ref[k].double().norm() ** 2 for k in cheb)))
    for k in cheb + ["flow.chain.3.shift_transformer.encoder_layers.1.self_attn.values_proj.weight", "flow.atom_embedder.weight"]:
        ref = grads_ref[k].double()
        err = float((grads[k].double() - ref).norm())
        assert err <= GRAD_RTOL * max(float(ref.norm()), 1e-2 * total), (k, err, float(ref.norm()))
    # frozen coefficients: the extra kernels are skipped, the other gradients do not move
    for n, p in m.named_parameters():
        if n.endswith("cheb_coeffs"):
            p.requires_grad_(False)
    _, grads2 = _loss_and_grads(m, g)
    assert not any(k.endswith("cheb_coeffs") for k in grads2)
    assert _rel(grads2["flow.atom_embedder.weight"], grads["flow.atom_embedder.weight"]) < 1e-5


def test_inference_result_unchanged_and_no_grad_path():
    g = _load("grads_full_ad22")
    m, _ = build_model(FULL_O, "bf16x3", 0)
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    with torch.no_grad():
        a = m.log_likelihood(**kw)
    b = m.log_likelihood(**kw)  # taped forward
    assert b.requires_grad and not a.requires_grad
    torch.testing.assert_close(a, b.detach(), rtol=1e-6, atol=1e-4)


def test_fp32_precision_has_no_training_path():
    g = _load("grads_full_ad22")
    m, _ = build_model(FULL_O, "fp32", 0)
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    m.train()
    with pytest.raises(_lib.TimewarpB200Error):
        m(**kw)
    m.eval()  # evaluation mode: the result is computed, but carries no graph
    assert not m.log_likelihood(**kw).requires_grad


def test_fused_optimizer_updates_reach_the_packed_weights():
    """torch.optim.Adam(fused=True) updates the parameters without bumping their version counters: the taped forward must
    re-pack the bf16 weight images anyway (regression: the big GEMM weights stayed stale and only biases / LayerNorm trained)."""
    g = _load("grads_full_ad22")
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    traj = {}
    for name, okw in (("foreach", dict()), ("fused", dict(fused=True))):
        m, _ = build_model(FULL_O, "bf16x3", 0)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, **okw)
        losses = []
        for _ in range(4):
            opt.zero_grad(set_to_none=True)
            loss = m(**kw)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        traj[name] = losses
        with torch.no_grad():  # the inference path sees the trained weights too
            m.eval()
            traj[name + "_eval"] = float(m(**kw))
    assert abs(traj["foreach"][0] - traj["fused"][0]) < 1e-6
    assert abs(traj["foreach"][1] - traj["foreach"][0]) > 1e-3  # the step really moves the loss
    for a, b in zip(traj["foreach"], traj["fused"]):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a)), traj
    assert abs(traj["foreach_eval"] - traj["fused_eval"]) < 2e-3 * max(1.0, abs(traj["foreach_eval"])), traj


def test_flat_adam_kernel_matches_torch_adam():
    """tw_adam_step against torch.optim.Adam (utilities/training_utils.py:356-368: lr + L2 weight decay) on identical
    gradients, six steps: parameters and both moments."""
    from timewarp_b200 import _lib as L
    n = 4 * 50_001
    g = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(n, device="cuda", generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    ref = torch.optim.Adam([ref_p], lr=3e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    hyper = torch.tensor([3e-4, 0.9, 0.99, 1e-8, 1e-2, 0.0, 0.0, 0.0], device="cuda")
    scale = 10.0 ** torch.randint(-6, 2, (n,), device="cuda", generator=g).float()  # per-element gradient magnitude
    for it in range(6):
        grad = torch.randn(n, device="cuda", generator=g) * scale
        ref_p.grad = grad.clone()
        ref.step()
        hyper[5] += 1
        L.check(L.load().tw_adam_step(L.ptr(p), L.ptr(grad), L.ptr(m), L.ptr(v), n, L.ptr(hyper), torch.cuda.current_stream().cuda_stream), "adam")
        st = ref.state[ref_p]
        # moments relative to the element's gradient scale (the weight-decay term adds up to 1e-2 |p| to the small ones)
        s1 = scale + 1e-2 * p0.abs()
        torch.testing.assert_close(m / s1, st["exp_avg"] / s1, rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(v / s1 ** 2, st["exp_avg_sq"] / s1 ** 2, rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(p, ref_p.detach(), rtol=1e-6, atol=2e-8)  # a few ulps of p; one step moves p by 3e-4


def test_flat_adam_trains_like_torch_adam():
    """optim.FlatAdam (parameters re-homed into one flat buffer, one launch per step) against torch.optim.Adam on the full
    model: same loss trajectory, state-dict keys and shapes untouched, the inference path sees the trained weights; the fast path
    (gradients read in place from the model's flat buffer) and the gather path (gradients that live elsewhere) agree."""
    from timewarp_b200.optim import FlatAdam
    g = _load("grads_full_ad22")
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    traj, finals = {}, {}
    for name in ("torch", "flat", "flat_gather"):
        m, _ = build_model(FULL_O, "bf16x3", 0)
        m.train()
        keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-3) if name == "torch" else FlatAdam(m, lr=1e-4, weight_decay=1e-3)
        losses = []
        for _ in range(4):
            opt.zero_grad(set_to_none=True)
            loss = m(**kw)
            loss.backward()
            if name == "flat_gather":  # gradients outside the model's flat buffer (what autograd does when it clones them)
                for p in m.parameters():
                    if p.grad is not None:
                        p.grad = p.grad.clone()
            opt.step()
            losses.append(float(loss.detach()))
        traj[name] = losses
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == keys
        finals[name] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        with torch.no_grad():
            m.eval()
            traj[name + "_eval"] = float(m(**kw))
        if name == "flat":
            assert (opt.fast_path_steps, opt.gather_steps) == (4, 0)
            assert all(p.data_ptr() >= opt._flat_p.data_ptr() for p in m.parameters() if p.requires_grad)
        if name == "flat_gather":
            assert (opt.fast_path_steps, opt.gather_steps) == (0, 4)
    assert abs(traj["torch"][0] - traj["flat"][0]) < 1e-6
    assert abs(traj["torch"][1] - traj["torch"][0]) > 1e-3
    for other in ("flat", "flat_gather"):
        for a, b in zip(traj["torch"], traj[other]):
            assert abs(a - b) < 2e-3 * max(1.0, abs(a)), traj
        assert abs(traj["torch_eval"] - traj[other + "_eval"]) < 2e-3 * max(1.0, abs(traj["torch_eval"])), traj
    # the big weights moved the same way (elements with a well-determined gradient sign: |delta| = lr-sized steps)
    k = "flow.chain.0.scale_transformer.encoder_layers.0.linear1.weight"
    k = k if k in finals["torch"] else next(x for x in finals["torch"] if x.endswith("linear1.weight"))
    for other in ("flat", "flat_gather"):
        d = (finals["torch"][k] - finals[other][k]).abs()
        assert float((d > 5e-5).float().mean()) < 0.05, float((d > 5e-5).float().mean())


def test_flat_adam_resumes_from_a_checkpoint():
    """utilities/model_utils.py:12-32 / tests/test_training_utils.py:23-67: a run resumed from (model, optimizer) state dicts -- the
    optimizer state loaded BEFORE its first step, into whatever flat layout the new process ends up with -- continues like the
    uninterrupted one."""
    from timewarp_b200.optim import FlatAdam
    g = _load("grads_full_ad22")
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())

    def steps(m, opt, n, clone_grads=False):
        out = []
        for _ in range(n):
            opt.zero_grad(set_to_none=True)
            loss = m(**kw)
            loss.backward()
            if clone_grads:  # forces the optimizer's own flat layout (gather path)
                for p in m.parameters():
                    if p.grad is not None:
                        p.grad = p.grad.clone()
            opt.step()
            out.append(float(loss.detach()))
        return out

    m, _ = build_model(FULL_O, "bf16x3", 0)
    m.train()
    opt = FlatAdam(m, lr=1e-4)
    first = steps(m, opt, 2)
    model_sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    opt_sd = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in opt.state_dict().items()}
    straight = steps(m, opt, 2)
    for clone_grads in (False, True):
        m2, _ = build_model(FULL_O, "bf16x3", 1)  # different weights until the checkpoint is loaded
        m2.load_state_dict(model_sd)
        m2.train()
        opt2 = FlatAdam(m2, lr=1e-4)
        opt2.load_state_dict(opt_sd)
        resumed = steps(m2, opt2, 2, clone_grads)
        assert float(opt2._hyper[5]) == 4.0
        for a, b in zip(straight, resumed):
            assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (first, straight, resumed)
    assert abs(straight[0] - first[1]) > 1e-4  # the run was still moving


def test_training_trajectory_matches_oracle():
    """three Adam steps of NLL training (forward + hand-written backward + optimizer, weights
    re-packed every step) against the same three steps of the CPU oracle (fp64 autograd + torch Adam): the loss trajectory."""
    g = _load("grads_full_ad22_ragged")
    kw = dict(atom_types=g["atom_types"].cuda(), x_coords=g["x_coords"].cuda(), x_velocs=g["x_velocs"].cuda(), y_coords=g["y_coords"].cuda(),
              y_velocs=g["y_velocs"].cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=g["masked_elements"].cuda())
    m, sd = build_model(FULL_O, "bf16x3", int(g["weight_seed"]))
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-5)
    ours = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        loss = m(**kw)
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))
    leaves = {k: v.double().clone().requires_grad_(not k.endswith(".lengthscales")) for k, v in sd.items()}
    opt_ref = torch.optim.Adam([v for v in leaves.values() if v.requires_grad], lr=1e-5)
    ref = []
    for _ in range(3):
        opt_ref.zero_grad(set_to_none=True)
        loss = fo.nll_loss(leaves, FULL_O, g["atom_types"], g["x_coords"].double(), g["x_velocs"].double(), g["y_coords"].double(),
                           g["y_velocs"].double(), g["masked_elements"], "direct")
        loss.backward()
        opt_ref.step()
        ref.append(float(loss.detach()))
    assert abs(ref[2] - ref[0]) > 0.05  # the three steps move the loss well beyond the tolerance
    for a, b in zip(ours, ref):
        assert abs(a - b) < 2e-4 * max(1.0, abs(b)), (ours, ref)


def _check_first_order_decrease(m, loss_value, evaluate, delta=2e-3):
    """Composed-gradient check: a plain gradient step p -= eta * g with eta = delta / |g|^2 must lower the loss by ~delta
    (first-order Taylor; `evaluate()` recomputes the loss on the same random draws)."""
    g2 = float(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
    eta = delta / g2
    with torch.no_grad():
        for p in m.parameters():
            if p.grad is not None:
                p.sub_(eta * p.grad)
    m.invalidate_packed_weights()
    after = evaluate()
    drop = loss_value - after
    assert 0.5 * delta < drop < 1.5 * delta, (loss_value, after, drop, delta)


def test_sampling_backward_matches_oracle_autograd():
    """conditional_sample_with_logp under autograd (the energy-based losses, losses.py:396-664): gradients of a generic scalar
    L = <y_coords, G1> + <y_velocs, G2> + <log p, g3> w.r.t. every parameter, against the oracle's fp64 autograd through its
    own sampling pass with the same latent draws -- ragged batch, prior log-scales included (the draws are eps * exp(log_scale))."""
    torch.manual_seed(13)
    B, V = 5, 30
    lengths = [30, 22, 17, 30, 9]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    xv = torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    eps_c, eps_v = torch.randn(1, B, V, 3), torch.randn(1, B, V, 3)
    G1, G2, g3 = torch.randn(B, V, 3) * keep, 0.1 * torch.randn(B, V, 3) * keep, torch.randn(B) / V
    m, sd = build_model(FULL_O, "bf16x3", 5)
    m.train()
    m.zero_grad(set_to_none=True)
    zc = eps_c.cuda() * torch.exp(m.coords_prior_log_scale)
    zv = eps_v.cuda() * torch.exp(m.velocs_prior_log_scale)
    yc, yv, lp = m.sample_from_latents(at.cuda(), x.cuda(), xv.cuda(), mask.cuda(), zc, zv)
    assert yc.requires_grad and lp.shape == (1, B)
    loss = (yc[0] * G1.cuda()).sum() + (yv[0] * G2.cuda()).sum() + (lp[0] * g3.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    # the same under no_grad (inference kernels): identical samples and densities
    with torch.no_grad():
        yc0, yv0, lp0 = m.sample_from_latents(at.cuda(), x.cuda(), xv.cuda(), mask.cuda(), zc.detach(), zv.detach())
    torch.testing.assert_close(yc.detach(), yc0, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(lp.detach(), lp0, rtol=1e-6, atol=2e-4)
    # oracle, fp64
    leaves = {k: v.double().clone().requires_grad_(not k.endswith(".lengthscales")) for k, v in sd.items()}
    zc64 = eps_c.double() * torch.exp(leaves["coords_prior_log_scale"])
    zv64 = eps_v.double() * torch.exp(leaves["velocs_prior_log_scale"])
    ryc, ryv, rlp = fo.conditional_sample_with_logp(leaves, FULL_O, at, x.double(), xv.double(), mask, 1, zc64, zv64, distance_mode="direct")
    rloss = (ryc[0] * G1.double()).sum() + (ryv[0] * G2.double()).sum() + (rlp[0] * g3.double()).sum()
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-4 * max(1.0, abs(float(rloss.detach())))
    names = [k for k, v in leaves.items() if v.requires_grad]
    ref = dict(zip(names, torch.autograd.grad(rloss, [leaves[k] for k in names], allow_unused=True)))
    total = float(torch.sqrt(sum(g.norm() ** 2 for g in ref.values() if g is not None)))
    errs = []
    for k, r in ref.items():
        r = torch.zeros_like(leaves[k]) if r is None else r
        err = float((grads[k].double() - r).norm())
        scale = max(float(r.norm()), 1e-4 * total)
        errs.append(err / scale)
        assert err <= _tol(k) * scale, (k, err, float(r.norm()))
    assert float(np.median(errs)) < 2.5 * GRAD_MEDIAN_RTOL  # (random upstream gradients: measured median 2.1e-4)
    print("sampling backward: worst", max(errs), "median", float(np.median(errs)))


def test_energy_loss_trains_through_the_sampler():
    """EnergyLoss (losses.py:558-664) on the GPU path: value == the hand-composed E(y)/kT + KE + log p with the fp64 energy
    oracle on the same samples; its gradient == the generic sampling backward fed with -F/kT, v and 1/n_atoms; a gradient
    step lowers the loss by the first-order prediction."""
    from oracle import energy_oracle as eo
    from timewarp_b200 import losses
    from timewarp_b200.energy import PeptidePotentialEnergy
    from timewarp_b200.forcefield import amber_like_system
    from timewarp_b200.peptides import alanine_dipeptide

    pep = alanine_dipeptide()
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    provider = losses.EnergyProvider({"ad": energy}, {"ad": torch.tensor(pep.masses, dtype=torch.float32)})
    B, V = 6, pep.num_atoms
    g = torch.Generator().manual_seed(3)

    class Batch:
        atom_coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, V, 3, generator=g)
        atom_velocs = torch.zeros(B, V, 3)
        atom_types = torch.tensor(pep.atom_types)[None].repeat(B, 1)
        masked_elements = torch.zeros(B, V, dtype=torch.bool)
        adj_list, edge_batch_idx = EMPTY_ADJ, EMPTY_EBI
        names, segments = ["ad"] * B, [0, B]

    import bench
    m, _ = build_model(FULL_O, "bf16x3", 0)
    m.load_state_dict({k: v.cuda() for k, v in bench.bench_state_dict(m, "proposal").items()})  # local moves: finite energies
    m.train()
    spec = losses.EnergyLoss(provider, random_velocs=True, num_samples=1)
    torch.manual_seed(21)
    loss = losses.energy_loss(spec, m, Batch, device="cuda")
    m.zero_grad(set_to_none=True)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == len(list(m.parameters())) and all(torch.isfinite(v).all() for v in grads.values())
    # value: replay the draws (x_velocs, then the two latent draws) and compose the loss by hand with the fp64 energy oracle
    torch.manual_seed(21)
    xv = torch.randn_like(Batch.atom_coords.cuda())
    with torch.no_grad():
        yc, yv, lp = m.conditional_sample_with_logp(atom_types=Batch.atom_types.cuda(), x_coords=Batch.atom_coords.cuda(), x_velocs=xv,
                                                    adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(),
                                                    masked_elements=Batch.masked_elements.cuda(), num_samples=1)
    u = eo.potential_energy(sysd.as_float32(), yc[0].cpu().numpy().astype(np.float64)) / energy.kbT
    ke = 0.5 * (yv[0].double() ** 2).sum((-1, -2)).cpu().numpy()
    want = float(((u + ke + lp[0].double().cpu().numpy()) / V).mean())
    assert abs(float(loss.detach()) - want) < 2e-4 * max(1.0, abs(want)), (float(loss.detach()), want)
    # the composed gradient (sampler backward + forces + prior terms) predicts the change of the loss along itself
    def evaluate():
        torch.manual_seed(21)
        with torch.no_grad():
            return float(losses.energy_loss(spec, m, Batch, device="cuda"))

    _check_first_order_decrease(m, float(loss.detach()), evaluate)


def test_density_backward_input_gradients_match_oracle_autograd():
    """log_likelihood differentiated w.r.t. its INPUTS (conditioning coordinates / velocities through the conditioner inputs,
    the attention scores and the centring; target through the flow input), against the oracle's fp64 autograd on a ragged batch."""
    torch.manual_seed(17)
    B, V = 4, 26
    lengths = [26, 19, 26, 8]
    mask = torch.zeros(B, V, dtype=torch.bool)
    for b, n in enumerate(lengths):
        mask[b, n:] = True
    keep = (~mask)[:, :, None]
    x = 0.3 * torch.randn(B, V, 3) * keep
    y = (x + 0.02 * torch.randn(B, V, 3)) * keep
    xv, yv = torch.randn(B, V, 3) * keep, torch.randn(B, V, 3) * keep
    at = torch.randint(0, 5, (B, V)) * (~mask)
    w = torch.randn(B)
    m, sd = build_model(FULL_O, "bf16x3", 6)
    m.train()
    leaves = [t.clone().cuda().requires_grad_(True) for t in (x, xv, y, yv)]
    ll = m.log_likelihood(atom_types=at.cuda(), x_coords=leaves[0], x_velocs=leaves[1], y_coords=leaves[2], y_velocs=leaves[3],
                          adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=mask.cuda())
    (ll * w.cuda()).sum().backward()
    ref_leaves = [t.double().clone().requires_grad_(True) for t in (x, xv, y, yv)]
    rll = fo.log_likelihood(fo.to_dtype(sd, torch.float64), FULL_O, at, ref_leaves[0], ref_leaves[1], ref_leaves[2], ref_leaves[3], mask,
                            distance_mode="direct_sq")
    ref = torch.autograd.grad((rll * w.double()).sum(), ref_leaves)
    torch.testing.assert_close(ll.detach().cpu().double(), rll.detach(), rtol=1e-5, atol=1e-3)
    for name, got, want in zip(("x_coords", "x_velocs", "y_coords", "y_velocs"), leaves, ref):
        rows = keep.expand_as(want)
        err = float((got.grad.cpu().double() - want)[rows].norm() / want[rows].norm())
        assert err < GRAD_RTOL, (name, err)
        assert float(got.grad.cpu()[~rows].abs().max()) < 1e-4 * float(want.abs().max()), name  # padding atoms: no gradient


def test_acceptance_loss_matches_hand_composition():
    """AcceptanceLoss (losses.py:358-555): value against the hand-composed (E(y) - E(x))/kT + log p(y|x) - log p(x|y) with the
    fp64 energy oracle and inference-path densities on the same draws; gradients finite on every parameter; a gradient step
    lowers it by the first-order prediction.  clamp / beta / high-energy filter variants on the same draws."""
    from oracle import energy_oracle as eo
    from timewarp_b200 import losses
    from timewarp_b200.energy import PeptidePotentialEnergy
    from timewarp_b200.forcefield import amber_like_system
    from timewarp_b200.peptides import alanine_dipeptide
    import bench

    pep = alanine_dipeptide()
    sysd = amber_like_system(pep)
    energy = PeptidePotentialEnergy(sysd)
    provider = losses.EnergyProvider({"ad": energy}, {"ad": torch.tensor(pep.masses, dtype=torch.float32)})
    B, V = 6, pep.num_atoms
    g = torch.Generator().manual_seed(4)

    class Batch:
        atom_coords = torch.tensor(pep.coords_nm, dtype=torch.float32)[None] + 0.005 * torch.randn(B, V, 3, generator=g)
        atom_velocs = torch.zeros(B, V, 3)
        atom_types = torch.tensor(pep.atom_types)[None].repeat(B, 1)
        masked_elements = torch.zeros(B, V, dtype=torch.bool)
        adj_list, edge_batch_idx = EMPTY_ADJ, EMPTY_EBI
        names, segments = ["ad"] * B, [0, B]

    m, _ = build_model(FULL_O, "bf16x3", 0)
    m.load_state_dict({k: v.cuda() for k, v in bench.bench_state_dict(m, "proposal").items()})
    m.train()
    spec = losses.AcceptanceLoss(provider, random_velocs=True, num_samples=1)
    torch.manual_seed(33)
    loss = losses.acceptance_loss(spec, m, Batch, device="cuda")
    m.zero_grad(set_to_none=True)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # hand composition on the same draws (inference kernels + fp64 energy oracle)
    torch.manual_seed(33)
    xc = Batch.atom_coords.cuda()
    xv = torch.randn_like(xc)
    kw = dict(atom_types=Batch.atom_types.cuda(), adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda(), masked_elements=Batch.masked_elements.cuda())
    with torch.no_grad():
        yc, yv, lp = m.conditional_sample_with_logp(x_coords=xc, x_velocs=xv, num_samples=1, **kw)
        p_yx = m.log_likelihood(x_coords=yc[0], x_velocs=yv[0], y_coords=xc, y_velocs=xv, **kw)
    s32 = sysd.as_float32()
    de = (eo.potential_energy(s32, yc[0].cpu().numpy().astype(np.float64)) - eo.potential_energy(s32, xc.cpu().numpy().astype(np.float64))) / energy.kbT
    dk = (0.5 * (yv[0].double() ** 2).sum((-1, -2)) - 0.5 * (xv.double() ** 2).sum((-1, -2))).cpu().numpy()
    nla = de + dk + (lp[0] - p_yx).double().cpu().numpy()
    want = float((nla / V).mean())
    assert abs(float(loss.detach()) - want) < 5e-4 * max(1.0, abs(want)), (float(loss.detach()), want)
    # variants on the same draws
    for kwargs, expect in ((dict(clamp=True), float((np.minimum(nla, 0.0) / V).mean())),
                           (dict(beta=0.5), float(((nla + 0.5 * lp[0].double().cpu().numpy()) / V).mean()))):
        torch.manual_seed(33)
        with torch.no_grad():
            v = float(losses.acceptance_loss(losses.AcceptanceLoss(provider, random_velocs=True, **kwargs), m, Batch, device="cuda"))
        assert abs(v - expect) < 5e-4 * max(1.0, abs(expect)), (kwargs, v, expect)
    torch.manual_seed(33)
    with torch.no_grad():  # every proposal flagged: the reference's constant 10000 (losses.py:535-537)
        spec_bad = losses.AcceptanceLoss(provider, high_energy_threshold=300.0, chirality_checker=lambda b, y, mk: torch.ones(len(y), dtype=torch.bool, device=y.device))
        assert float(losses.acceptance_loss(spec_bad, m, Batch, device="cuda")) == 10000.0
    with pytest.raises(ValueError):
        losses.AcceptanceLoss(provider, high_energy_threshold=300.0)
    # the composed gradient (sampler backward -> density backward w.r.t. its conditioning -> forces) predicts the change of the loss
    def evaluate():
        torch.manual_seed(33)
        with torch.no_grad():
            return float(losses.acceptance_loss(spec, m, Batch, device="cuda"))

    _check_first_order_decrease(m, float(loss.detach()), evaluate)


def test_learnable_lengthscales_with_both_passes_in_one_graph():
    """learnable_kernel with a sampling pass AND a density pass in one autograd graph (the AcceptanceLoss shape): the sampling
    pass reads the LAST coupling layer's log_lengthscales, the density pass the FIRST one's, both through one shared buffer --
    each backward must see its own pass's values.  Gradients of both against the oracle's fp64 autograd."""
    torch.manual_seed(23)
    B, V = 4, 24
    x, xv = 0.25 * torch.randn(B, V, 3), torch.randn(B, V, 3)
    at = torch.randint(0, 5, (B, V))
    mask = torch.zeros(B, V, dtype=torch.bool)
    eps_c, eps_v = torch.randn(1, B, V, 3), torch.randn(1, B, V, 3)
    G1, g3 = torch.randn(B, V, 3), torch.randn(B) / V
    m, sd = build_model(FULL_L, "bf16x3", 9)
    m.train()
    m.zero_grad(set_to_none=True)
    kw = dict(adj_list=EMPTY_ADJ.cuda(), edge_batch_idx=EMPTY_EBI.cuda())
    zc = eps_c.cuda() * torch.exp(m.coords_prior_log_scale)
    zv = eps_v.cuda() * torch.exp(m.velocs_prior_log_scale)
    yc, yv, lp = m.sample_from_latents(at.cuda(), x.cuda(), xv.cuda(), mask.cuda(), zc, zv)
    ll = m.log_likelihood(atom_types=at.cuda(), x_coords=yc[0], x_velocs=yv[0], y_coords=x.cuda(), y_velocs=xv.cuda(), masked_elements=mask.cuda(), **kw)
    loss = (yc[0] * G1.cuda()).sum() + ((lp[0] - ll) * g3.cuda()).sum()
    loss.backward()
    first = "flow.chain.0.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"
    last = "flow.chain.7.scale_transformer.encoder_layers.0.self_attn.attention.log_lengthscales"
    got = {k: p.grad.detach().cpu().double() for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(k for k in got if k.endswith("log_lengthscales")) == sorted([first, last])
    leaves = {k: v.double().clone().requires_grad_(not k.endswith(".lengthscales")) for k, v in sd.items()}
    zc64 = eps_c.double() * torch.exp(leaves["coords_prior_log_scale"])
    zv64 = eps_v.double() * torch.exp(leaves["velocs_prior_log_scale"])
    ryc, ryv, rlp = fo.conditional_sample_with_logp(leaves, FULL_L, at, x.double(), xv.double(), mask, 1, zc64, zv64, distance_mode="direct_sq")
    rll = fo.log_likelihood(leaves, FULL_L, at, ryc[0], ryv[0], x.double(), xv.double(), mask, distance_mode="direct_sq")
    rloss = (ryc[0] * G1.double()).sum() + ((rlp[0] - rll) * g3.double()).sum()
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-4 * max(1.0, abs(float(rloss.detach())))
    keys = [first, last, "flow.chain.3.shift_transformer.encoder_layers.1.self_attn.values_proj.weight", "coords_prior_log_scale"]
    ref = dict(zip(keys, torch.autograd.grad(rloss, [leaves[k] for k in keys])))
    assert float((ref[first] - ref[last]).abs().max()) > 1e-6  # the two lengthscale gradients differ: a mix-up would show
    for k in keys:
        err = float((got[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-12))
        assert err < 2 * GRAD_RTOL, (k, err, got[k], ref[k])

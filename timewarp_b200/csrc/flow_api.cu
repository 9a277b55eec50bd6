// C-ABI entry points of the flow (include/timewarp_b200.h) and the per-pass orchestration.
// One call = one flow pass: every kernel is enqueued on the caller's stream, nothing syncs.
#include "flow_simt.cuh"
#include "flow_tc.cuh"

namespace tw {

struct FlowBuffers {
  float *xc, *com, *scores, *feat;
  float *actA[2], *actB[2], *hidA[2], *hidB[2], *vals[2], *att[2], *ffn[2], *st[2];
  float *zc, *zv, *delta;
  TcScratch tc;
};

static int max_hidden(const tw_flow_config* c) {
  int m = 1;
  for (int i = 0; i < c->num_mlp_hidden; i++) m = c->mlp_hidden_dims[i] > m ? c->mlp_hidden_dims[i] : m;
  return m;
}

static int validate_cfg(const tw_flow_config* c) {
  TW_CHECK_ARG(c != nullptr, "cfg is NULL");
  TW_CHECK_ARG(c->atom_embedding_dim >= 1 && c->atom_embedding_dim <= 4096, "bad atom_embedding_dim");
  TW_CHECK_ARG(c->num_mlp_hidden >= 0 && c->num_mlp_hidden <= TW_MAX_MLP_HIDDEN, "num_mlp_hidden out of range");
  TW_CHECK_ARG(c->num_coupling_layers >= 2 && c->num_coupling_layers % 2 == 0,
               "Real NVP should have an even number of coupling layers");  // model_constructor.py:156-158
  TW_CHECK_ARG(c->position_layer_index_mod_2 == 0 || c->position_layer_index_mod_2 == 1,
               "positions_layer_index can only be 0 or 1");  // model_constructor.py:160-163
  TW_CHECK_ARG(c->num_transformer_layers >= 1, "num_transformer_layers must be >= 1");
  TW_CHECK_ARG(c->d_model >= 1 && c->dim_feedforward >= 1, "bad d_model / dim_feedforward");
  TW_CHECK_ARG(c->num_heads >= 1 && c->num_heads <= TW_MAX_HEADS, "num_heads out of range");
  TW_CHECK_ARG(c->num_atom_types >= 1, "bad num_atom_types");
  TW_CHECK_ARG(c->precision >= TW_PRECISION_FP32 && c->precision <= TW_PRECISION_BF16, "unknown precision");
  TW_CHECK_ARG(c->attention_type == TW_ATTENTION_KERNEL || c->attention_type == TW_ATTENTION_CHEBYSHEV ||
                   c->attention_type == TW_ATTENTION_LOCAL, "unknown attention_type");
  TW_CHECK_ARG(c->attention_type != TW_ATTENTION_LOCAL || c->max_radius > 0.f, "local attention needs max_radius > 0");
  TW_CHECK_ARG(c->attention_type != TW_ATTENTION_CHEBYSHEV || (c->cheb_order >= 1 && c->cheb_order <= TW_MAX_CHEB_ORDER),
               "cheb_order out of range (1..32)");
  if (c->precision != TW_PRECISION_FP32 && !tc_supported(c))
    return fail(TW_ERR_UNSUPPORTED, "tensor-core path needs d_model=128, dim_feedforward%%128==0, one MLP hidden layer of 256");
  return TW_OK;
}

// Carve the workspace.  With base == NULL only computes the size.
static size_t carve(const tw_flow_config* c, int64_t n, int64_t n_cond, int64_t V, void* base, size_t cap, FlowBuffers* fb) {
  Arena ar(base, cap);
  const int64_t M = n * V;
  const int D = c->d_model, H = c->num_heads, F = c->dim_feedforward, E = c->atom_embedding_dim;
  FlowBuffers b;
  b.xc = ar.take<float>(n_cond * V * 3);
  b.com = ar.take<float>(n_cond * 3);
  b.scores = ar.take<float>(n_cond * H * V * V);
  b.zc = ar.take<float>(M * 3);
  b.zv = ar.take<float>(M * 3);
  b.delta = ar.take<float>(n);
  for (int i = 0; i < 2; i++) b.st[i] = ar.take<float>(M * 3);
  for (int i = 0; i < 2; i++) b.actA[i] = ar.take<float>(M * D);
  for (int i = 0; i < 2; i++) b.actB[i] = ar.take<float>(M * D);
  const int hid = max_hidden(c);
  b.feat = ar.take<float>(M * (E + 9));
  for (int i = 0; i < 2; i++) {
    b.hidA[i] = ar.take<float>(M * hid);
    b.hidB[i] = ar.take<float>(M * hid);
    b.vals[i] = ar.take<float>(M * H * D * (c->attention_type == TW_ATTENTION_LOCAL ? 3 : 1));  // local: q | k | v per head
    b.att[i] = ar.take<float>(M * H * D);
    // the fused tensor-core FFN keeps the hidden activation on chip
    b.ffn[i] = (c->precision == TW_PRECISION_FP32 || !(tc_stage_mask() & TC_FFN)) ? ar.take<float>(M * F) : nullptr;
  }
  b.tc = TcScratch{};
  if (c->precision != TW_PRECISION_FP32) tc_carve(c, n, n_cond, V, ar, &b.tc);
  if (fb) *fb = b;
  return align_up(ar.off, 256);
}

struct PassCtx {
  const tw_flow_config* c;
  ParamView pv;
  FlowBuffers fb;
  const int64_t* atom_types;
  const float* x_velocs;
  const uint8_t* mask;
  int64_t n, n_cond;
  int V;
  cudaStream_t st;
  const uint8_t* packed = nullptr;  // packed tensor-core weight images (precision != fp32)
};

// One transformer block pair (scale net, shift net) of coupling layer k on CUDA cores:
// custom_transformer_block.py:46-82, custom_attention_encoder.py:82-114.
static int conditioner(PassCtx& p, int k) {
  const tw_flow_config* c = p.c;
  FlowBuffers& b = p.fb;
  const int64_t M = p.n * p.V;
  const int D = c->d_model, H = c->num_heads, F = c->dim_feedforward, E = c->atom_embedding_dim, nh = c->num_mlp_hidden;
  const bool pos = (k % 2) == c->position_layer_index_mod_2;
  uint32_t tcs = (c->precision == TW_PRECISION_FP32) ? 0u : tc_stage_mask();
  // dot-product attention, and samples of more than 128 atoms (the tensor-core attention kernels hold one sample's scores on
  // chip), run their attention on the CUDA-core kernels; the MLPs and the FFN stay on tcgen05
  if (p.pv.local() || p.V > 128) tcs &= ~(uint32_t)(TC_MIX | TC_ATTN_PROJ);
  TcScratch tcx = b.tc;
  tcx.packed = p.packed;
  const float* cur[2] = {b.feat, b.feat};
  int cur_dim = E + 9;
  float* hid[2][2] = {{b.hidA[0], b.hidA[1]}, {b.hidB[0], b.hidB[1]}};
  if (tcs & TC_IN_MLP) {
    TW_TRY(tc_in_mlp(c, p.pv, k, tcx, p.atom_types, b.xc, p.x_velocs, pos ? b.zv : b.zc, b.actA, p.n, p.n_cond, p.V, p.st));
  } else {
    TW_TRY(launch_features(p.pv.embed(), p.atom_types, b.xc, p.x_velocs, pos ? b.zv : b.zc, p.n, p.n_cond, p.V, E,
                           c->num_atom_types, b.feat, p.st));
    for (int i = 0; i < nh; i++) {
      Lin2 a{};
      for (int s = 0; s < 2; s++) a.X[s] = cur[s], a.W[s] = p.pv.in_w(k, s, i), a.b[s] = p.pv.in_b(k, s, i), a.Y[s] = hid[i & 1][s];
      TW_TRY(launch_linear(a, 2, M, c->mlp_hidden_dims[i], cur_dim, cur_dim, 0, c->mlp_hidden_dims[i], ACT_SILU, p.st));
      cur[0] = hid[i & 1][0], cur[1] = hid[i & 1][1], cur_dim = c->mlp_hidden_dims[i];
    }
    Lin2 a{};
    for (int s = 0; s < 2; s++) a.X[s] = cur[s], a.W[s] = p.pv.in_w(k, s, nh), a.b[s] = p.pv.in_b(k, s, nh), a.Y[s] = b.actA[s];
    TW_TRY(launch_linear(a, 2, M, D, cur_dim, cur_dim, 0, D, ACT_NONE, p.st));
  }
  const bool cheb = p.pv.chebyshev();
  const float* ls = p.pv.enc(0, 0, 0, 1);
  for (int t = 0; t < c->num_transformer_layers; t++) {
    // chebyshev_kernel: every attention layer of every network has its own basis function, hence its own scores
    // (the reference's cache key contains the per-module basis lambda, kernel_attention.py:333-335): the scores are
    // recomputed here and the two networks run one after the other.  Otherwise one score set serves the whole pass.
    if (p.pv.local()) {  // qkv projection -> masked softmax attention within max_radius -> output projection + residual -> LN1
      Lin2 a{};
      for (int s = 0; s < 2; s++) a.X[s] = b.actA[s], a.W[s] = p.pv.enc(k, s, t, 0), a.Y[s] = b.vals[s];
      TW_TRY(launch_linear(a, 2, M, H * 3 * D, D, D, 0, H * 3 * D, ACT_NONE, p.st));
      TW_TRY(launch_local_attn(b.vals[0], b.vals[1], b.att[0], b.att[1], 2, p.n, p.n_cond, p.V, H, D, b.xc, p.mask, c->max_radius, p.st));
      Lin2 o{};
      for (int s = 0; s < 2; s++) o.X[s] = b.att[s], o.W[s] = p.pv.enc(k, s, t, 2), o.R[s] = b.actA[s], o.Y[s] = b.actB[s];
      TW_TRY(launch_linear(o, 2, M, D, H * D, H * D, D, D, ACT_NONE, p.st));
      TW_TRY(launch_layernorm(b.actB[0], b.actB[1], p.pv.enc(k, 0, t, 7), p.pv.enc(k, 1, t, 7), p.pv.enc(k, 0, t, 8),
                              p.pv.enc(k, 1, t, 8), 2, M, D, c->layer_norm_eps, p.st));
    }
    for (int pass = 0; pass < (p.pv.local() ? 0 : (cheb ? 2 : 1)); pass++) {
      const int only = cheb ? pass : -1;
      if ((tcs & TC_MIX) && (tcs & TC_ATTN_PROJ)) {
        if (cheb) TW_TRY(tc_begin_pass_direct(c, tcx, b.xc, p.mask, ls, p.n_cond, p.V, p.st, p.pv.cheb(k, only, t), false));
        TW_TRY(tc_attention_layer(c, p.pv, k, t, tcx, b.actA, b.actB, p.n, p.n_cond, p.V, p.st, nullptr, only));
      } else {
        if (cheb)
          TW_TRY(launch_scores(b.xc, p.mask, ls, p.n_cond, p.V, H, b.scores, p.st, p.pv.cheb(k, only, t), c->cheb_order,
                               c->force_asymptotic_zero));
        const int s0 = cheb ? only : 0, s1 = cheb ? only : 1, nets = cheb ? 1 : 2;
        Lin2 a{};
        a.X[0] = b.actA[s0], a.W[0] = p.pv.enc(k, s0, t, 0), a.Y[0] = b.vals[s0];
        a.X[1] = b.actA[s1], a.W[1] = p.pv.enc(k, s1, t, 0), a.Y[1] = b.vals[s1];
        TW_TRY(launch_linear(a, nets, M, H * D, D, D, 0, H * D, ACT_NONE, p.st));
        TW_TRY(launch_attn_mix(b.scores, b.vals[s0], b.vals[s1], b.att[s0], b.att[s1], nets, p.n, p.n_cond, p.V, H, D, p.st));
        Lin2 o{};
        o.X[0] = b.att[s0], o.W[0] = p.pv.enc(k, s0, t, 2), o.R[0] = b.actA[s0], o.Y[0] = b.actB[s0];
        o.X[1] = b.att[s1], o.W[1] = p.pv.enc(k, s1, t, 2), o.R[1] = b.actA[s1], o.Y[1] = b.actB[s1];
        TW_TRY(launch_linear(o, nets, M, D, H * D, H * D, D, D, ACT_NONE, p.st));
        TW_TRY(launch_layernorm(b.actB[s0], b.actB[s1], p.pv.enc(k, s0, t, 7), p.pv.enc(k, s1, t, 7), p.pv.enc(k, s0, t, 8),
                                p.pv.enc(k, s1, t, 8), nets, M, D, c->layer_norm_eps, p.st));
      }
    }
    if (tcs & TC_FFN) {
      TW_TRY(tc_ffn_layer(c, p.pv, k, t, tcx, b.actB, b.actA, M, p.st));  // fused linear1+ReLU+linear2+residual+LN2
    } else {
      {
        ProfScope prof_ffn(PROF_FFN, p.st);
        Lin2 f1{};
        for (int s = 0; s < 2; s++) f1.X[s] = b.actB[s], f1.W[s] = p.pv.enc(k, s, t, 3), f1.b[s] = p.pv.enc(k, s, t, 4), f1.Y[s] = b.ffn[s];
        TW_TRY(launch_linear(f1, 2, M, F, D, D, 0, F, ACT_RELU, p.st));
        Lin2 f2{};
        for (int s = 0; s < 2; s++)
          f2.X[s] = b.ffn[s], f2.W[s] = p.pv.enc(k, s, t, 5), f2.b[s] = p.pv.enc(k, s, t, 6), f2.R[s] = b.actB[s], f2.Y[s] = b.actA[s];
        TW_TRY(launch_linear(f2, 2, M, D, F, F, D, D, ACT_NONE, p.st));
      }
      TW_TRY(launch_layernorm(b.actA[0], b.actA[1], p.pv.enc(k, 0, t, 9), p.pv.enc(k, 1, t, 9), p.pv.enc(k, 0, t, 10),
                              p.pv.enc(k, 1, t, 10), 2, M, D, c->layer_norm_eps, p.st));
    }
  }
  // out_mlp
  if (tcs & TC_OUT_MLP) {
    TW_TRY(tc_out_mlp(c, p.pv, k, tcx, b.actA, b.st, M, p.st));
    return TW_OK;
  }
  cur[0] = b.actA[0], cur[1] = b.actA[1], cur_dim = D;
  for (int i = 0; i < nh; i++) {
    Lin2 a{};
    for (int s = 0; s < 2; s++) a.X[s] = cur[s], a.W[s] = p.pv.out_w(k, s, i), a.b[s] = p.pv.out_b(k, s, i), a.Y[s] = hid[i & 1][s];
    TW_TRY(launch_linear(a, 2, M, c->mlp_hidden_dims[i], cur_dim, cur_dim, 0, c->mlp_hidden_dims[i], ACT_SILU, p.st));
    cur[0] = hid[i & 1][0], cur[1] = hid[i & 1][1], cur_dim = c->mlp_hidden_dims[i];
  }
  {
    Lin2 a{};
    for (int s = 0; s < 2; s++) a.X[s] = cur[s], a.W[s] = p.pv.out_w(k, s, nh), a.b[s] = p.pv.out_b(k, s, nh), a.Y[s] = b.st[s];
    TW_TRY(launch_linear(a, 2, M, 3, cur_dim, cur_dim, 0, 3, ACT_NONE, p.st));
  }
  return TW_OK;
}

// Prepare a pass: centre the conditioning coordinates, attention scores once per pass
// (the reference's Cache: model_constructor.py:189-196; 1 miss + 47 hits).
static int begin_pass(PassCtx& p, const float* x_coords) {
  TW_TRY(launch_prep(x_coords, p.mask, p.n_cond, p.V, p.fb.xc, p.fb.com, p.st));
  // lengthscales of chain[0].scale_transformer.encoder_layers[0] (cache key maps lengthscales -> 0)
  const float* ls = p.pv.enc(0, 0, 0, 1);
  if (p.pv.chebyshev() || p.pv.local()) {  // scores are per attention layer (conditioner()) / not position-only at all
    if (p.c->precision != TW_PRECISION_FP32) TW_TRY(tc_begin_pass(p.c, p.pv, p.fb.tc, nullptr, p.mask, p.n, p.n_cond, p.V, p.st));
    return TW_OK;
  }
  if (p.c->precision != TW_PRECISION_FP32 && tc_supported(p.c) && tc_scores_direct_supported(p.V))
    return tc_begin_pass_direct(p.c, p.fb.tc, p.fb.xc, p.mask, ls, p.n_cond, p.V, p.st);  // no fp32 score tensor
  TW_TRY(launch_scores(p.fb.xc, p.mask, ls, p.n_cond, p.V, p.c->num_heads, p.fb.scores, p.st));
  if (p.c->precision != TW_PRECISION_FP32)  // (V > 128: fp32 scores only, no operand images)
    TW_TRY(tc_begin_pass(p.c, p.pv, p.fb.tc, p.V > 128 ? nullptr : p.fb.scores, p.mask, p.n, p.n_cond, p.V, p.st));
  return TW_OK;
}

static int run_layers(PassCtx& p, bool reverse) {
  const int L = p.c->num_coupling_layers;
  for (int step = 0; step < L; step++) {
    const int k = reverse ? L - 1 - step : step;
    TW_TRY(conditioner(p, k));
    const bool pos = (k % 2) == p.c->position_layer_index_mod_2;
    TW_TRY(launch_coupling(p.fb.st[0], p.fb.st[1], pos ? p.fb.zc : p.fb.zv, p.mask, p.fb.delta, p.n, p.n_cond, p.V,
                           reverse ? 1 : 0, nullptr, nullptr, p.st));
  }
  return TW_OK;
}

static int set_packed(PassCtx& p, const void* packed) {
  if (p.c->precision == TW_PRECISION_FP32) return TW_OK;
  TW_CHECK_ARG(packed != nullptr, "packed_weights is NULL: call tw_flow_pack_weights first (tensor-core precisions)");
  TW_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "packed_weights must be 1024-byte aligned");
  p.packed = (const uint8_t*)packed;
  return TW_OK;
}

static int check_common(const tw_flow_config* cfg, const void* const* params, int64_t n, int64_t n_cond, int64_t V) {
  TW_TRY(validate_cfg(cfg));
  TW_CHECK_ARG(params != nullptr, "params is NULL");
  TW_CHECK_ARG(n >= 0 && n_cond >= 0 && V >= 1, "bad sizes");
  TW_CHECK_ARG(V <= 1024, "n_atoms > 1024 not supported");
  TW_CHECK_ARG(n_cond == 0 ? n == 0 : n % n_cond == 0, "n_samples must be a multiple of n_cond");
  TW_CHECK_ARG(n * V < (1LL << 31), "too many tokens for one call");
  return TW_OK;
}

}  // namespace tw

using namespace tw;

extern "C" {

int tw_abi_version(void) { return TW_ABI_VERSION; }
const char* tw_last_error(void) { return err_buf(); }

int tw_flow_num_params(const tw_flow_config* cfg) {
  if (validate_cfg(cfg) != TW_OK) return -1;
  ParamView pv{cfg, nullptr};
  return pv.total();
}

int tw_flow_workspace_bytes(const tw_flow_config* cfg, int64_t n_samples, int64_t n_cond, int64_t n_atoms, size_t* bytes) {
  TW_TRY(validate_cfg(cfg));
  TW_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  TW_CHECK_ARG(n_samples >= 0 && n_cond >= 0 && n_atoms >= 1, "bad sizes");
  *bytes = carve(cfg, n_samples, n_cond, n_atoms, nullptr, 0, nullptr) + 256;
  return TW_OK;
}

int tw_flow_packed_bytes(const tw_flow_config* cfg, size_t* bytes) {
  TW_TRY(validate_cfg(cfg));
  TW_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  *bytes = (cfg->precision == TW_PRECISION_FP32) ? 0 : tc_packed_bytes(cfg);
  return TW_OK;
}

int tw_flow_pack_weights(const tw_flow_config* cfg, const void* const* params, void* packed, size_t packed_bytes, void* stream) {
  TW_TRY(validate_cfg(cfg));
  TW_CHECK_ARG(params != nullptr, "params is NULL");
  if (cfg->precision == TW_PRECISION_FP32) return TW_OK;
  return tc_pack_weights(cfg, ParamView{cfg, params}, (uint8_t*)packed, packed_bytes, (cudaStream_t)stream);
}

int tw_attn_scores(const float* coords, const uint8_t* mask, const float* lengthscales, int64_t B, int64_t V, int32_t H,
                   float* out, void* stream) {
  TW_CHECK_ARG(B >= 0 && V >= 1 && H >= 1, "bad sizes");
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(coords && mask && lengthscales && out, "NULL pointer");
  return launch_scores(coords, mask, lengthscales, B, (int)V, H, out, (cudaStream_t)stream);
}

int tw_flow_scale_shift(const tw_flow_config* cfg, const void* const* params, int32_t layer_idx, const int64_t* atom_types,
                        const float* x_coords_centred, const float* x_velocs, const float* z_coords, const float* z_velocs,
                        const uint8_t* mask, int64_t B, int64_t V, float* out_scale, float* out_shift, const void* packed_weights,
                        void* workspace, size_t workspace_bytes, void* stream) {
  TW_TRY(check_common(cfg, params, B, B, V));
  TW_CHECK_ARG(layer_idx >= 0 && layer_idx < cfg->num_coupling_layers, "layer_idx out of range");
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(atom_types && x_coords_centred && x_velocs && z_coords && z_velocs && mask && out_scale && out_shift, "NULL pointer");
  PassCtx p{cfg, ParamView{cfg, params}, {}, atom_types, x_velocs, mask, B, B, (int)V, (cudaStream_t)stream};
  TW_TRY(set_packed(p, packed_weights));
  size_t need = carve(cfg, B, B, V, workspace, workspace_bytes, &p.fb);
  if (need > workspace_bytes || !workspace) return fail(TW_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, need);
  const size_t zb = (size_t)B * V * 3 * sizeof(float);
  TW_CUDA(cudaMemcpyAsync(p.fb.xc, x_coords_centred, zb, cudaMemcpyDeviceToDevice, p.st));
  TW_CUDA(cudaMemcpyAsync(p.fb.zc, z_coords, zb, cudaMemcpyDeviceToDevice, p.st));
  TW_CUDA(cudaMemcpyAsync(p.fb.zv, z_velocs, zb, cudaMemcpyDeviceToDevice, p.st));
  if (p.pv.chebyshev() || p.pv.local()) {
    if (cfg->precision != TW_PRECISION_FP32) TW_TRY(tc_begin_pass(cfg, p.pv, p.fb.tc, nullptr, mask, B, B, (int)V, p.st));
  } else {
    TW_TRY(launch_scores(p.fb.xc, mask, p.pv.enc(0, 0, 0, 1), B, (int)V, cfg->num_heads, p.fb.scores, p.st));
    if (cfg->precision != TW_PRECISION_FP32) TW_TRY(tc_begin_pass(cfg, p.pv, p.fb.tc, p.fb.scores, mask, B, B, (int)V, p.st));
  }
  TW_TRY(conditioner(p, layer_idx));
  return launch_coupling(p.fb.st[0], p.fb.st[1], nullptr, mask, nullptr, B, B, (int)V, 0, out_scale, out_shift, p.st);
}

int tw_flow_log_likelihood(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types,
                           const float* x_coords, const float* x_velocs, const float* y_coords, const float* y_velocs,
                           const uint8_t* mask, int64_t B, int64_t V, int32_t flags, float* out_log_prob,
                           float* out_z_coords, float* out_z_velocs, const void* packed_weights, void* workspace,
                           size_t workspace_bytes, void* stream) {
  TW_TRY(check_common(cfg, params, B, B, V));
  if (B == 0) return TW_OK;  // empty batch: nothing to do (empty tensors have NULL data pointers)
  TW_CHECK_ARG(atom_types && x_coords && x_velocs && y_coords && y_velocs && mask && out_log_prob, "NULL pointer");
  PassCtx p{cfg, ParamView{cfg, params}, {}, atom_types, x_velocs, mask, B, B, (int)V, (cudaStream_t)stream};
  TW_TRY(set_packed(p, packed_weights));
  size_t need = carve(cfg, B, B, V, workspace, workspace_bytes, &p.fb);
  if (need > workspace_bytes || !workspace) return fail(TW_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, need);
  const int64_t cnt = B * V * 3;
  TW_TRY(begin_pass(p, x_coords));
  if (flags & TW_FLOW_DISPLACEMENT_TARGET)
    TW_TRY(launch_sub(y_coords, x_coords, cnt, p.fb.zc, p.st));  // residual target w.r.t. the un-centred x (flow.py:148-149)
  else
    TW_CUDA(cudaMemcpyAsync(p.fb.zc, y_coords, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));  // flow.py:151
  TW_CUDA(cudaMemcpyAsync(p.fb.zv, y_velocs, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  TW_CUDA(cudaMemsetAsync(p.fb.delta, 0, B * sizeof(float), p.st));
  TW_TRY(run_layers(p, false));
  TW_TRY(launch_prior(p.fb.zc, p.fb.zv, mask, p.pv.log_scale_c(), p.pv.log_scale_v(), p.fb.delta, -1.f, B, B, (int)V,
                      out_log_prob, p.st));
  if (out_z_coords) TW_CUDA(cudaMemcpyAsync(out_z_coords, p.fb.zc, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  if (out_z_velocs) TW_CUDA(cudaMemcpyAsync(out_z_velocs, p.fb.zv, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  return TW_OK;
}

int tw_flow_sample(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types, const float* x_coords,
                   const float* x_velocs, const uint8_t* mask, int64_t n_cond, int64_t V, int64_t S, int32_t flags,
                   const float* z_coords, const float* z_velocs, float* out_y_coords, float* out_y_velocs, float* out_log_prob,
                   const void* packed_weights, void* workspace, size_t workspace_bytes, void* stream) {
  TW_CHECK_ARG(S >= 0, "bad num_samples");
  const int64_t n = S * n_cond;
  TW_TRY(check_common(cfg, params, n, n_cond, V));
  if (n == 0) return TW_OK;
  TW_CHECK_ARG(atom_types && x_coords && x_velocs && mask && z_coords && z_velocs && out_y_coords && out_y_velocs, "NULL pointer");
  PassCtx p{cfg, ParamView{cfg, params}, {}, atom_types, x_velocs, mask, n, n_cond, (int)V, (cudaStream_t)stream};
  TW_TRY(set_packed(p, packed_weights));
  size_t need = carve(cfg, n, n_cond, V, workspace, workspace_bytes, &p.fb);
  if (need > workspace_bytes || !workspace) return fail(TW_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, need);
  const int64_t cnt = n * V * 3;
  TW_TRY(begin_pass(p, x_coords));
  TW_CUDA(cudaMemcpyAsync(p.fb.zc, z_coords, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  TW_CUDA(cudaMemcpyAsync(p.fb.zv, z_velocs, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  TW_CUDA(cudaMemsetAsync(p.fb.delta, 0, n * sizeof(float), p.st));
  TW_TRY(run_layers(p, true));
  if (flags & TW_FLOW_DISPLACEMENT_TARGET)
    TW_TRY(launch_uncentre(p.fb.xc, p.fb.com, p.fb.zc, n, n_cond, (int)V, out_y_coords, p.st));
  else
    TW_CUDA(cudaMemcpyAsync(out_y_coords, p.fb.zc, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));  // flow.py:310
  TW_CUDA(cudaMemcpyAsync(out_y_velocs, p.fb.zv, cnt * sizeof(float), cudaMemcpyDeviceToDevice, p.st));
  if (out_log_prob)  // log p(y|x) = log p(z) + delta_logp, prior evaluated at the ORIGINAL draws (flow.py:322-334)
    TW_TRY(launch_prior(z_coords, z_velocs, mask, p.pv.log_scale_c(), p.pv.log_scale_v(), p.fb.delta, +1.f, n, n_cond,
                        (int)V, out_log_prob, p.st));
  return TW_OK;
}

int tw_debug_set_ffn_trace(long long* device_buf) {
  tc_set_ffn_trace(device_buf);
  return TW_OK;
}
int tw_debug_set_trace(int kernel_class, long long* device_buf) {
  tc_set_trace(kernel_class, device_buf);
  return TW_OK;
}

}  // extern "C"

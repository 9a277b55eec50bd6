// Tensor-core (tcgen05 / TMEM / bulk-TMA) path of the conditioner networks for the flagship
// layer sizes: d_model 128, one MLP hidden layer of 256, dim_feedforward % 128 == 0,
// atom_embedding_dim + 9 <= 64, num_heads * 128 = value width.
//
// Weights are re-packed once (tw_flow_pack_weights) into bf16 "operand images": [rows x 64] K-major
// tiles in the 128-byte-swizzled shared-memory layout tcgen05.mma reads, hi part (bf16(w)) and lo
// part (bf16(w - hi)), so that a kernel fetches a tile with ONE bulk async copy (no tensor map).
#pragma once
#include "flow_simt.cuh"

namespace tw {

// byte sizes / offsets of the packed image of ONE conditioner network (scale or shift) of ONE coupling layer
struct TcLayout {
  int D, F, H, E, hid, T;
  size_t in_w1, in_w2, out_w1;       // MLP weight images
  size_t enc0, enc_stride;            // first encoder layer, stride between encoder layers
  size_t enc_wc, enc_ffn;             // inside an encoder layer: combined attention projection, FFN chunks
  size_t net_bytes;
  __host__ static TcLayout make(const tw_flow_config* c);
  __host__ size_t net_offset(int k, int net) const { return (size_t)(k * 2 + net) * net_bytes; }
  __host__ size_t total(const tw_flow_config* c) const { return (size_t)c->num_coupling_layers * 2 * net_bytes; }
};

struct TcScratch {
  uint8_t* mixed_img[2];  // A-operand images of the per-head neighbourhood averages: [tile][H*2][hi|lo][16 KB]
  uint8_t* scores_img;    // B-operand images of the attention weights: [n_cond][H][hi|lo][VP*VP*2]
  const uint8_t* packed;  // packed weights (caller-owned, persistent)
  float* ffn_tail;        // split-tile scratch of the pair FFN: [2 nets][kFfnTailTiles][16 parts][2 ranks][128 x 128] fp32 partial sums,
                          // followed by [2 nets][kFfnTailTiles][2 ranks | 2 ranks] arrival and done counters; zeroed once per pass,
                          // self-cleaning
};
constexpr int kFfnTailTiles = 4;
constexpr size_t kFfnTailFloats = (size_t)2 * kFfnTailTiles * 16 * 2 * 128 * 128;
constexpr size_t kFfnTailBytes = kFfnTailFloats * 4 + 2 * kFfnTailTiles * 4 * 4;

enum TcStage : uint32_t { TC_FFN = 1, TC_ATTN_PROJ = 2, TC_MIX = 4, TC_IN_MLP = 8, TC_OUT_MLP = 16, TC_ALL = 31 };
constexpr uint32_t TC_IMPLEMENTED = TC_ALL;  // stages with a tensor-core kernel; the rest run on CUDA cores

bool tc_supported(const tw_flow_config* c);
uint32_t tc_stage_mask();  // TW_TC_STAGES environment override (bring-up), default: every stage that exists
void tc_carve(const tw_flow_config* c, int64_t n, int64_t n_cond, int64_t V, Arena& ar, TcScratch* out);
int tc_pack_weights(const tw_flow_config* c, const ParamView& pv, uint8_t* packed, size_t bytes, cudaStream_t st);
int tc_begin_pass(const tw_flow_config* c, const ParamView& pv, TcScratch& tc, const float* scores, const uint8_t* mask,
                  int64_t n, int64_t n_cond, int V, cudaStream_t st);
bool tc_scores_direct_supported(int V);
int tc_begin_pass_direct(const tw_flow_config* c, TcScratch& tc, const float* xc, const uint8_t* mask, const float* lengthscales,
                         int64_t n_cond, int V, cudaStream_t st, const float* cheb = nullptr, bool clear_tail = true);
size_t tc_packed_bytes(const tw_flow_config* c);
void tc_set_ffn_trace(long long* buf);
void tc_set_trace(int cls, long long* buf);  // 1 fused FFN, 2 mixing kernel  // debug: event trace of the fused FFN (see FfnArgs::trace)
// fused FFN + residual + LayerNorm of encoder layer t for both networks: out = LN2(x + FFN(x))
// out = LN1(x + sum_h W_c,h (A_h x)) for both networks (tensor-core mixing + projection)
int tc_attention_layer(const tw_flow_config* c, const ParamView& pv, int k, int t, const TcScratch& tc, float* const x[2],
                       float* const out[2], int64_t n, int64_t n_cond, int V, cudaStream_t st, float* const* pre = nullptr,
                       int only_net = -1);
// feature-major fused attention layer (attn_fm.cu): project every head first, mix per sample afterwards; any atom count <= 128
bool tc_attn_fm_supported(int V, int64_t n);
// ... with groups of at most 80 tokens and a deeper pipeline (attn_fm3.cu: four projection buffers, two accumulators, two tile sets)
bool tc_attn_fm3_supported(int V, int64_t n);
int tc_attn_fm3(const tw_flow_config* c, const float* const x[2], float* const out[2], const uint8_t* const wc[2],
                const float* const gamma[2], const float* const beta[2], const uint8_t* scores_img, int64_t n, int64_t n_cond, int V,
                int nets, cudaStream_t st);
int tc_attn_fm(const tw_flow_config* c, const float* const x[2], float* const out[2], const uint8_t* const wc[2],
               const float* const gamma[2], const float* const beta[2], const uint8_t* scores_img, int64_t n, int64_t n_cond, int V,
               int nets, cudaStream_t st);
void tc_set_fm_trace(long long* buf);
// the mixing step alone (also used by the backward pass with the transposed score images)
int tc_mix(const tw_flow_config* c, const float* const x[2], uint8_t* const img[2], const uint8_t* scores_img, int64_t n,
           int64_t n_cond, int V, cudaStream_t st, int nets = 2);
int tc_scores_images(const tw_flow_config* c, const float* scores, int64_t n_cond, int V, uint8_t* img, int transpose, cudaStream_t st);
size_t tc_scores_img_bytes(const tw_flow_config* c, int64_t n_cond, int V);
size_t tc_mixed_img_bytes(const tw_flow_config* c, int64_t M);
// in_mlp (feature gather + 2 linears) and out_mlp (2 linears -> s or t [M,3]) of both networks
int tc_in_mlp(const tw_flow_config* c, const ParamView& pv, int k, const TcScratch& tc, const int64_t* atom_types, const float* xc,
              const float* xv, const float* z_other, float* const out[2], int64_t n, int64_t n_cond, int V, cudaStream_t st);
int tc_out_mlp(const tw_flow_config* c, const ParamView& pv, int k, const TcScratch& tc, float* const x[2], float* const out[2],
               int64_t M, cudaStream_t st);
int tc_ffn_layer(const tw_flow_config* c, const ParamView& pv, int k, int t, const TcScratch& tc, float* const x[2],
                 float* const out[2], int64_t M, cudaStream_t st, float* const* pre = nullptr);

}  // namespace tw

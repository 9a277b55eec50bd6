// Generic (any layer size) CUDA-core fp32 kernels of the flow + small fused kernels shared by
// every precision path: conditioning prep, attention scores, feature gather, LayerNorm,
// coupling update with log-det, prior log-prob.
#pragma once
#include "common.cuh"

namespace tw {

// View of the flat parameter table documented in include/timewarp_b200.h.
struct ParamView {
  const tw_flow_config* c;
  const void* const* p;
  __host__ int per_mlp() const { return 2 * (c->num_mlp_hidden + 1); }
  __host__ int per_net() const { return 2 * per_mlp() + 11 * c->num_transformer_layers; }
  __host__ int regular() const { return 3 + c->num_coupling_layers * 2 * per_net(); }
  __host__ bool chebyshev() const { return c->attention_type == TW_ATTENTION_CHEBYSHEV; }
  __host__ bool local() const { return c->attention_type == TW_ATTENTION_LOCAL; }
  __host__ int total() const { return regular() + (chebyshev() ? c->num_coupling_layers * 2 * c->num_transformer_layers : 0); }
  // chebyshev_kernel: cheb_coeffs [H, cheb_order] of encoder layer t of network `net` of coupling layer k (trailing section)
  __host__ const float* cheb(int k, int net, int t) const { return at(regular() + (k * 2 + net) * c->num_transformer_layers + t); }
  __host__ const float* at(int i) const { return (const float*)p[i]; }
  __host__ int net_base(int k, int net) const { return 3 + (k * 2 + net) * per_net(); }
  __host__ const float* embed() const { return at(0); }
  __host__ const float* log_scale_c() const { return at(1); }
  __host__ const float* log_scale_v() const { return at(2); }
  __host__ const float* in_w(int k, int net, int i) const { return at(net_base(k, net) + 2 * i); }
  __host__ const float* in_b(int k, int net, int i) const { return at(net_base(k, net) + 2 * i + 1); }
  // j: 0 wv, 1 lengthscales, 2 wo, 3 w1, 4 b1, 5 w2, 6 b2, 7 g1, 8 be1, 9 g2, 10 be2
  __host__ const float* enc(int k, int net, int t, int j) const { return at(net_base(k, net) + per_mlp() + 11 * t + j); }
  __host__ const float* out_w(int k, int net, int i) const {
    return at(net_base(k, net) + per_mlp() + 11 * c->num_transformer_layers + 2 * i);
  }
  __host__ const float* out_b(int k, int net, int i) const {
    return at(net_base(k, net) + per_mlp() + 11 * c->num_transformer_layers + 2 * i + 1);
  }
};

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

// Y[net] = act(X[net] W[net]^T + b[net]) (+ R[net]);  blockIdx.z selects the network so that the
// scale and shift conditioners of one coupling layer run in the same launch.
struct Lin2 {
  const float* X[2];
  const float* W[2];
  const float* b[2];
  const float* R[2];
  float* Y[2];
};

int launch_linear(const Lin2& a, int nets, int64_t M, int N, int K, int ldx, int ldr, int ldy, int act, cudaStream_t st);
int launch_layernorm(float* x0, float* x1, const float* g0, const float* g1, const float* b0, const float* b1, int nets,
                     int64_t M, int D, float eps, cudaStream_t st);
int launch_attn_mix(const float* scores, const float* v0, const float* v1, float* o0, float* o1, int nets, int64_t n,
                    int64_t n_cond, int V, int H, int Dv, cudaStream_t st);
// LocalSelfAttention (local_self_attention.py:46-119) on qkv [n*V, H*3*D] (per head: q | k | v): masked softmax over the
// atoms within max_radius of the conditioning positions xc [n_cond, V, 3]; out [n*V, H*D]
int launch_local_attn_bwd(const float* qkv0, const float* qkv1, const float* do0, const float* do1, float* dqkv0, float* dqkv1, int nets,
                          int64_t n, int64_t n_cond, int V, int H, int D, const float* xc, const uint8_t* mask, float max_radius,
                          cudaStream_t st);  // d(out) -> d(q | k | v)
int launch_local_attn(const float* qkv0, const float* qkv1, float* o0, float* o1, int nets, int64_t n, int64_t n_cond, int V, int H,
                      int D, const float* xc, const uint8_t* mask, float max_radius, cudaStream_t st);
int launch_prep(const float* x, const uint8_t* mask, int64_t n_cond, int V, float* xc, float* com, cudaStream_t st);
// cheb != nullptr: Chebyshev-rational basis with coefficients [H, order] instead of the Gaussian
int launch_scores(const float* xc, const uint8_t* mask, const float* ls, int64_t B, int V, int H, float* out, cudaStream_t st,
                  const float* cheb = nullptr, int order = 0, int force_zero = 0);
int launch_features(const float* embed, const int64_t* atom_types, const float* xc, const float* xv, const float* z_other,
                    int64_t n, int64_t n_cond, int V, int E, int n_types, float* feat, cudaStream_t st);
// forward: z = z*exp(s)+t ; reverse: z = (z-t)/exp(s); delta -= (+/-) sum log(exp(s)) over unmasked atoms
int launch_coupling(const float* s, const float* t, float* z, const uint8_t* mask, float* delta, int64_t n, int64_t n_cond,
                    int V, int reverse, float* out_scale, float* out_shift, cudaStream_t st);
// out[n] = prior(z) - delta (sign=-1) or + delta (sign=+1)
int launch_prior(const float* zc, const float* zv, const uint8_t* mask, const float* log_scale_c, const float* log_scale_v,
                 const float* delta, float sign, int64_t n, int64_t n_cond, int V, float* out, cudaStream_t st);
// y = (xc + com) + z  (flow.py:303-308)
int launch_uncentre(const float* xc, const float* com, const float* z, int64_t n, int64_t n_cond, int V, float* y, cudaStream_t st);
int launch_sub(const float* a, const float* b, int64_t count, float* out, cudaStream_t st);

}  // namespace tw

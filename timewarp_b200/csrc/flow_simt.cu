// CUDA-core fp32 kernels of the flow (generic layer sizes) + the small fused kernels.
// Arithmetic follows SURVEY.md Appendix A / the reference lines cited at each kernel.
#include <atomic>
#include <mutex>
#include <vector>

#include "flow_simt.cuh"

namespace tw {

// ------------------------------------------------------------------------------------------
// thread-local error string
char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- profiling hooks -----------------------------------------------------------------------
static int g_prof_class = PROF_NONE;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static size_t g_prof_used = 0;
static std::mutex g_prof_mu;

ProfScope::ProfScope(int cls, cudaStream_t s) : st(s), slot(-1) {
  if (cls != g_prof_class || cls == PROF_NONE) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_used == g_prof_events.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof_events.emplace_back(a, b);
  }
  slot = (int)g_prof_used++;
  cudaEventRecord(g_prof_events[slot].first, st);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof_events[slot].second, st);
}

// ------------------------------------------------------------------------------------------
// Y = act(X W^T + b) (+R).  64x64x16 tiles, 256 threads, 4x4 outputs per thread.
constexpr int LBM = 64, LBN = 64, LBK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SILU) return v / (1.f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(256) k_linear(Lin2 a, int64_t M, int N, int K, int ldx, int ldr, int ldy, int act) {
  const int net = blockIdx.z;
  const float* __restrict__ X = a.X[net];
  const float* __restrict__ W = a.W[net];
  const float* __restrict__ bias = a.b[net];
  const float* __restrict__ R = a.R[net];
  float* __restrict__ Y = a.Y[net];
  __shared__ __align__(16) float Xs[LBK][LBM + 4];
  __shared__ __align__(16) float Ws[LBK][LBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * LBM;
  const int n0 = blockIdx.x * LBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += LBK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256, r = e >> 4, c = e & 15;
      int gk = k0 + c;
      int64_t gm = m0 + r;
      int gn = n0 + r;
      Xs[c][r] = (gm < M && gk < K) ? X[gm * ldx + gk] : 0.f;
      Ws[c][r] = (gn < N && gk < K) ? W[(int64_t)gn * K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; kk++) {
      float4 xa = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      float4 wb = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float xv[4] = {xa.x, xa.y, xa.z, xa.w}, wv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      v = apply_act(v, act);
      if (R) v += R[gm * ldr + gn];
      Y[gm * ldy + gn] = v;
    }
  }
}

int launch_linear(const Lin2& a, int nets, int64_t M, int N, int K, int ldx, int ldr, int ldy, int act, cudaStream_t st) {
  if (M == 0) return TW_OK;
  int64_t gy = (M + LBM - 1) / LBM;
  TW_CHECK_ARG(gy <= 65535 * 32767LL, "linear: M too large");
  // gridDim.y is limited to 65535: fold the excess into several launches
  const int64_t max_rows = 65535LL * LBM;
  for (int64_t r0 = 0; r0 < M; r0 += max_rows) {
    int64_t rows = (M - r0 < max_rows) ? (M - r0) : max_rows;
    Lin2 b = a;
    for (int i = 0; i < 2; i++) {
      if (b.X[i]) b.X[i] += r0 * ldx;
      if (b.R[i]) b.R[i] += r0 * ldr;
      if (b.Y[i]) b.Y[i] += r0 * ldy;
    }
    dim3 grid((N + LBN - 1) / LBN, (unsigned)((rows + LBM - 1) / LBM), nets);
    k_linear<<<grid, 256, 0, st>>>(b, rows, N, K, ldx, ldr, ldy, act);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// In-place LayerNorm over the last dim (biased variance, eps inside sqrt): one warp per row.
// custom_attention_encoder.py:110,113 (nn.LayerNorm).
struct LN2 {
  float* x[2];
  const float* g[2];
  const float* b[2];
};
__global__ void __launch_bounds__(256) k_layernorm(LN2 a, int64_t M, int D, float eps) {
  const int net = blockIdx.y;
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float* x = a.x[net] + row * D;
  const float* g = a.g[net];
  const float* b = a.b[net];
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += x[i];
  float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) {
    float d = x[i] - mean;
    q = fmaf(d, d, q);
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
  for (int i = lane; i < D; i += 32) x[i] = (x[i] - mean) * rstd * g[i] + b[i];
}

int launch_layernorm(float* x0, float* x1, const float* g0, const float* g1, const float* b0, const float* b1, int nets,
                     int64_t M, int D, float eps, cudaStream_t st) {
  if (M == 0) return TW_OK;
  LN2 a;
  a.x[0] = x0, a.x[1] = x1, a.g[0] = g0, a.g[1] = g1, a.b[0] = b0, a.b[1] = b1;
  dim3 grid((unsigned)((M + 7) / 8), nets);
  k_layernorm<<<grid, 256, 0, st>>>(a, M, D, eps);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// attend + flatten_multihead (kernel_attention.py:124-156):
// out[n,i,h*Dv+d] = sum_j scores[n % n_cond,h,i,j] * vals[n,j,h*Dv+d]
constexpr int MIX_TI = 8;
__global__ void __launch_bounds__(128) k_attn_mix(const float* __restrict__ scores, const float* __restrict__ v0,
                                                  const float* __restrict__ v1, float* __restrict__ o0,
                                                  float* __restrict__ o1, int64_t n_cond, int V, int H, int Dv) {
  extern __shared__ float sA[];  // [MIX_TI][V]
  const int64_t n = blockIdx.x;
  const int h = blockIdx.y, net = blockIdx.z;
  const float* A = scores + ((n % n_cond) * H + h) * (int64_t)V * V;
  const int ld = H * Dv;
  const float* v = (net ? v1 : v0) + n * (int64_t)V * ld + h * Dv;
  float* o = (net ? o1 : o0) + n * (int64_t)V * ld + h * Dv;
  for (int i0 = 0; i0 < V; i0 += MIX_TI) {
    int ti = min(MIX_TI, V - i0);
    for (int e = threadIdx.x; e < ti * V; e += blockDim.x) sA[e] = A[(int64_t)i0 * V + e];
    __syncthreads();
    for (int d = threadIdx.x; d < Dv; d += blockDim.x) {
      float acc[MIX_TI];
#pragma unroll
      for (int ii = 0; ii < MIX_TI; ii++) acc[ii] = 0.f;
      for (int j = 0; j < V; j++) {
        float vj = v[(int64_t)j * ld + d];
#pragma unroll
        for (int ii = 0; ii < MIX_TI; ii++)
          if (ii < ti) acc[ii] = fmaf(sA[ii * V + j], vj, acc[ii]);
      }
#pragma unroll
      for (int ii = 0; ii < MIX_TI; ii++)
        if (ii < ti) o[(int64_t)(i0 + ii) * ld + d] = acc[ii];
    }
    __syncthreads();
  }
}

int launch_attn_mix(const float* scores, const float* v0, const float* v1, float* o0, float* o1, int nets, int64_t n,
                    int64_t n_cond, int V, int H, int Dv, cudaStream_t st) {
  if (n == 0) return TW_OK;
  TW_CHECK_ARG(n <= 2147483647LL, "attn_mix: too many samples");
  size_t smem = (size_t)MIX_TI * V * sizeof(float);
  TW_CHECK_ARG(smem <= 48 * 1024, "attn_mix: V=%d too large", V);
  dim3 grid((unsigned)n, H, nets);
  k_attn_mix<<<grid, 128, smem, st>>>(scores, v0, v1, o0, o1, n_cond, V, H, Dv);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// LocalSelfAttention (modules/layers/local_self_attention.py:46-119).  The reference gathers the atoms within max_radius
// (topk) and softmaxes over them; equivalently a masked softmax over all atoms of the sample.  One CTA per (sample, head,
// network): K and V of the head in shared memory, one warp per query row (lane = 4 features of d_model = 128... generic D).
__global__ void __launch_bounds__(128) k_local_attn(const float* __restrict__ qkv0, const float* __restrict__ qkv1,
                                                    float* __restrict__ o0, float* __restrict__ o1, int64_t n_cond, int V, int H,
                                                    int D, const float* __restrict__ xc, const uint8_t* __restrict__ mask,
                                                    float max_radius, float inv_sqrt_d) {
  extern __shared__ float sm[];
  float* sK = sm;                    // [V][D]
  float* sV = sK + (size_t)V * D;    // [V][D]
  float* sX = sV + (size_t)V * D;    // [V][3]
  float* sW = sX + (size_t)V * 3;    // [4 warps][V] attention weights of the row in flight
  const int64_t n = blockIdx.x;
  const int h = blockIdx.y, net = blockIdx.z;
  const int64_t nc = n % n_cond;
  const int ld = H * 3 * D;
  const float* qkv = (net ? qkv1 : qkv0) + n * (int64_t)V * ld + (size_t)h * 3 * D;
  float* out = (net ? o1 : o0) + n * (int64_t)V * (H * D) + (size_t)h * D;
  const uint8_t* mb = mask + nc * V;
  for (int e = threadIdx.x; e < V * D; e += blockDim.x) {
    const int j = e / D, d = e % D;
    sK[e] = qkv[(int64_t)j * ld + D + d];
    sV[e] = qkv[(int64_t)j * ld + 2 * D + d];
  }
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) sX[e] = xc[nc * V * 3 + e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* w = sW + warp * V;
  for (int i = warp; i < V; i += 4) {
    const float* q = qkv + (int64_t)i * ld;
    const float xi = sX[i * 3], yi = sX[i * 3 + 1], zi = sX[i * 3 + 2];
    float mx = -INFINITY;
    for (int j = 0; j < V; j++) {
      float part = 0.f;
      for (int d = lane; d < D; d += 32) part = fmaf(q[d], sK[j * D + d], part);
      part = warp_sum(part) * inv_sqrt_d;
      const float dx = xi - sX[j * 3], dy = yi - sX[j * 3 + 1], dz = zi - sX[j * 3 + 2];
      const bool outside = mb[i] || mb[j] || sqrtf(dx * dx + dy * dy + dz * dz) > max_radius;
      const float sc = outside ? -INFINITY : part;
      if (lane == 0) w[j] = sc;
      mx = fmaxf(mx, sc);
    }
    __syncwarp();
    float den = 0.f;
    for (int j = lane; j < V; j += 32) {
      const float e = (w[j] == -INFINITY) ? 0.f : expf(w[j] - mx);  // a row with no neighbour (padding) stays all zero
      w[j] = e;
      den += e;
    }
    den = warp_sum(den);
    const float inv = den > 0.f ? 1.f / den : 0.f;
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < V; j++) acc = fmaf(w[j] * inv, sV[j * D + d], acc);
      out[(int64_t)i * (H * D) + d] = acc;
    }
    __syncwarp();
  }
}

int launch_local_attn(const float* qkv0, const float* qkv1, float* o0, float* o1, int nets, int64_t n, int64_t n_cond, int V, int H,
                      int D, const float* xc, const uint8_t* mask, float max_radius, cudaStream_t st) {
  if (n == 0) return TW_OK;
  const size_t smem = ((size_t)2 * V * D + (size_t)V * 3 + (size_t)4 * V) * sizeof(float);
  TW_CHECK_ARG(smem <= 200 * 1024, "local attention: V * d_model too large for shared memory (V=%d)", V);
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_local_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done.mark();
  }
  dim3 grid((unsigned)n, H, nets);
  k_local_attn<<<grid, 128, smem, st>>>(qkv0, qkv1, o0, o1, n_cond, V, H, D, xc, mask, max_radius, 1.0f / sqrtf((float)D));
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// Backward of the masked-softmax attention above (training path of `local` attention, local_self_attention.py:46-119 under
// autograd): given d out, the gradient w.r.t. q | k | v of every head.  The radius mask is a hard function of the positions, so no
// gradient flows to them.  One CTA per (sample, head, network); pass A: one warp per query row i -- P_i. and dS_i. = P_i. (dP_i. -
// sum_j P_ij dP_ij) into shared memory, dq_i = dS_i. K / sqrt(D); pass B: one warp per key row j -- dk_j = dS_.j^T Q / sqrt(D),
// dv_j = P_.j^T d out.  K / V / Q rows are read from global memory (L1 / L2: one head of one sample is V * 3 D floats).
__global__ void __launch_bounds__(128) k_local_attn_bwd(const float* __restrict__ qkv0, const float* __restrict__ qkv1,
                                                        const float* __restrict__ do0, const float* __restrict__ do1,
                                                        float* __restrict__ dqkv0, float* __restrict__ dqkv1, int64_t n_cond, int V, int H,
                                                        int D, const float* __restrict__ xc, const uint8_t* __restrict__ mask,
                                                        float max_radius, float inv_sqrt_d) {
  extern __shared__ float sm[];
  float* sP = sm;                     // [V][V] attention weights
  float* sS = sP + (size_t)V * V;     // [V][V] gradient w.r.t. the scores
  float* sX = sS + (size_t)V * V;     // [V][3]
  const int64_t n = blockIdx.x;
  const int h = blockIdx.y, net = blockIdx.z;
  const int64_t nc = n % n_cond;
  const int ld = H * 3 * D;
  const float* qkv = (net ? qkv1 : qkv0) + n * (int64_t)V * ld + (size_t)h * 3 * D;
  const float* dout = (net ? do1 : do0) + n * (int64_t)V * (H * D) + (size_t)h * D;
  float* dqkv = (net ? dqkv1 : dqkv0) + n * (int64_t)V * ld + (size_t)h * 3 * D;
  const uint8_t* mb = mask + nc * V;
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) sX[e] = xc[nc * V * 3 + e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < V; i += 4) {
    const float* q = qkv + (int64_t)i * ld;
    const float* g = dout + (int64_t)i * (H * D);
    float* p = sP + (size_t)i * V;
    float* ds = sS + (size_t)i * V;
    const float xi = sX[i * 3], yi = sX[i * 3 + 1], zi = sX[i * 3 + 2];
    float mx = -INFINITY;
    for (int j = 0; j < V; j++) {
      const float* kj = qkv + (int64_t)j * ld + D;
      float part = 0.f;
      for (int d = lane; d < D; d += 32) part = fmaf(q[d], kj[d], part);
      part = warp_sum(part) * inv_sqrt_d;
      const float dx = xi - sX[j * 3], dy = yi - sX[j * 3 + 1], dz = zi - sX[j * 3 + 2];
      const bool outside = mb[i] || mb[j] || sqrtf(dx * dx + dy * dy + dz * dz) > max_radius;
      const float sc = outside ? -INFINITY : part;
      if (lane == 0) p[j] = sc;
      mx = fmaxf(mx, sc);
    }
    __syncwarp();
    float den = 0.f;
    for (int j = lane; j < V; j += 32) {
      const float e = (p[j] == -INFINITY) ? 0.f : expf(p[j] - mx);
      p[j] = e;
      den += e;
    }
    den = warp_sum(den);
    const float inv = den > 0.f ? 1.f / den : 0.f;
    __syncwarp();
    float rs = 0.f;  // sum_j P_ij dP_ij (every lane ends up with the same value)
    for (int j = 0; j < V; j++) {
      const float pij = p[j] * inv;
      float dp = 0.f;
      if (pij != 0.f) {  // (warp-uniform)
        const float* vj = qkv + (int64_t)j * ld + 2 * D;
        for (int d = lane; d < D; d += 32) dp = fmaf(g[d], vj[d], dp);
        dp = warp_sum(dp);
      }
      if (lane == 0) ds[j] = dp;
      rs = fmaf(pij, dp, rs);
    }
    __syncwarp();
    for (int j = lane; j < V; j += 32) {
      const float pij = p[j] * inv;
      p[j] = pij;
      ds[j] = pij * (ds[j] - rs);
    }
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < V; j++) {
        const float w = ds[j];
        if (w != 0.f) acc = fmaf(w, qkv[(int64_t)j * ld + D + d], acc);
      }
      dqkv[(int64_t)i * ld + d] = acc * inv_sqrt_d;
    }
  }
  __syncthreads();
  for (int j = warp; j < V; j += 4) {
    for (int d = lane; d < D; d += 32) {
      float dk = 0.f, dv = 0.f;
      for (int i = 0; i < V; i++) {
        const float w = sS[(size_t)i * V + j], pw = sP[(size_t)i * V + j];
        if (w != 0.f) dk = fmaf(w, qkv[(int64_t)i * ld + d], dk);
        if (pw != 0.f) dv = fmaf(pw, dout[(int64_t)i * (H * D) + d], dv);
      }
      dqkv[(int64_t)j * ld + D + d] = dk * inv_sqrt_d;
      dqkv[(int64_t)j * ld + 2 * D + d] = dv;
    }
  }
}

int launch_local_attn_bwd(const float* qkv0, const float* qkv1, const float* do0, const float* do1, float* dqkv0, float* dqkv1, int nets,
                          int64_t n, int64_t n_cond, int V, int H, int D, const float* xc, const uint8_t* mask, float max_radius,
                          cudaStream_t st) {
  if (n == 0) return TW_OK;
  const size_t smem = ((size_t)2 * V * V + (size_t)V * 3) * sizeof(float);
  TW_CHECK_ARG(smem <= 200 * 1024, "local attention backward: V=%d too large for shared memory", V);
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_local_attn_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done.mark();
  }
  dim3 grid((unsigned)n, H, nets);
  k_local_attn_bwd<<<grid, 128, smem, st>>>(qkv0, qkv1, do0, do1, dqkv0, dqkv1, n_cond, V, H, D, xc, mask, max_radius,
                                            1.0f / sqrtf((float)D));
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// get_centre_of_mass + centring (utils/molecule_utils.py:15-29, flow.py:156-157): one block / state
__global__ void __launch_bounds__(128) k_prep(const float* __restrict__ x, const uint8_t* __restrict__ mask, int V,
                                              float* __restrict__ xc, float* __restrict__ com) {
  __shared__ float red[33];
  const int64_t b = blockIdx.x;
  const float* xb = x + b * V * 3;
  const uint8_t* mb = mask + b * V;
  float s[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    if (!mb[v]) {
      s[0] += xb[v * 3 + 0], s[1] += xb[v * 3 + 1], s[2] += xb[v * 3 + 2];
      cnt += 1.f;
    }
  }
  float c[3];
  cnt = block_sum(cnt, red);
  for (int k = 0; k < 3; k++) c[k] = block_sum(s[k], red) / cnt;
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) xc[b * V * 3 + e] = xb[e] - c[e % 3];
  if (threadIdx.x < 3) com[b * 3 + threadIdx.x] = c[threadIdx.x];
}

int launch_prep(const float* x, const uint8_t* mask, int64_t n_cond, int V, float* xc, float* com, cudaStream_t st) {
  if (n_cond == 0) return TW_OK;
  k_prep<<<(unsigned)n_cond, 128, 0, st>>>(x, mask, V, xc, com);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// compute_kernel_attention_scores (kernel_attention.py:69-121).  One warp per query row (b,i);
// distances from direct differences (exactly 0 on the diagonal).
__global__ void __launch_bounds__(256) k_scores(const float* __restrict__ xc, const uint8_t* __restrict__ mask,
                                                const float* __restrict__ ls, int64_t B, int V, int H,
                                                float* __restrict__ out, const float* __restrict__ cheb, int order, int force_zero) {
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B * V) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / V;
  const int i = (int)(row % V);
  const float* xb = xc + b * V * 3;
  const uint8_t* mb = mask + b * V;
  const float xi = xb[i * 3], yi = xb[i * 3 + 1], zi = xb[i * 3 + 2];
  for (int h = 0; h < H; h++) {
    const float l = ls[h];
    const float* coef = cheb ? cheb + (size_t)h * order : nullptr;
    const float cmean = cheb_mean(coef, order, force_zero);
    float* o = out + ((b * H + h) * V + i) * (int64_t)V;
    float sum = 0.f;
    for (int j = lane; j < V; j += 32) {
      float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
      float d = sqrtf(dx * dx + dy * dy + dz * dz);
      float a = d / l;
      float w = mb[j] ? 0.f : attention_basis(a, coef, order, cmean);
      o[j] = w;
      sum += fabsf(w);
    }
    sum = warp_sum(sum) + 1e-5f;
    for (int j = lane; j < V; j += 32) o[j] = o[j] / sum;
  }
}

int launch_scores(const float* xc, const uint8_t* mask, const float* ls, int64_t B, int V, int H, float* out, cudaStream_t st,
                  const float* cheb, int order, int force_zero) {
  if (B == 0) return TW_OK;
  int64_t rows = B * V;
  k_scores<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(xc, mask, ls, B, V, H, out, cheb, order, force_zero);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// nn.Embedding + torch.cat (flow.py:172, custom_transformer_nvp.py:64-71): feat[m] = (emb, xc, xv, z_other)
__global__ void __launch_bounds__(256) k_features(const float* __restrict__ embed, const int64_t* __restrict__ atom_types,
                                                  const float* __restrict__ xc, const float* __restrict__ xv,
                                                  const float* __restrict__ z_other, int64_t n, int64_t n_cond, int V,
                                                  int E, int n_types, float* __restrict__ feat) {
  const int F = E + 9;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * V * F) return;
  int c = (int)(idx % F);
  int64_t m = idx / F;
  int64_t s = m / V;
  int v = (int)(m % V);
  int64_t mc = (s % n_cond) * V + v;
  float val;
  if (c < E) {
    int64_t t = atom_types[mc];
    t = t < 0 ? 0 : (t >= n_types ? n_types - 1 : t);
    val = embed[t * E + c];
  } else if (c < E + 3)
    val = xc[mc * 3 + (c - E)];
  else if (c < E + 6)
    val = xv[mc * 3 + (c - E - 3)];
  else
    val = z_other[m * 3 + (c - E - 6)];
  feat[idx] = val;
}

int launch_features(const float* embed, const int64_t* atom_types, const float* xc, const float* xv, const float* z_other,
                    int64_t n, int64_t n_cond, int V, int E, int n_types, float* feat, cudaStream_t st) {
  int64_t total = n * V * (E + 9);
  if (total == 0) return TW_OK;
  k_features<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(embed, atom_types, xc, xv, z_other, n, n_cond, V, E, n_types, feat);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// Affine coupling + log-det (nvp.py:127-133, 175-181, 86).  One block per sample.
__global__ void __launch_bounds__(128) k_coupling(const float* __restrict__ s, const float* __restrict__ t,
                                                  float* __restrict__ z, const uint8_t* __restrict__ mask,
                                                  float* __restrict__ delta, int64_t n_cond, int V, int reverse,
                                                  float* __restrict__ out_scale, float* __restrict__ out_shift) {
  __shared__ float red[33];
  const int64_t n = blockIdx.x;
  const uint8_t* mb = mask + (n % n_cond) * V;
  float acc = 0.f;
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) {
    int64_t g = n * V * 3 + e;
    float sv = s[g], tv = t[g];
    float sc = expf(sv);          // scale = exp(scale_transformer(...))   custom_transformer_nvp.py:76
    float ls = logf(sc);          // torch.log(scale)                      nvp.py:127
    if (!mb[e / 3]) acc += ls;
    if (z) {
      float zv = z[g];
      z[g] = reverse ? (zv - tv) / sc : fmaf(zv, sc, tv);  // fmaf vs mul+add: <= 1 ulp
    }
    if (out_scale) out_scale[g] = sc;
    if (out_shift) out_shift[g] = tv;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && delta) delta[n] -= reverse ? -acc : acc;
}

int launch_coupling(const float* s, const float* t, float* z, const uint8_t* mask, float* delta, int64_t n, int64_t n_cond,
                    int V, int reverse, float* out_scale, float* out_shift, cudaStream_t st) {
  if (n == 0) return TW_OK;
  k_coupling<<<(unsigned)n, 128, 0, st>>>(s, t, z, mask, delta, n_cond, V, reverse, out_scale, out_shift);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ------------------------------------------------------------------------------------------
// Gaussian prior log-prob with learned log-scales (flow.py:159-166,191-203 / 322-334)
__global__ void __launch_bounds__(128) k_prior(const float* __restrict__ zc, const float* __restrict__ zv,
                                               const uint8_t* __restrict__ mask, const float* __restrict__ lsc,
                                               const float* __restrict__ lsv, const float* __restrict__ delta, float sign,
                                               int64_t n_cond, int V, float* __restrict__ out) {
  __shared__ float red[33];
  const int64_t n = blockIdx.x;
  const uint8_t* mb = mask + (n % n_cond) * V;
  const float sc = expf(lsc[0]), sv = expf(lsv[0]);
  const float var_c = sc * sc, var_v = sv * sv, log_c = logf(sc), log_v = logf(sv);
  const float half_log_2pi = 0.91893853320467274178f;  // log(sqrt(2*pi))
  float acc = 0.f;
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) {
    if (mb[e / 3]) continue;
    int64_t g = n * V * 3 + e;
    float a = zc[g], b = zv[g];
    acc += -(a * a) / (2.f * var_c) - log_c - half_log_2pi;
    acc += -(b * b) / (2.f * var_v) - log_v - half_log_2pi;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[n] = acc + sign * delta[n];
}

int launch_prior(const float* zc, const float* zv, const uint8_t* mask, const float* log_scale_c, const float* log_scale_v,
                 const float* delta, float sign, int64_t n, int64_t n_cond, int V, float* out, cudaStream_t st) {
  if (n == 0) return TW_OK;
  k_prior<<<(unsigned)n, 128, 0, st>>>(zc, zv, mask, log_scale_c, log_scale_v, delta, sign, n_cond, V, out);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

__global__ void k_uncentre(const float* __restrict__ xc, const float* __restrict__ com, const float* __restrict__ z,
                           int64_t total, int64_t n_cond, int V, float* __restrict__ y) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t m = idx / 3;
  int c = (int)(idx % 3);
  int64_t s = m / V;
  int v = (int)(m % V);
  int64_t b = s % n_cond;
  y[idx] = (xc[(b * V + v) * 3 + c] + com[b * 3 + c]) + z[idx];
}

int launch_uncentre(const float* xc, const float* com, const float* z, int64_t n, int64_t n_cond, int V, float* y, cudaStream_t st) {
  int64_t total = n * V * 3;
  if (total == 0) return TW_OK;
  k_uncentre<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xc, com, z, total, n_cond, V, y);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

__global__ void k_sub(const float* __restrict__ a, const float* __restrict__ b, int64_t count, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) out[idx] = a[idx] - b[idx];
}
int launch_sub(const float* a, const float* b, int64_t count, float* out, cudaStream_t st) {
  if (count == 0) return TW_OK;
  k_sub<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(a, b, count, out);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // namespace tw

extern "C" {

long long tw_debug_launch_count(void) { return tw::g_launches.load(); }

int tw_prof_enable(int kernel_class) {
  std::lock_guard<std::mutex> lk(tw::g_prof_mu);
  tw::g_prof_class = kernel_class;
  tw::g_prof_used = 0;
  return TW_OK;
}

int tw_prof_collect(double* total_ms, long long* scopes) {
  std::lock_guard<std::mutex> lk(tw::g_prof_mu);
  double tot = 0;
  for (size_t i = 0; i < tw::g_prof_used; i++) {
    float ms = 0;
    cudaError_t e = cudaEventSynchronize(tw::g_prof_events[i].second);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, tw::g_prof_events[i].first, tw::g_prof_events[i].second);
    if (e != cudaSuccess) return tw::fail(TW_ERR_CUDA, "prof collect: %s", cudaGetErrorString(e));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (scopes) *scopes = (long long)tw::g_prof_used;
  tw::g_prof_used = 0;
  return TW_OK;
}

}  // extern "C"

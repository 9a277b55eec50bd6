// Metropolis-Hastings / exploration acceptance, kinetic energy and chirality veto -- the small
// per-chain kernels that keep the sampling loop on the device (no host sync per iteration).
#include <limits.h>

#include "common.cuh"

namespace tw {

// utils/evaluation_utils.py:663-674 (one row per proposal)
__global__ void k_mh_decide(const float* __restrict__ e_pot_x, const float* __restrict__ e_pot_y,
                            const float* __restrict__ e_kin_x, const float* __restrict__ e_kin_y,
                            const float* __restrict__ p_xy, const float* __restrict__ p_yx, const float* __restrict__ u,
                            int64_t n, float* __restrict__ out_exponent, float* __restrict__ out_p_acc,
                            uint8_t* __restrict__ out_accepted) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e_pot = e_pot_y[i] - e_pot_x[i];              // :644
  float e_kin = e_kin_y[i] - e_kin_x[i];              // :633
  float energy = e_pot + e_kin;                       // :646
  float ex = (energy + p_xy[i]) - p_yx[i];            // :663  exp = energy + p_xy - p_yx
  float p = expf(-ex);
  float p_acc = (p != p) ? p : fminf(1.0f, p);        // torch.min(1, exp(-exp)) propagates NaN (:665)
  bool acc = u[i] < p_acc;                            // :668 (NaN compares false => reject)
  if (out_exponent) out_exponent[i] = ex;
  if (out_p_acc) out_p_acc[i] = p_acc;
  out_accepted[i] = acc ? 1 : 0;
}

// x[n] <- y[n] where accepted (independent-chains form of :675-678)
__global__ void k_select_rows(float* __restrict__ xa, const float* __restrict__ ya, float* __restrict__ xb,
                              const float* __restrict__ yb, const uint8_t* __restrict__ take, int64_t n, int row) {
  int64_t i = blockIdx.x;
  if (!take[i]) return;
  for (int e = threadIdx.x; e < row; e += blockDim.x) {
    if (xa) xa[i * row + e] = ya[i * row + e];
    if (xb) xb[i * row + e] = yb[i * row + e];
  }
}

// first accepted proposal (:669-674) or -1
__global__ void k_first_accept(const uint8_t* __restrict__ acc, int64_t n, int32_t* __restrict__ out) {
  __shared__ int best[32];
  int v = INT_MAX;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
    if (acc[i]) {
      v = (int)i;
      break;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = INT_MAX;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = min(m, best[w]);
    out[0] = (m == INT_MAX) ? -1 : m;
  }
}

// exploration.py:243-246
__global__ void k_threshold(float* __restrict__ x, float* __restrict__ e_old, const float* __restrict__ y,
                            const float* __restrict__ e_new, float threshold, int row, uint8_t* __restrict__ out_accepted) {
  int64_t i = blockIdx.x;
  const bool keep_old = (e_new[i] - e_old[i]) > threshold;  // NaN => takes the new state, as the reference does
  __syncthreads();
  if (!keep_old) {
    for (int e = threadIdx.x; e < row; e += blockDim.x) x[i * row + e] = y[i * row + e];
    if (threadIdx.x == 0) e_old[i] = e_new[i];
  }
  if (threadIdx.x == 0 && out_accepted) out_accepted[i] = keep_old ? 0 : 1;
}

// utils/evaluation_utils.py:416-436
__global__ void __launch_bounds__(128) k_kinetic(const float* __restrict__ v, const float* __restrict__ masses, float inv_kbT,
                                                 int V, float* __restrict__ out) {
  __shared__ float red[33];
  int64_t b = blockIdx.x;
  float acc = 0.f;
  for (int a = threadIdx.x; a < V; a += blockDim.x) {
    const float* p = v + (b * V + a) * 3;
    float s = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    acc += masses ? masses[a] * s : s;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[b] = masses ? 0.5f * acc * inv_kbT : 0.5f * acc;
}

// utils/chirality.py:41-80; products/sums rounded separately (no FMA contraction) so that the sign
// of a near-zero triple product is evaluated like the torch ops it restates.
__global__ void k_chirality(const float* __restrict__ coords, const int64_t* __restrict__ centers,
                            const float* __restrict__ ref_signs, int64_t B, int V, int C, uint8_t* __restrict__ out,
                            float* __restrict__ out_signs) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* x = coords + b * V * 3;
  bool changed = false;
  for (int c = 0; c < C; c++) {
    const float* p0 = x + centers[c * 4] * 3;
    float d[3][3];
    for (int k = 0; k < 3; k++) {
      const float* pk = x + centers[c * 4 + 1 + k] * 3;
      for (int a = 0; a < 3; a++) d[k][a] = __fsub_rn(pk[a], p0[a]);
    }
    float cx = __fsub_rn(__fmul_rn(d[1][1], d[2][2]), __fmul_rn(d[1][2], d[2][1]));
    float cy = __fsub_rn(__fmul_rn(d[1][2], d[2][0]), __fmul_rn(d[1][0], d[2][2]));
    float cz = __fsub_rn(__fmul_rn(d[1][0], d[2][1]), __fmul_rn(d[1][1], d[2][0]));
    float t = __fadd_rn(__fadd_rn(__fmul_rn(d[0][0], cx), __fmul_rn(d[0][1], cy)), __fmul_rn(d[0][2], cz));
    float sg = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : t);  // torch.sign (NaN stays NaN)
    if (out_signs) out_signs[b * C + c] = sg;
    if (ref_signs && sg != ref_signs[c]) changed = true;
  }
  if (out) out[b] = changed ? 1 : 0;
}

}  // namespace tw

using namespace tw;

extern "C" {

int tw_mh_accept(const float* e_pot_x, const float* e_pot_y, const float* e_kin_x, const float* e_kin_y, const float* p_xy,
                 const float* p_yx, const float* u, int64_t n, int64_t V, float* x_coords, float* x_velocs,
                 const float* y_coords, const float* y_velocs, float* out_exponent, float* out_p_acc, uint8_t* out_accepted,
                 int32_t* out_first_accept, void* stream) {
  TW_CHECK_ARG(e_pot_x && e_pot_y && e_kin_x && e_kin_y && p_xy && p_yx && u && out_accepted, "NULL pointer");
  TW_CHECK_ARG(n >= 0 && n <= 2147483647LL && V >= 1, "bad sizes");
  TW_CHECK_ARG((!x_coords || y_coords) && (!x_velocs || y_velocs), "state update needs the proposal tensors");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (out_first_accept) TW_CUDA(cudaMemsetAsync(out_first_accept, 0xff, sizeof(int32_t), st));
    return TW_OK;
  }
  k_mh_decide<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e_pot_x, e_pot_y, e_kin_x, e_kin_y, p_xy, p_yx, u, n, out_exponent,
                                                          out_p_acc, out_accepted);
  TW_LAUNCH_CHECK();
  if (x_coords || x_velocs) {
    k_select_rows<<<(unsigned)n, 128, 0, st>>>(x_coords, y_coords, x_velocs, y_velocs, out_accepted, n, (int)V * 3);
    TW_LAUNCH_CHECK();
  }
  if (out_first_accept) {
    k_first_accept<<<1, 256, 0, st>>>(out_accepted, n, out_first_accept);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

int tw_threshold_accept(float* x_coords, float* e_old, const float* y_coords, const float* e_new, float threshold, int64_t n,
                        int64_t V, uint8_t* out_accepted, void* stream) {
  TW_CHECK_ARG(x_coords && e_old && y_coords && e_new, "NULL pointer");
  TW_CHECK_ARG(n >= 0 && n <= 2147483647LL && V >= 1, "bad sizes");
  if (n == 0) return TW_OK;
  k_threshold<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(x_coords, e_old, y_coords, e_new, threshold, (int)V * 3, out_accepted);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

int tw_kinetic_energy(const float* velocs, const float* masses, float inv_kbT, int64_t B, int64_t V, float* out, void* stream) {
  TW_CHECK_ARG(velocs && out, "NULL pointer");
  TW_CHECK_ARG(B >= 0 && B <= 2147483647LL && V >= 1, "bad sizes");
  if (B == 0) return TW_OK;
  k_kinetic<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(velocs, masses, inv_kbT, (int)V, out);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

int tw_chirality(const float* coords, const int64_t* centers, const float* ref_signs, int64_t B, int64_t V, int32_t C,
                 uint8_t* out_changed, float* out_signs, void* stream) {
  TW_CHECK_ARG(coords && (out_changed || out_signs) && (C == 0 || centers), "NULL pointer");
  TW_CHECK_ARG(!out_changed || C == 0 || ref_signs, "out_changed needs ref_signs");
  TW_CHECK_ARG(B >= 0 && V >= 1 && C >= 0, "bad sizes");
  if (B == 0) return TW_OK;
  k_chirality<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(coords, centers, ref_signs, B, (int)V, C, out_changed,
                                                                             out_signs);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // extern "C"

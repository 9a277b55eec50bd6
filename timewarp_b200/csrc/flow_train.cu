// Training path of the flow: log-likelihood forward that records a tape, and the backward pass
// (gradients of sum_b g_b * log p(y_b | x_b) w.r.t. every trainable parameter).
//
// Reference semantics: torch autograd through ConditionalFlowDensityModel.log_likelihood
// (modules/model_wrappers/flow.py:131-215) as driven by the NLL loss (losses.py:321-356,
// density_model_base.py:27-42).  The reference has no hand-written backward; the arithmetic restated
// here is the chain rule of the forward pass documented in flow_api.cu / flow_tc.cu.
//
// Design.  The forward pass is the inference pass (same fused tcgen05 kernels) writing every layer
// boundary to its own buffer (the tape): layer inputs, post-LN1 activations, the two pre-LayerNorm
// sums, the per-head neighbourhood averages (already stored as bf16 operand images) and the flow state
// before each coupling layer.  The backward pass re-computes the wide intermediates it needs (FFN / MLP
// pre-activations) and evaluates every contraction with ONE generic tcgen05 GEMM over operand images:
//   NT  C[m,n] = sum_k A[m,k] B[n,k]      forward-shaped (re-computation)
//   NN  C[m,j] = sum_n A[m,n] W[n,j]      data gradient: the SAME packed weight image read MN-major
//   TN  C[i,j] = sum_m A[m,i] B[m,j]      weight gradient: both activations read MN-major, split over
//                                         token blocks, fp32 atomics into the gradient tensor
// An operand image is a grid of [128 rows x 64 cols] bf16 tiles in the 128-byte-swizzled K-major layout
// (hi and lo parts); the identical bytes are a valid MN-major SW128 tile for the transposed contraction,
// so no transposed copy of any weight or activation is ever materialised.
#include "flow_tc.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

// ============================================================================================
// operand images
struct ImgRef {
  const uint8_t* base;
  uint32_t sr;   // bytes between row tiles (128 rows)
  uint32_t sc1;  // bytes between column tile 2q and 2q+1
  uint32_t sc2;  // bytes between column tile pairs
  uint32_t lo;   // offset of the lo part of a tile
};
__host__ __device__ __forceinline__ const uint8_t* img_tile(const ImgRef& r, int tr, int tc) {
  return r.base + (size_t)tr * r.sr + (size_t)(tc >> 1) * r.sc2 + (size_t)(tc & 1) * r.sc1;
}
// plain activation image of a [rows x 64*n_ct] matrix: tile (tr, tc) at (tr*n_ct + tc) * 32 KB, hi 16 KB | lo 16 KB
static ImgRef plain_img(const uint8_t* base, int n_ct) { return ImgRef{base, (uint32_t)n_ct * 32768u, 32768u, 65536u, 16384u}; }
static size_t plain_img_bytes(int64_t rows, int n_ct) { return (size_t)((rows + 127) / 128) * n_ct * 32768; }

enum GemmMode { GEMM_NT = 0, GEMM_NN = 1, GEMM_NN_HEADED = 2, GEMM_TN = 3 };

struct GemmArgs {
  ImgRef A[2], B[2];
  float* C[2];
  const float* bias[2];
  const float* resid[2];
  int ldc, ldr;
  int mode;
  int rows, cols;        // valid extent of C
  int tiles_m, tiles_n;  // output tiles of 128 x bn
  int bn;                // 128 or 64
  int KB;                // 64-wide blocks of the contraction
  int splits;            // TN: the contraction is split over `splits` CTAs per output tile
  int accumulate;        // NT/NN: C += result
  int nets;              // grid.y (0 -> 2: scale and shift network in one launch)
};

constexpr int kGemmStages = 3;
constexpr int kGemmStageBytes = 65536;  // A hi 16K | A lo 16K | B hi 16K | B lo 16K

// Warp 0 streams operand tiles (bulk async copies), warp 1 issues the MMAs, warps 2-5 drain the
// double-buffered TMEM accumulator.  Persistent over (output tile, split) work items; blockIdx.y = network.
template <int kSplit>
__global__ void __launch_bounds__(192, 1) k_gemm_img(GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kGemmStages;
  uint64_t* acc_full = empty + kGemmStages;  // [2]
  uint64_t* acc_free = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
  if (tid == 0) {
    for (int i = 0; i < kGemmStages; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; i++) mbar_init(&acc_full[i], 1), mbar_init(&acc_free[i], 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int mode = a.mode, bn = a.bn;
  const bool a_mn = (mode == GEMM_TN), b_mn = (mode != GEMM_NT);
  const int n_work = a.tiles_m * a.tiles_n * a.splits;
  const int kb_chunk = (a.KB + a.splits - 1) / a.splits;
  const uint32_t parts = kSplit == 3 ? 2u : 1u;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    const ImgRef A = a.A[net], B = a.B[net];
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w % a.splits, t = w / a.splits, tn = t % a.tiles_n, tm = t / a.tiles_n;
      const int k0 = sp * kb_chunk, k1 = min(a.KB, k0 + kb_chunk);
      for (int kb = k0; kb < k1; kb++) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = ring + stage * kGemmStageBytes;
          mbar_arrive_expect_tx(&full[stage], parts * (16384u + (uint32_t)bn * 128u));
          if (!a_mn) {
            const uint8_t* src = img_tile(A, tm, kb);
            bulk_g2s(dst, src, 16384, &full[stage]);
            if (kSplit == 3) bulk_g2s(dst + 16384, src + A.lo, 16384, &full[stage]);
          } else {  // 64 token rows of the two column tiles that hold output rows tm*128 .. +127
            for (int c = 0; c < 2; c++) {
              const uint8_t* src = img_tile(A, kb >> 1, 2 * tm + c) + (kb & 1) * 8192;
              bulk_g2s(dst + c * 8192, src, 8192, &full[stage]);
              if (kSplit == 3) bulk_g2s(dst + 16384 + c * 8192, src + A.lo, 8192, &full[stage]);
            }
          }
          if (!b_mn) {
            const uint8_t* src = img_tile(B, tn, kb);
            bulk_g2s(dst + 32768, src, bn * 128, &full[stage]);
            if (kSplit == 3) bulk_g2s(dst + 49152, src + B.lo, bn * 128, &full[stage]);
          } else {
            const int tr = (mode == GEMM_NN_HEADED) ? 0 : (kb >> 1);
            const int tc0 = (mode == GEMM_NN_HEADED) ? 2 * (kb >> 1) : (bn == 128 ? 2 * tn : tn);
            for (int c = 0; c < bn / 64; c++) {
              const uint8_t* src = img_tile(B, tr, tc0 + c) + (kb & 1) * 8192;
              bulk_g2s(dst + 32768 + c * 8192, src, 8192, &full[stage]);
              if (kSplit == 3) bulk_g2s(dst + 49152 + c * 8192, src + B.lo, 8192, &full[stage]);
            }
          }
        }
        __syncwarp();
        if (++stage == kGemmStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, ph_free[2] = {0, 0};
    const uint32_t idesc = make_idesc_bf16(128, bn, a_mn ? 1 : 0, b_mn ? 1 : 0);
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w % a.splits;
      const int k0 = sp * kb_chunk, k1 = min(a.KB, k0 + kb_chunk);
      if (k0 >= k1) continue;
      const int tb = it & 1;
      if (it >= 2) {
        mbar_wait(&acc_free[tb], ph_free[tb]);
        ph_free[tb] ^= 1;
      }
      it++;
      tc_fence_after();
      const uint32_t d = tmem + tb * 128;
      for (int kb = k0; kb < k1; kb++) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = smem_u32(ring + stage * kGemmStageBytes);
          const uint32_t ahi = base, alo = base + 16384, bhi = base + 32768, blo = base + 49152;
#pragma unroll
          for (int term = 0; term < (kSplit == 3 ? 3 : 1); term++) {
            const uint32_t ab = (term == 1) ? alo : ahi, bb = (term == 2) ? blo : bhi;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint64_t ad = a_mn ? make_smem_desc(ab + k * 2048, 8192, 1024, LAYOUT_SW128) : desc_kmajor_sw128(ab + k * 32);
              const uint64_t bd = b_mn ? make_smem_desc(bb + k * 2048, 8192, 1024, LAYOUT_SW128) : desc_kmajor_sw128(bb + k * 32);
              mma_ss(d, ad, bd, idesc, (kb > k0 || term > 0 || k > 0) ? 1 : 0);
            }
          }
          mma_commit(&empty[stage]);
          if (kb == k1 - 1) mma_commit(&acc_full[tb]);
        }
        __syncwarp();
        if (++stage == kGemmStages) stage = 0, phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_full[2] = {0, 0};
    float* C = a.C[net];
    const float* bias = a.bias[net];
    const float* resid = a.resid[net];
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int sp = w % a.splits, t = w / a.splits, tn = t % a.tiles_n, tm = t / a.tiles_n;
      const int k0 = sp * kb_chunk, k1 = min(a.KB, k0 + kb_chunk);
      if (k0 >= k1) continue;
      const int tb = it & 1;
      it++;
      mbar_wait(&acc_full[tb], ph_full[tb]);
      ph_full[tb] ^= 1;
      tc_fence_after();
      const int64_t grow = (int64_t)tm * 128 + row;
      const bool valid = grow < a.rows;
#pragma unroll 1
      for (int g = 0; g < bn / 32; g++) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + tb * 128 + g * 32, r);
        tmem_ld_wait();
        const int col0 = tn * bn + g * 32;
        float* crow = C + grow * a.ldc + col0;
        if (bias && mode != GEMM_TN) {  // (warp-uniform) one coalesced load of the group's 32 bias values, handed round by shuffles
          const float bl = (col0 + lane < a.cols) ? __ldg(bias + col0 + lane) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j));
        }
        if (!valid) {
          // rows beyond the matrix: nothing to store (the TMEM load above stays warp-uniform)
        } else if (mode == GEMM_TN) {
          if ((a.cols & 3) == 0 && (a.ldc & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < a.cols)
                atomicAdd(reinterpret_cast<float4*>(crow + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (col0 + j < a.cols) atomicAdd(crow + j, __uint_as_float(r[j]));
          }
        } else {
          // every load of the 32-column group is requested before the first store: inside the store loop each load waited for
          // its own L2 round trip (~4 us per 128 x 128 tile with a bias, more with a residual)
          float4 e[8];
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) e[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (resid) {
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++)
              if (col0 + 4 * j4 < a.cols) e[j4] = __ldg(reinterpret_cast<const float4*>(resid + grow * a.ldr + col0 + 4 * j4));
          }
          if (a.accumulate) {
            float4 c[8];
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++)
              c[j4] = (col0 + 4 * j4 < a.cols) ? *reinterpret_cast<const float4*>(crow + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) e[j4].x += c[j4].x, e[j4].y += c[j4].y, e[j4].z += c[j4].z, e[j4].w += c[j4].w;
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            if (col0 + 4 * j4 >= a.cols) continue;
            const int j = 4 * j4;
            *reinterpret_cast<float4*>(crow + j) = make_float4(__uint_as_float(r[j]) + e[j4].x, __uint_as_float(r[j + 1]) + e[j4].y,
                                                               __uint_as_float(r[j + 2]) + e[j4].z, __uint_as_float(r[j + 3]) + e[j4].w);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[tb]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

static int launch_gemm(const tw_flow_config* c, GemmArgs& a, cudaStream_t st) {
  static DeviceOnce attr_done;
  const int smem = kGemmStages * kGemmStageBytes + 256 + 1024;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_gemm_img<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TW_CUDA(cudaFuncSetAttribute(k_gemm_img<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.mark();
  }
  TW_CHECK_ARG(a.bn == 128 || a.bn == 64, "gemm: bn must be 64 or 128");
  TW_CHECK_ARG(a.mode != GEMM_NN_HEADED || (a.tiles_n == 1 && a.bn == 128), "gemm: headed mode has one column tile");
  if (a.mode != GEMM_TN) a.splits = 1;
  if (a.splits < 1) a.splits = 1;
  if (a.splits > a.KB) a.splits = a.KB > 0 ? a.KB : 1;
  const int64_t n_work = (int64_t)a.tiles_m * a.tiles_n * a.splits;
  if (n_work == 0 || a.KB == 0) return TW_OK;
  const int nets = a.nets ? a.nets : 2;
  const int per = 148 / nets;
  dim3 grid((unsigned)(n_work < per ? n_work : per), nets);
  if (c->precision == TW_PRECISION_BF16X3)
    k_gemm_img<3><<<grid, 192, smem, st>>>(a);
  else
    k_gemm_img<1><<<grid, 192, smem, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ============================================================================================
// FFN backward, first half, in ONE launch per encoder layer (custom_attention_encoder.py:102-113 differentiated):
//   pre  = y1 W1^T + b1   (NT, re-computed)            dhid = dr W2   (NN)
//   act  = relu(pre)  -> operand image (for dW2 += dr^T act)
//   dpre = dhid where pre > 0, else 0  -> operand image (for dW1 += dpre^T y1 and dy1 = dr + dpre W1),  db1 += column sums of dpre
// Both products of an output tile [128 tokens x 128 hidden] accumulate side by side in TMEM (2 x 256 columns, double
// buffered), so neither [M, F] fp32 matrix is ever written: the epilogue reads the pair and stores the two images with
// 16-byte swizzled-chunk stores.  Same roles as k_gemm_img: warp 0 streams tiles, warp 1 issues, warps 2-5 drain.
struct FfnBwdPreArgs {
  ImgRef Ay[2], Ad[2];  // y1 and dr images (plain, 2 column tiles)
  ImgRef W1[2], W2[2];
  const float* b1[2];
  uint8_t* img_act[2];   // plain images, n_ct column tiles
  uint8_t* img_dpre[2];
  float* db1[2];
  int rows, tiles_m, tiles_n, n_ct;
};

// 256-bit global store of two 16-byte chunks (sm_100: st.global.v8)
__device__ __forceinline__ void st_global_v8(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ uint4 sel4(bool c, const uint4& a, const uint4& b) {
  return make_uint4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w);
}

// column sums of a [32 rows (lanes) x 32 columns (registers)] block: lane j returns the sum of column j (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; i++) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Token-tile stationary: the y1 and dr tiles of a token tile (128 KB with the lo parts) stay in shared memory while the CTA
// walks over the hidden chunks, only the weight tiles stream through the ring -- 144 KB instead of 256 KB of L2 reads per
// output tile.  Every CTA takes a contiguous range of (token tile, hidden chunk) items, so it re-loads the resident tiles at
// most twice.
constexpr int kFbpStages = 3;
constexpr int kFbpStageBytes = 32768;   // W hi 16K | W lo 16K
constexpr int kFbpResident = 131072;    // y1 kb0, kb1, dr kb0, kb1: each hi 16K | lo 16K
constexpr int kFbpThreads = 576;        // producer, MMA issuer, 16 epilogue warps (four per TMEM lane quarter: one 32-column group each)
template <int kSplit>
__global__ void __launch_bounds__(kFbpThreads, 1) k_ffn_bwd_pre(FfnBwdPreArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* resident = smem;
  uint8_t* ring = smem + kFbpResident;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kFbpStages * kFbpStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kFbpStages;
  uint64_t* acc_full = empty + kFbpStages;  // [2]
  uint64_t* acc_free = acc_full + 2;        // [2]
  uint64_t* res_full = acc_free + 2;        // the resident tiles have landed
  uint64_t* res_free = res_full + 1;        // every MMA that reads them has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_free + 1);
  if (tid == 0) {
    for (int i = 0; i < kFbpStages; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; i++) mbar_init(&acc_full[i], 1), mbar_init(&acc_free[i], 512);
    mbar_init(res_full, 1), mbar_init(res_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_work = a.tiles_m * a.tiles_n;
  const int w0 = (int)((int64_t)blockIdx.x * n_work / gridDim.x), w1 = (int)((int64_t)(blockIdx.x + 1) * n_work / gridDim.x);
  const uint32_t parts = kSplit == 3 ? 2u : 1u;

  if (warp == 0) {
    uint32_t stage = 0, phase = 0, ph_res = 0;
    int cur_tm = -1;
    const ImgRef Ay = a.Ay[net], Ad = a.Ad[net], W1 = a.W1[net], W2 = a.W2[net];
    for (int w = w0; w < w1; w++) {
      const int tn = w % a.tiles_n, tm = w / a.tiles_n;
      if (tm != cur_tm) {
        if (cur_tm >= 0) {
          mbar_wait(res_free, ph_res);
          ph_res ^= 1;
        }
        cur_tm = tm;
        if (elect_one()) {
          mbar_arrive_expect_tx(res_full, parts * 4u * 16384u);
          for (int u = 0; u < 4; u++) {
            const ImgRef& A = u < 2 ? Ay : Ad;
            const uint8_t* src = img_tile(A, tm, u & 1);
            bulk_g2s(resident + u * 32768, src, 16384, res_full);
            if (kSplit == 3) bulk_g2s(resident + u * 32768 + 16384, src + A.lo, 16384, res_full);
          }
        }
        __syncwarp();
      }
      for (int u = 0; u < 4; u++) {  // u = 2 * product + K block
        const int kb = u & 1;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = ring + stage * kFbpStageBytes;
          mbar_arrive_expect_tx(&full[stage], parts * 16384u);
          if (u < 2) {
            const uint8_t* wsrc = img_tile(W1, tn, kb);
            bulk_g2s(dst, wsrc, 16384, &full[stage]);
            if (kSplit == 3) bulk_g2s(dst + 16384, wsrc + W1.lo, 16384, &full[stage]);
          } else {
            for (int c = 0; c < 2; c++) {
              const uint8_t* wsrc = img_tile(W2, 0, 2 * tn + c) + kb * 8192;
              bulk_g2s(dst + c * 8192, wsrc, 8192, &full[stage]);
              if (kSplit == 3) bulk_g2s(dst + 16384 + c * 8192, wsrc + W2.lo, 8192, &full[stage]);
            }
          }
        }
        __syncwarp();
        if (++stage == kFbpStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, ph_free[2] = {0, 0}, ph_res = 0;
    const uint32_t idesc_nt = make_idesc_bf16(128, 128, 0, 0), idesc_nn = make_idesc_bf16(128, 128, 0, 1);
    const uint32_t res_base = smem_u32(resident);
    int it = 0, cur_tm = -1;
    for (int w = w0; w < w1; w++) {
      const int tm = w / a.tiles_n;
      if (tm != cur_tm) {
        mbar_wait(res_full, ph_res);
        ph_res ^= 1;
        cur_tm = tm;
      }
      const bool last_of_tm = (w + 1 == w1) || ((w + 1) / a.tiles_n != tm);
      const int tb = it & 1;
      if (it >= 2) {
        mbar_wait(&acc_free[tb], ph_free[tb]);
        ph_free[tb] ^= 1;
      }
      it++;
      tc_fence_after();
      for (int u = 0; u < 4; u++) {
        const uint32_t d = tmem + tb * 256 + (u >> 1) * 128;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ahi = res_base + u * 32768, alo = ahi + 16384;
          const uint32_t bhi = smem_u32(ring + stage * kFbpStageBytes), blo = bhi + 16384;
#pragma unroll
          for (int term = 0; term < (kSplit == 3 ? 3 : 1); term++) {
            const uint32_t ab = (term == 1) ? alo : ahi, bb = (term == 2) ? blo : bhi;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint64_t ad = desc_kmajor_sw128(ab + k * 32);
              const uint64_t bd = u >= 2 ? make_smem_desc(bb + k * 2048, 8192, 1024, LAYOUT_SW128) : desc_kmajor_sw128(bb + k * 32);
              mma_ss(d, ad, bd, u >= 2 ? idesc_nn : idesc_nt, ((u & 1) || term > 0 || k > 0) ? 1 : 0);
            }
          }
          mma_commit(&empty[stage]);
          if (u == 3) {
            mma_commit(&acc_full[tb]);
            if (last_of_tm) mma_commit(res_free);
          }
        }
        __syncwarp();
        if (++stage == kFbpStages) stage = 0, phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)row & 7u;
    uint32_t ph_full[2] = {0, 0};
    const float* b1 = a.b1[net];
    float* db1 = a.db1[net];
    int it = 0;
    for (int w = w0; w < w1; w++) {
      const int tn = w % a.tiles_n, tm = w / a.tiles_n;
      const int tb = it & 1;
      it++;
      mbar_wait(&acc_full[tb], ph_full[tb]);
      ph_full[tb] ^= 1;
      tc_fence_after();
      const bool valid = (int64_t)tm * 128 + row < a.rows;
      {
        const int col0 = tn * 128 + g * 32;
        const size_t toff = ((size_t)tm * a.n_ct + (size_t)(col0 >> 6)) * 32768 + (size_t)row * 128;
        uint8_t* act_hi = a.img_act[net] + toff;
        uint8_t* dp_hi = a.img_dpre[net] + toff;
        const uint32_t chunk0 = (uint32_t)(col0 & 63) >> 3;
        const float bl = __ldg(b1 + col0 + lane);
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + tb * 256 + g * 32, r);
        tmem_ld_wait();
        uint32_t keep = 0;
        // two 16-byte chunks (8 columns each) of a swizzled 128-byte row form an aligned 32-byte pair: one 256-bit store
#pragma unroll
        for (int u = 0; u < 2; u++) {
          uint4 h[2], l[2];
#pragma unroll
          for (int c = 0; c < 2; c++) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int e = u * 16 + c * 8 + j;
              const float p = __uint_as_float(r[e]) + __shfl_sync(0xffffffffu, bl, e);
              const bool on = valid && p > 0.f;
              keep |= on ? (1u << e) : 0u;
              v[j] = on ? p : 0.f;
            }
            split2(v[0], v[1], h[c].x, l[c].x), split2(v[2], v[3], h[c].y, l[c].y), split2(v[4], v[5], h[c].z, l[c].z), split2(v[6], v[7], h[c].w, l[c].w);
          }
          const uint32_t off = (((chunk0 + 2 * u) ^ sw) & ~1u) << 4;
          const bool odd = (sw & 1u) != 0;  // odd rows: the even chunk sits in the upper half of the pair
          st_global_v8(act_hi + off, sel4(odd, h[1], h[0]), sel4(odd, h[0], h[1]));
          if (kSplit == 3) st_global_v8(act_hi + 16384 + off, sel4(odd, l[1], l[0]), sel4(odd, l[0], l[1]));
        }
        tmem_ld32(tmem + lane_base + tb * 256 + 128 + g * 32, r);
        tmem_ld_wait();
        float dv[32];
#pragma unroll
        for (int j = 0; j < 32; j++) dv[j] = ((keep >> j) & 1u) ? __uint_as_float(r[j]) : 0.f;
#pragma unroll
        for (int u = 0; u < 2; u++) {
          uint4 h[2], l[2];
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int e = u * 16 + c * 8;
            split2(dv[e], dv[e + 1], h[c].x, l[c].x), split2(dv[e + 2], dv[e + 3], h[c].y, l[c].y);
            split2(dv[e + 4], dv[e + 5], h[c].z, l[c].z), split2(dv[e + 6], dv[e + 7], h[c].w, l[c].w);
          }
          const uint32_t off = (((chunk0 + 2 * u) ^ sw) & ~1u) << 4;
          const bool odd = (sw & 1u) != 0;
          st_global_v8(dp_hi + off, sel4(odd, h[1], h[0]), sel4(odd, h[0], h[1]));
          if (kSplit == 3) st_global_v8(dp_hi + 16384 + off, sel4(odd, l[1], l[0]), sel4(odd, l[0], l[1]));
        }
        const float cs = warp_colsum32(dv, lane);
        atomicAdd(db1 + col0 + lane, cs);
      }
      tc_fence_before();
      mbar_arrive(&acc_free[tb]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static int launch_ffn_bwd_pre(const tw_flow_config* c, FfnBwdPreArgs& a, cudaStream_t st) {
  static DeviceOnce attr_done;
  const int smem = kFbpResident + kFbpStages * kFbpStageBytes + 256 + 1024;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_ffn_bwd_pre<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TW_CUDA(cudaFuncSetAttribute(k_ffn_bwd_pre<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done.mark();
  }
  const int n_work = a.tiles_m * a.tiles_n;
  if (n_work == 0) return TW_OK;
  dim3 grid((unsigned)(n_work < 74 ? n_work : 74), 2);
  if (c->precision == TW_PRECISION_BF16X3)
    k_ffn_bwd_pre<3><<<grid, kFbpThreads, smem, st>>>(a);
  else
    k_ffn_bwd_pre<1><<<grid, kFbpThreads, smem, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ============================================================================================
// tile-wise element kernels.  One CTA = one [128 rows x 64 cols] image tile, 256 threads; thread e covers
// row (e >> 5) + 8*i, columns 2*(e & 31), +1  -> 4-byte stores that fill 128-byte swizzled rows.
__device__ __forceinline__ float silu_fwd(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ float silu_grad(float v) {
  const float s = 1.f / (1.f + __expf(-v));
  return s * (1.f + v * (1.f - s));
}

__device__ __forceinline__ void img_store2(uint8_t* tile_hi, uint32_t lo_off, int r, int kp, float v0, float v1) {
  uint32_t h, l;
  split2(v0, v1, h, l);
  const uint32_t off = sw128_offset(r, kp, 128);
  *reinterpret_cast<uint32_t*>(tile_hi + off) = h;
  *reinterpret_cast<uint32_t*>(tile_hi + lo_off + off) = l;
}

// column sums of per-thread pairs (s0, s1) over the 8 warps of the CTA -> atomicAdd into dst[col0 + 2*lane (+1)]
__device__ __forceinline__ void tile_colsum_add(float s0, float s1, float* red /*[8][64]*/, float* dst, int col0, int n_cols) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  red[w * 64 + 2 * lane] = s0;
  red[w * 64 + 2 * lane + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i * 64 + threadIdx.x];
    if (col0 + (int)threadIdx.x < n_cols) atomicAdd(dst + col0 + threadIdx.x, t);
  }
  __syncthreads();
}

struct PackArgs {
  const float* X[2];
  uint8_t* img[2];
  float* colsum[2];  // optional: += column sums of X
  int64_t M;
  int C, ld, n_ct;
};
__global__ void __launch_bounds__(256) k_pack_act(PackArgs a) {
  // one 16-byte chunk (8 columns) per thread and pass: eight neighbouring threads read 256 contiguous bytes and fill one
  // 128-byte swizzled row of each part (C is a multiple of 64, rows are 16-byte aligned)
  __shared__ float red[8 * 64];
  const int net = blockIdx.z, tc = blockIdx.x, tr = blockIdx.y;
  const float* X = a.X[net];
  uint8_t* hi = a.img[net] + ((size_t)tr * a.n_ct + tc) * 32768;
  const int ch = threadIdx.x & 7, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = threadIdx.x >> 3; r < 128; r += 32) {
    const int64_t gr = (int64_t)tr * 128 + r;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (gr < a.M) {
      const float4* src = reinterpret_cast<const float4*>(X + gr * a.ld + tc * 64 + ch * 8);
      v0 = src[0], v1 = src[1];
    }
    s[0] += v0.x, s[1] += v0.y, s[2] += v0.z, s[3] += v0.w, s[4] += v1.x, s[5] += v1.y, s[6] += v1.z, s[7] += v1.w;
    uint4 h, l;
    split2(v0.x, v0.y, h.x, l.x), split2(v0.z, v0.w, h.y, l.y), split2(v1.x, v1.y, h.z, l.z), split2(v1.z, v1.w, h.w, l.w);
    const uint32_t off = sw128_offset(r, ch * 8, 128);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(hi + 16384 + off) = l;
  }
  if (a.colsum[net]) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; j++) red[w * 64 + lane * 8 + j] = s[j];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) t += red[i * 64 + threadIdx.x];
      if (tc * 64 + (int)threadIdx.x < a.C) atomicAdd(a.colsum[net] + tc * 64 + threadIdx.x, t);
    }
  }
}

// Backward through an activation: given the pre-activation `pre` (bias included) and the gradient w.r.t.
// the activation output, write act(pre) and d_pre = d_act * act'(pre) as operand images and add the
// column sums of d_pre into the bias gradient.
struct ActBwdArgs {
  const float* pre[2];
  const float* dact[2];
  uint8_t* img_act[2];
  uint8_t* img_dpre[2];
  float* dbias[2];
  int64_t M;
  int C, n_ct;
};
template <int kAct>
__global__ void __launch_bounds__(256) k_act_bwd(ActBwdArgs a) {
  // one 16-byte chunk (8 columns) per thread and pass, like k_pack_act
  __shared__ float red[8 * 64];
  const int net = blockIdx.z, tc = blockIdx.x, tr = blockIdx.y;
  const size_t toff = ((size_t)tr * a.n_ct + tc) * 32768;
  uint8_t* act_hi = a.img_act[net] + toff;
  uint8_t* dp_hi = a.img_dpre[net] + toff;
  const int ch = threadIdx.x & 7, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = threadIdx.x >> 3; r < 128; r += 32) {
    const int64_t gr = (int64_t)tr * 128 + r;
    float p[8], g[8], av[8], dv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) p[j] = g[j] = 0.f;
    if (gr < a.M) {
      const float4* ps = reinterpret_cast<const float4*>(a.pre[net] + gr * a.C + tc * 64 + ch * 8);
      const float4* gs = reinterpret_cast<const float4*>(a.dact[net] + gr * a.C + tc * 64 + ch * 8);
      const float4 p0 = ps[0], p1 = ps[1], g0 = gs[0], g1 = gs[1];
      p[0] = p0.x, p[1] = p0.y, p[2] = p0.z, p[3] = p0.w, p[4] = p1.x, p[5] = p1.y, p[6] = p1.z, p[7] = p1.w;
      g[0] = g0.x, g[1] = g0.y, g[2] = g0.z, g[3] = g0.w, g[4] = g1.x, g[5] = g1.y, g[6] = g1.z, g[7] = g1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (gr >= a.M) {
        av[j] = dv[j] = 0.f;
      } else if (kAct == ACT_RELU) {
        av[j] = fmaxf(p[j], 0.f), dv[j] = p[j] > 0.f ? g[j] : 0.f;
      } else {
        av[j] = silu_fwd(p[j]), dv[j] = g[j] * silu_grad(p[j]);
      }
      s[j] += dv[j];
    }
    uint4 h, l;
    const uint32_t off = sw128_offset(r, ch * 8, 128);
    split2(av[0], av[1], h.x, l.x), split2(av[2], av[3], h.y, l.y), split2(av[4], av[5], h.z, l.z), split2(av[6], av[7], h.w, l.w);
    *reinterpret_cast<uint4*>(act_hi + off) = h;
    *reinterpret_cast<uint4*>(act_hi + 16384 + off) = l;
    split2(dv[0], dv[1], h.x, l.x), split2(dv[2], dv[3], h.y, l.y), split2(dv[4], dv[5], h.z, l.z), split2(dv[6], dv[7], h.w, l.w);
    *reinterpret_cast<uint4*>(dp_hi + off) = h;
    *reinterpret_cast<uint4*>(dp_hi + 16384 + off) = l;
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; j++) red[w * 64 + lane * 8 + j] = s[j];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i * 64 + threadIdx.x];
    if (tc * 64 + (int)threadIdx.x < a.C) atomicAdd(a.dbias[net] + tc * 64 + threadIdx.x, t);
  }
}

// Last layer of the out_mlp (hidden -> 3, mlp.py:18-23) backward, fused with the SiLU backward of the hidden
// layer: pre [M,hid] is the re-computed pre-activation, dst [M,3] the gradient of the network output.
struct OutLastArgs {
  const float* pre[2];
  const float* dst[2];
  const float* w4[2];   // [3,hid]
  uint8_t* img_dpre[2];
  float* dw4[2];        // [3,hid]
  float* db4[2];        // [3]
  float* db3[2];        // [hid]
  int64_t M;
  int hid, n_ct;
};
__global__ void __launch_bounds__(256) k_out_last_bwd(OutLastArgs a) {
  __shared__ float red[8 * 64];
  __shared__ float sd[128 * 3];
  const int net = blockIdx.z, tc = blockIdx.x, tr = blockIdx.y;
  uint8_t* dp_hi = a.img_dpre[net] + ((size_t)tr * a.n_ct + tc) * 32768;
  const int lane = threadIdx.x & 31, kp = lane * 2, gc = tc * 64 + kp;
  for (int i = threadIdx.x; i < 128 * 3; i += 256) {
    const int64_t gr = (int64_t)tr * 128 + i / 3;
    sd[i] = gr < a.M ? a.dst[net][gr * 3 + i % 3] : 0.f;
  }
  __syncthreads();
  float w[3][2];
#pragma unroll
  for (int i = 0; i < 3; i++) w[i][0] = a.w4[net][i * a.hid + gc], w[i][1] = a.w4[net][i * a.hid + gc + 1];
  float s0 = 0.f, s1 = 0.f, gw[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  float2 pv[16];  // the thread's 16 pre-activation pairs, requested at once
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int64_t gr = (int64_t)tr * 128 + (threadIdx.x >> 5) + 8 * i;
    pv[i] = gr < a.M ? __ldg(reinterpret_cast<const float2*>(a.pre[net] + gr * a.hid + gc)) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int r = (threadIdx.x >> 5) + 8 * i;
    const int64_t gr = (int64_t)tr * 128 + r;
    float d0 = 0.f, d1 = 0.f;
    if (gr < a.M) {
      const float2 p = pv[i];
      const float g0 = sd[r * 3], g1 = sd[r * 3 + 1], g2 = sd[r * 3 + 2];
      const float act0 = silu_fwd(p.x), act1 = silu_fwd(p.y);
      gw[0][0] = fmaf(g0, act0, gw[0][0]), gw[0][1] = fmaf(g0, act1, gw[0][1]);
      gw[1][0] = fmaf(g1, act0, gw[1][0]), gw[1][1] = fmaf(g1, act1, gw[1][1]);
      gw[2][0] = fmaf(g2, act0, gw[2][0]), gw[2][1] = fmaf(g2, act1, gw[2][1]);
      d0 = (g0 * w[0][0] + g1 * w[1][0] + g2 * w[2][0]) * silu_grad(p.x);
      d1 = (g0 * w[0][1] + g1 * w[1][1] + g2 * w[2][1]) * silu_grad(p.y);
    }
    s0 += d0, s1 += d1;
    img_store2(dp_hi, 16384, r, kp, d0, d1);
  }
  tile_colsum_add(s0, s1, red, a.db3[net], tc * 64, a.hid);
#pragma unroll
  for (int i = 0; i < 3; i++) tile_colsum_add(gw[i][0], gw[i][1], red, a.dw4[net] + i * a.hid, tc * 64, a.hid);
  if (tc == 0 && threadIdx.x < 3) {
    float t = 0.f;
    for (int r = 0; r < 128; r++) t += sd[r * 3 + threadIdx.x];
    atomicAdd(a.db4[net] + threadIdx.x, t);
  }
}

// LayerNorm backward over 128 features (custom_attention_encoder.py:110,113): one CTA per 128-row tile, one
// warp per row (lane l owns columns 4l..4l+3).  dr = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// Writes dr (fp32 + operand image) and adds dgamma, dbeta and (optionally) the column sums of dr.
struct LnBwdArgs {
  const float* pre[2];   // [M,128] LayerNorm input
  const float* dy[2];    // [M,128]
  const float* gamma[2];
  float* dr[2];          // [M,128]
  uint8_t* img_dr[2];    // plain image, 2 column tiles
  float* dgamma[2];
  float* dbeta[2];
  float* dbias[2];       // optional
  int64_t M;
  float eps;
};
__global__ void __launch_bounds__(256) k_ln_bwd(LnBwdArgs a) {
  __shared__ float red[3][8][128];
  // one CTA per 32 rows of a tile (four rows per warp, all requested at once): 44 token tiles alone leave most SMs idle
  const int net = blockIdx.y, tr = blockIdx.x >> 2, rbase = (blockIdx.x & 3) * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float4 gm = *reinterpret_cast<const float4*>(a.gamma[net] + 4 * lane);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, as = ag;
  uint8_t* tile_hi = a.img_dr[net] + ((size_t)tr * 2 + (lane >> 4)) * 32768;
  {
   const int rb = rbase + w;
   float4 xs[4], dys[4];
#pragma unroll
   for (int u = 0; u < 4; u++) {
     const int64_t gr = (int64_t)tr * 128 + rb + 8 * u;
     xs[u] = dys[u] = make_float4(0.f, 0.f, 0.f, 0.f);
     if (gr < a.M) {
       xs[u] = __ldg(reinterpret_cast<const float4*>(a.pre[net] + gr * 128 + 4 * lane));
       dys[u] = __ldg(reinterpret_cast<const float4*>(a.dy[net] + gr * 128 + 4 * lane));
     }
   }
#pragma unroll
   for (int u = 0; u < 4; u++) {
    const int r = rb + 8 * u;
    const int64_t gr = (int64_t)tr * 128 + r;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < a.M) {
      const float4 x = xs[u];
      const float4 dy = dys[u];
      const float mean = warp_sum((x.x + x.y) + (x.z + x.w)) * (1.f / 128.f);
      const float4 xc = make_float4(x.x - mean, x.y - mean, x.z - mean, x.w - mean);
      const float var = warp_sum(xc.x * xc.x + xc.y * xc.y + xc.z * xc.z + xc.w * xc.w) * (1.f / 128.f);
      const float rstd = 1.0f / sqrtf(var + a.eps);
      const float4 xh = make_float4(xc.x * rstd, xc.y * rstd, xc.z * rstd, xc.w * rstd);
      const float4 g = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
      const float mg = warp_sum((g.x + g.y) + (g.z + g.w)) * (1.f / 128.f);
      const float mgx = warp_sum(g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w) * (1.f / 128.f);
      d = make_float4(rstd * (g.x - mg - xh.x * mgx), rstd * (g.y - mg - xh.y * mgx), rstd * (g.z - mg - xh.z * mgx),
                      rstd * (g.w - mg - xh.w * mgx));
      *reinterpret_cast<float4*>(a.dr[net] + gr * 128 + 4 * lane) = d;
      ag.x = fmaf(dy.x, xh.x, ag.x), ag.y = fmaf(dy.y, xh.y, ag.y), ag.z = fmaf(dy.z, xh.z, ag.z), ag.w = fmaf(dy.w, xh.w, ag.w);
      ab.x += dy.x, ab.y += dy.y, ab.z += dy.z, ab.w += dy.w;
      as.x += d.x, as.y += d.y, as.z += d.z, as.w += d.w;
    }
    uint32_t h0, l0, h1, l1;
    split2(d.x, d.y, h0, l0);
    split2(d.z, d.w, h1, l1);
    const uint32_t off = sw128_offset(r, (4 * lane) & 63, 128);
    *reinterpret_cast<uint2*>(tile_hi + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(tile_hi + 16384 + off) = make_uint2(l0, l1);
   }
  }
  *reinterpret_cast<float4*>(&red[0][w][4 * lane]) = ag;
  *reinterpret_cast<float4*>(&red[1][w][4 * lane]) = ab;
  *reinterpret_cast<float4*>(&red[2][w][4 * lane]) = as;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 128; i += 256) {
    const int which = i >> 7, col = i & 127;
    float* dst = which == 0 ? a.dgamma[net] : (which == 1 ? a.dbeta[net] : a.dbias[net]);
    if (!dst) continue;
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) t += red[which][j][col];
    atomicAdd(dst + col, t);
  }
}

// Conditioner input of both networks as an operand image [M, 64] (one column tile): the feature gather of
// flow.py:172 / custom_transformer_nvp.py:64-71, zero padded from E+9 to 64 columns.
__global__ void __launch_bounds__(256) k_features_img(const float* __restrict__ embed, const int64_t* __restrict__ atom_types,
                                                      const float* __restrict__ xc, const float* __restrict__ xv,
                                                      const float* __restrict__ z_other, int64_t M, int V, int E, int n_types,
                                                      uint8_t* __restrict__ img) {
  // one CTA per 32 rows of a tile: a launch over the 44 token tiles of a training batch is pure load latency
  const int tr = blockIdx.x >> 2, rbase = (blockIdx.x & 3) * 32;
  uint8_t* hi = img + (size_t)tr * 32768;
  const int lane = threadIdx.x & 31, kp = lane * 2;
  for (int r = rbase + (threadIdx.x >> 5); r < rbase + 32; r += 8) {
    const int64_t m = (int64_t)tr * 128 + r;
    float v[2] = {0.f, 0.f};
    if (m < M) {
      int64_t t = atom_types[m];
      t = t < 0 ? 0 : (t >= n_types ? n_types - 1 : t);
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int e = kp + u;
        if (e < E) v[u] = embed[t * E + e];
        else if (e < E + 3) v[u] = xc[m * 3 + (e - E)];
        else if (e < E + 6) v[u] = xv[m * 3 + (e - E - 3)];
        else if (e < E + 9) v[u] = z_other[m * 3 + (e - E - 6)];
      }
    }
    img_store2(hi, 16384, r, kp, v[0], v[1]);
  }
}

// Gradient of the conditioner input [M,64] (both networks) -> flow state (columns E+6..E+8) and atom embedding
__global__ void __launch_bounds__(256) k_du_scatter(const float* __restrict__ du0, const float* __restrict__ du1,
                                                    const int64_t* __restrict__ atom_types, int64_t M, int E, int n_types,
                                                    float* __restrict__ dz_other, float* __restrict__ dembed,
                                                    float* __restrict__ dxc, float* __restrict__ dxv) {
  extern __shared__ float acc[];  // [n_types * E]
  for (int i = threadIdx.x; i < n_types * E; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int64_t m0 = (int64_t)blockIdx.x * 32;  // 32 rows per CTA (see k_features_img)
  for (int idx = threadIdx.x; idx < 32 * 64; idx += blockDim.x) {
    const int64_t m = m0 + (idx >> 6);
    const int e = idx & 63;
    if (m >= M) continue;
    const float g = du0[m * 64 + e] + du1[m * 64 + e];
    if (e < E) {
      int64_t t = atom_types[m];
      t = t < 0 ? 0 : (t >= n_types ? n_types - 1 : t);
      atomicAdd(&acc[t * E + e], g);
    } else if (e >= E + 6 && e < E + 9) {
      dz_other[m * 3 + (e - E - 6)] += g;
    } else if (dxc && e < E + 3) {  // conditioning coordinates / velocities as conditioner inputs (AcceptanceLoss)
      dxc[m * 3 + (e - E)] += g;
    } else if (dxv && e >= E + 3 && e < E + 6) {
      dxv[m * 3 + (e - E - 3)] += g;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_types * E; i += blockDim.x)
    if (acc[i] != 0.f) atomicAdd(dembed + i, acc[i]);
}

// Affine coupling backward (nvp.py:127-133): z' = z * exp(s) + t, log-det = sum s over unmasked atoms.
// In: dz (gradient w.r.t. z'), z (state BEFORE the layer), s, g[n] = d(objective)/d(log p).  Out: dz <- dz * exp(s),
// ds = dz' * z * exp(s) + g * keep, dt = dz'.
__global__ void __launch_bounds__(128) k_coupling_bwd(const float* __restrict__ s, const float* __restrict__ z_in,
                                                      float* __restrict__ dz, const uint8_t* __restrict__ mask,
                                                      const float* __restrict__ g, int V, float* __restrict__ ds,
                                                      float* __restrict__ dt) {
  const int64_t n = blockIdx.x;
  const float gn = g[n];
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) {
    const int64_t i = n * V * 3 + e;
    const float sc = expf(s[i]), d = dz[i];
    ds[i] = d * z_in[i] * sc + (mask[n * V + e / 3] ? 0.f : gn);
    dt[i] = d;
    dz[i] = d * sc;
  }
}

// Sampling direction (nvp.py:137-183): z_out = (z_in - t) e^{-s}, delta += sum s.  Given d/dz_out in dz and g = d/d(delta):
__global__ void __launch_bounds__(128) k_coupling_rev_bwd(const float* __restrict__ s, const float* __restrict__ z_out,
                                                          float* __restrict__ dz, const uint8_t* __restrict__ mask,
                                                          const float* __restrict__ g, int V, float* __restrict__ ds,
                                                          float* __restrict__ dt) {
  const int64_t n = blockIdx.x;
  const float gn = g[n];
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) {
    const int64_t i = n * V * 3 + e;
    const float isc = expf(-s[i]), d = dz[i];
    ds[i] = -d * z_out[i] + (mask[n * V + e / 3] ? 0.f : gn);
    dt[i] = -d * isc;
    dz[i] = d * isc;
  }
}

// Prior backward (flow.py:159-166,191-203): lp = sum keep * (-z^2 / (2 e^{2 sigma}) - sigma - const)
__global__ void __launch_bounds__(128) k_prior_bwd(const float* __restrict__ zc, const float* __restrict__ zv,
                                                   const uint8_t* __restrict__ mask, const float* __restrict__ lsc,
                                                   const float* __restrict__ lsv, const float* __restrict__ g, int V,
                                                   float* __restrict__ dzc, float* __restrict__ dzv, float* __restrict__ dlsc,
                                                   float* __restrict__ dlsv) {
  __shared__ float red[33];
  const int64_t n = blockIdx.x;
  const float gn = g[n];
  const float ivc = expf(-2.f * lsc[0]), ivv = expf(-2.f * lsv[0]);
  float ac = 0.f, av = 0.f;
  for (int e = threadIdx.x; e < V * 3; e += blockDim.x) {
    const int64_t i = n * V * 3 + e;
    const bool keep = !mask[n * V + e / 3];
    const float a = zc[i], b = zv[i];
    dzc[i] = keep ? -gn * a * ivc : 0.f;
    dzv[i] = keep ? -gn * b * ivv : 0.f;
    if (keep) ac += gn * (a * a * ivc - 1.f), av += gn * (b * b * ivv - 1.f);
  }
  ac = block_sum(ac, red);
  av = block_sum(av, red);
  if (threadIdx.x == 0) {
    if (dlsc) atomicAdd(dlsc, ac);
    if (dlsv) atomicAdd(dlsv, av);
  }
}

// W_c,h = W_o,h W_v,h (k_pack_all) backward: dW_o[i, hD+k] += sum_j dWc[i, hD+j] W_v[hD+k, j];
//                                              dW_v[hD+k, j] += sum_i W_o[i, hD+k] dWc[i, hD+j].
struct WcChainArgs {
  const float* dwc[2];
  const float* wo[2];
  const float* wv[2];
  float* dwo[2];
  float* dwv[2];
};
// One CTA per (head, block of 16 rows, network, product), 256 threads.  Product 0 (dW_o) keeps the head's W_v block (padded rows)
// and its 16 rows of dWc in shared memory, product 1 (dW_v) the head's dWc block and 16 columns of W_o: 74 KB per CTA, three CTAs
// per SM, 192 CTAs in one wave (the row-per-CTA first version moved 200 MB through L2 per launch: 49 us for 50 MFLOP; both
// products in one CTA of 138 KB: 22 us).  Thread t owns column (t & 127) of 8 output rows.
constexpr int kWcChainSmem = (128 * 129 + 128 * 16) * (int)sizeof(float);
__global__ void __launch_bounds__(256) k_wc_chain(WcChainArgs a, int H) {
  extern __shared__ __align__(16) float wsm[];
  float* sBig = wsm;                 // product 0: [128 k][129] W_v[hD + k, j];  product 1: [128 i][128] dWc[i, hD + j]
  float* sSmall = wsm + 128 * 129;   // product 0: [16 i][128] dWc[16 rb + i, hD + j];  product 1: [128 i][16] W_o[i, hD + 16 rb + kk]
  const int net = blockIdx.z >> 1, prod = blockIdx.z & 1, h = blockIdx.x, rb = blockIdx.y, t = threadIdx.x;
  const int HD = H * 128;
  const float* dwc = a.dwc[net] + h * 128;
  const float* wo = a.wo[net] + h * 128;
  const float* wv = a.wv[net] + (size_t)h * 128 * 128;
  const int col = t & 127, g8 = (t >> 7) * 8;
  if (prod == 0) {
    {  // float4 pieces: row = e >> 5, columns 4 * (e & 31); all 16 loads of a thread in flight before the first store
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = t + 256 * i;
        v[i] = __ldg(reinterpret_cast<const float4*>(wv + (size_t)(e >> 5) * 128 + (e & 31) * 4));
      }
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = t + 256 * i, r = e >> 5, c4 = (e & 31) * 4;
        sBig[r * 129 + c4] = v[i].x, sBig[r * 129 + c4 + 1] = v[i].y, sBig[r * 129 + c4 + 2] = v[i].z, sBig[r * 129 + c4 + 3] = v[i].w;
      }
    }
    for (int e = t; e < 16 * 32; e += 256) {
      const int r = e >> 5, c4 = (e & 31) * 4;
      *reinterpret_cast<float4*>(sSmall + r * 128 + c4) = __ldg(reinterpret_cast<const float4*>(dwc + (size_t)(rb * 16 + r) * HD + c4));
    }
    __syncthreads();
    // dW_o[i, hD + k] += sum_j dWc[i, hD + j] W_v[hD + k, j],   i = 16 rb + g8 + u, k = col
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* drow = sSmall + g8 * 128;
#pragma unroll 4
    for (int j = 0; j < 128; j += 4) {
      const float v0 = sBig[col * 129 + j], v1 = sBig[col * 129 + j + 1], v2 = sBig[col * 129 + j + 2], v3 = sBig[col * 129 + j + 3];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const float4 d = *reinterpret_cast<const float4*>(drow + u * 128 + j);  // broadcast
        acc[u] = fmaf(d.x, v0, fmaf(d.y, v1, fmaf(d.z, v2, fmaf(d.w, v3, acc[u]))));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) a.dwo[net][(size_t)(rb * 16 + g8 + u) * HD + h * 128 + col] += acc[u];
  } else {
    {
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = t + 256 * i;
        v[i] = __ldg(reinterpret_cast<const float4*>(dwc + (size_t)(e >> 5) * HD + (e & 31) * 4));
      }
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = t + 256 * i;
        *reinterpret_cast<float4*>(sBig + (e >> 5) * 128 + (e & 31) * 4) = v[i];
      }
    }
    for (int e = t; e < 128 * 4; e += 256) {
      const int r = e >> 2, c4 = (e & 3) * 4;
      *reinterpret_cast<float4*>(sSmall + r * 16 + c4) = __ldg(reinterpret_cast<const float4*>(wo + (size_t)r * HD + rb * 16 + c4));
    }
    __syncthreads();
    // dW_v[hD + k, j] += sum_i W_o[i, hD + k] dWc[i, hD + j],   k = 16 rb + g8 + u, j = col
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < 128; i++) {
      const float d = sBig[i * 128 + col];
      const float4 o0 = *reinterpret_cast<const float4*>(sSmall + i * 16 + g8), o1 = *reinterpret_cast<const float4*>(sSmall + i * 16 + g8 + 4);
      acc[0] = fmaf(o0.x, d, acc[0]), acc[1] = fmaf(o0.y, d, acc[1]), acc[2] = fmaf(o0.z, d, acc[2]), acc[3] = fmaf(o0.w, d, acc[3]);
      acc[4] = fmaf(o1.x, d, acc[4]), acc[5] = fmaf(o1.y, d, acc[5]), acc[6] = fmaf(o1.z, d, acc[6]), acc[7] = fmaf(o1.w, d, acc[7]);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) a.dwv[net][(size_t)(h * 128 + rb * 16 + g8 + u) * 128 + col] += acc[u];
  }
}

// ============================================================================================
// Gradient w.r.t. the attention lengthscales (`learnable_kernel`, kernel_attention.py:217-253).  One score set A_h(l_h)
// serves every attention layer of a pass (the reference's cache quirk), so
//   dL/dl_h = sum_{b,i,j} S_h[b,i,j] * dA_h[b,i,j]/dl_h,    S_h[b] = sum_{layers, nets} Q_h[b] X[b]^T,   Q_h = dr W_c,h
// (r1 = x + sum_h W_c,h (A_h x)  =>  dL/dA_h[i,j] = <dr_i W_c,h, x_j>).  Q comes from one NN GEMM per layer over the W_c
// image (all heads at once); k_score_grad accumulates S; k_ls_grad applies the derivative of the normalised Gaussian scores.
struct ScoreGradArgs {
  const float* q[2];  // [M, H*128]
  const float* x[2];  // [M, 128] layer input
  float* S;           // [B, H, V, V]
};
__global__ void __launch_bounds__(128) k_score_grad(ScoreGradArgs a, int V, int H, int net0, int net1) {
  extern __shared__ float sg[];
  float* sX = sg;                       // [V][129]
  float* sQ = sX + (size_t)V * 129;     // [4 warps][128]
  const int64_t b = blockIdx.x;
  const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* S = a.S + ((size_t)b * H + h) * V * V;
  for (int net = net0; net < net1; net++) {  // [0, 2): both networks share the scores; chebyshev_kernel: one at a time
    const float* X = a.x[net] + b * V * 128;
    for (int e = threadIdx.x; e < V * 128; e += blockDim.x) sX[(e >> 7) * 129 + (e & 127)] = X[e];
    __syncthreads();
    for (int i = warp; i < V; i += 4) {
      const float* qrow = a.q[net] + ((size_t)b * V + i) * (size_t)(H * 128) + h * 128;
      for (int f = lane; f < 128; f += 32) sQ[warp * 128 + f] = qrow[f];
      __syncwarp();
      for (int j = lane; j < V; j += 32) {
        float acc = 0.f;
#pragma unroll 8
        for (int f = 0; f < 128; f++) acc = fmaf(sQ[warp * 128 + f], sX[j * 129 + f], acc);
        S[(size_t)i * V + j] += acc;
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// A_ij = K_ij / (s_i + eps), K_ij = exp(-(d_ij / l)^2) [j not padding], s_i = sum_j K_ij (kernel_attention.py:105-119):
//   dA_ij/dl = K'_ij / (s_i + eps) - K_ij s'_i / (s_i + eps)^2,   K'_ij = K_ij * 2 d_ij^2 / l^3,   s'_i = sum_j K'_ij
__global__ void __launch_bounds__(256) k_ls_grad(const float* __restrict__ xc, const uint8_t* __restrict__ mask, const float* __restrict__ ls,
                                                 const float* __restrict__ S, int V, int H, float* __restrict__ dls) {
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = xc + b * V * 3;
  const uint8_t* mb = mask + b * V;
  __shared__ float red[8];
  for (int h = 0; h < H; h++) {
    const float l = ls[h], inv_l3 = 2.f / (l * l * l);
    const float* Sh = S + ((size_t)b * H + h) * V * V;
    float part = 0.f;
    for (int i = warp; i < V; i += 8) {
      const float xi = xb[i * 3], yi = xb[i * 3 + 1], zi = xb[i * 3 + 2];
      float s = 0.f, sp = 0.f, a1 = 0.f, a2 = 0.f;  // s, s', sum S K', sum S K
      for (int j = lane; j < V; j += 32) {
        if (mb[j]) continue;
        const float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float k = expf(-d2 / (l * l)), kp = k * d2 * inv_l3;
        const float sv = Sh[(size_t)i * V + j];
        s += k, sp += kp, a1 = fmaf(sv, kp, a1), a2 = fmaf(sv, k, a2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o), sp += __shfl_xor_sync(0xffffffffu, sp, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o), a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      const float den = s + 1e-5f;
      part += a1 / den - a2 * sp / (den * den);
    }
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; w++) t += red[w];
      atomicAdd(&dls[h], t);
    }
    __syncthreads();
  }
}

// chebyshev_kernel (kernel_attention.py:13-66,255-339): w_ij = sum_c (coef_c - mean) R_c((d_ij / l)^2) [j not padding],
// A_ij = w_ij / (sum_j |w_ij| + eps).  With S = dL/dA of ONE attention layer of ONE network:
//   dL/dw_ij = (S_ij - sign(w_ij) D_i) / s_i,  D_i = sum_j S_ij A_ij;   dL/dcoef_c = sum_ij dL/dw_ij R_c  (- the mean over c when
// the expansion is forced to vanish at infinity).  One block per state, one warp per row, per-lane partial sums per coefficient.
__global__ void __launch_bounds__(256) k_cheb_grad(const float* __restrict__ xc, const uint8_t* __restrict__ mask, const float* __restrict__ ls,
                                                   const float* __restrict__ coef_all, int order, int force_zero,
                                                   const float* __restrict__ S, int V, int H, float* __restrict__ dcoef) {
  __shared__ float acc[TW_MAX_HEADS * TW_MAX_CHEB_ORDER];
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = xc + b * V * 3;
  const uint8_t* mb = mask + b * V;
  for (int i = threadIdx.x; i < H * order; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int h = 0; h < H; h++) {
    const float l = ls[h];
    const float* coef = coef_all + (size_t)h * order;
    const float cmean = cheb_mean(coef, order, force_zero);
    const float* Sh = S + ((size_t)b * H + h) * V * V;
    float part[TW_MAX_CHEB_ORDER];
#pragma unroll
    for (int c = 0; c < TW_MAX_CHEB_ORDER; c++) part[c] = 0.f;
    for (int i = warp; i < V; i += 8) {
      const float xi = xb[i * 3], yi = xb[i * 3 + 1], zi = xb[i * 3 + 2];
      float s = 0.f, dot = 0.f;
      for (int j = lane; j < V; j += 32) {
        if (mb[j]) continue;
        const float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
        const float w = attention_basis(sqrtf(dx * dx + dy * dy + dz * dz) / l, coef, order, cmean);
        s += fabsf(w), dot = fmaf(Sh[(size_t)i * V + j], w, dot);
      }
      s = warp_sum(s) + 1e-5f, dot = warp_sum(dot) / s;  // dot = D_i
      for (int j = lane; j < V; j += 32) {
        if (mb[j]) continue;
        const float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
        const float a = sqrtf(dx * dx + dy * dy + dz * dz) / l;
        const float w = attention_basis(a, coef, order, cmean);
        const float gw = (Sh[(size_t)i * V + j] - (w > 0.f ? dot : (w < 0.f ? -dot : 0.f))) / s;
        const float y = a * a, rf = (y - 1.0f) / (y + 1.0f);
        float rprev = 1.0f, rcur = rf;
        part[0] += gw;
        if (order >= 2) part[1] = fmaf(gw, rcur, part[1]);
#pragma unroll
        for (int c = 2; c < TW_MAX_CHEB_ORDER; c++) {
          if (c < order) {
            const float rnext = 2.0f * rf * rcur - rprev;
            part[c] = fmaf(gw, rnext, part[c]);
            rprev = rcur, rcur = rnext;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < TW_MAX_CHEB_ORDER; c++) {
      if (c < order) {
        const float t = warp_sum(part[c]);
        if (lane == 0) atomicAdd(&acc[h * order + c], t);
      }
    }
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float m = 0.f;
    if (force_zero) {
      for (int c = 0; c < order; c++) m += acc[h * order + c];
      m /= (float)order;
    }
    for (int c = 0; c < order; c++) atomicAdd(&dcoef[h * order + c], acc[h * order + c] - m);
  }
}

// Gradient of the row-normalised Gaussian scores w.r.t. the (centred) conditioning coordinates: with gK_ij = dL/dK_ij =
// (S_ij - sum_j' S_ij' A_ij') / (s_i + eps) and K_ij = exp(-|x_i - x_j|^2 / l^2):  dL/dx_i += c_ij (x_i - x_j), dL/dx_j -= c_ij
// (x_i - x_j), c_ij = gK_ij K_ij (-2 / l^2).  One block per state; accumulators in shared memory.
__global__ void __launch_bounds__(256) k_score_coord_grad(const float* __restrict__ xc, const uint8_t* __restrict__ mask,
                                                          const float* __restrict__ ls, const float* __restrict__ S, int V, int H,
                                                          float* __restrict__ dxc) {
  extern __shared__ float acc[];  // [V*3]
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = xc + b * V * 3;
  const uint8_t* mb = mask + b * V;
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int h = 0; h < H; h++) {
    const float l = ls[h], inv_l2 = 1.f / (l * l);
    const float* Sh = S + ((size_t)b * H + h) * V * V;
    for (int i = warp; i < V; i += 8) {
      const float xi = xb[i * 3], yi = xb[i * 3 + 1], zi = xb[i * 3 + 2];
      float s = 0.f, dot = 0.f;
      for (int j = lane; j < V; j += 32) {
        if (mb[j]) continue;
        const float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
        const float k = expf(-(dx * dx + dy * dy + dz * dz) * inv_l2);
        s += k, dot = fmaf(Sh[(size_t)i * V + j], k, dot);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const float inv_den = 1.f / (s + 1e-5f);
      dot *= inv_den;  // sum_j S_ij A_ij
      float gx = 0.f, gy = 0.f, gz = 0.f;
      for (int j = lane; j < V; j += 32) {
        if (mb[j]) continue;
        const float dx = xi - xb[j * 3], dy = yi - xb[j * 3 + 1], dz = zi - xb[j * 3 + 2];
        const float k = expf(-(dx * dx + dy * dy + dz * dz) * inv_l2);
        const float cij = (Sh[(size_t)i * V + j] - dot) * inv_den * k * (-2.f * inv_l2);
        gx = fmaf(cij, dx, gx), gy = fmaf(cij, dy, gy), gz = fmaf(cij, dz, gz);
        atomicAdd(&acc[j * 3], -cij * dx), atomicAdd(&acc[j * 3 + 1], -cij * dy), atomicAdd(&acc[j * 3 + 2], -cij * dz);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        gx += __shfl_xor_sync(0xffffffffu, gx, o), gy += __shfl_xor_sync(0xffffffffu, gy, o), gz += __shfl_xor_sync(0xffffffffu, gz, o);
      if (lane == 0) atomicAdd(&acc[i * 3], gx), atomicAdd(&acc[i * 3 + 1], gy), atomicAdd(&acc[i * 3 + 2], gz);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) dxc[b * V * 3 + i] += acc[i];
}

// ============================================================================================
// tape layout
struct NetTape {
  float* st;            // [M,3] network output (s or t)
  float* h[8];          // h[0] = in_mlp output, h[t+1] = output of encoder layer t
  float* y1[8];         // post-LN1
  float* r1[8];         // pre-LN1
  float* r2[8];         // pre-LN2
  uint8_t* mixed[8];    // operand images of the per-head averages
};
struct Tape {
  float* xc;            // centred conditioning coordinates [B,V,3]
  float* com;
  float* scores;        // [B,H,V,V]
  uint8_t* scores_img;  // forward images
  uint8_t* scores_img_t;  // transposed images (backward)
  float* z[17][2];      // flow state (coords, velocs) before coupling layer k; [L] = final latent
  float* delta;
  NetTape net[16][2];
  // `local` attention (dot-product attention within max_radius): scratch of the taped forward, shared by every layer
  float* qkv[2];        // [M, 3 H D]
  float* att[2];        // [M, H D]
  uint8_t* img_x[2];    // image of the layer input [M,128]
  uint8_t* img_att[2];  // image of the attention output [M, H D]
  uint8_t* img_wq[2];   // image of qkv_proj.weight [3 H D, 128]
  uint8_t* img_wo[2];   // image of output_proj.weight [128, H D]
};

static size_t carve_tape(const tw_flow_config* c, int64_t B, int V, void* base, size_t cap, Tape* out) {
  Arena ar(base, cap);
  const int64_t M = B * V;
  const int L = c->num_coupling_layers, T = c->num_transformer_layers, H = c->num_heads;
  Tape t{};
  const bool local = c->attention_type == TW_ATTENTION_LOCAL;
  t.xc = ar.take<float>(M * 3);
  t.com = ar.take<float>(B * 3);
  t.scores = ar.take<float>(local ? 0 : (size_t)B * H * V * V);
  ar.off = align_up(ar.off, 1024);
  t.scores_img = ar.take<uint8_t>(local ? 0 : tc_scores_img_bytes(c, B, V));
  ar.off = align_up(ar.off, 1024);
  t.scores_img_t = ar.take<uint8_t>(local ? 0 : tc_scores_img_bytes(c, B, V));
  if (local) {
    const int HD = H * c->d_model;
    for (int s = 0; s < 2; s++) {
      t.qkv[s] = ar.take<float>(M * 3 * HD);
      t.att[s] = ar.take<float>(M * HD);
      ar.off = align_up(ar.off, 1024);
      t.img_x[s] = ar.take<uint8_t>(plain_img_bytes(M, 2));
      t.img_att[s] = ar.take<uint8_t>(plain_img_bytes(M, HD / 64));
      t.img_wq[s] = ar.take<uint8_t>(plain_img_bytes(3 * HD, 2));
      t.img_wo[s] = ar.take<uint8_t>(plain_img_bytes(128, HD / 64));
    }
  }
  t.delta = ar.take<float>(B);
  for (int k = 0; k <= L; k++)
    for (int i = 0; i < 2; i++) t.z[k][i] = ar.take<float>(M * 3);
  for (int k = 0; k < L; k++)
    for (int s = 0; s < 2; s++) {
      NetTape& n = t.net[k][s];
      n.st = ar.take<float>(M * 3);
      for (int i = 0; i <= T; i++) n.h[i] = ar.take<float>(M * 128);
      for (int i = 0; i < T; i++) {
        n.y1[i] = ar.take<float>(M * 128);
        n.r1[i] = ar.take<float>(M * 128);
        n.r2[i] = ar.take<float>(M * 128);
        ar.off = align_up(ar.off, 1024);
        n.mixed[i] = ar.take<uint8_t>(local ? 0 : tc_mixed_img_bytes(c, M));
      }
    }
  if (out) *out = t;
  return align_up(ar.off, 1024);
}

static int check_train(const tw_flow_config* c, const void* const* params, int64_t B, int64_t V) {
  if (c && c->attention_type != TW_ATTENTION_KERNEL && c->attention_type != TW_ATTENTION_CHEBYSHEV && c->attention_type != TW_ATTENTION_LOCAL)
    return fail(TW_ERR_UNSUPPORTED, "the training path is built for the `kernel` / `learnable_kernel` / `chebyshev_kernel` / `local` attention");
  if (c && c->attention_type == TW_ATTENTION_LOCAL && c->max_radius <= 0.f) return fail(TW_ERR_INVALID, "local attention needs max_radius > 0");
  TW_CHECK_ARG(c != nullptr && params != nullptr, "NULL cfg / params");
  if (c->precision == TW_PRECISION_FP32 || !tc_supported(c))
    return fail(TW_ERR_UNSUPPORTED, "the training path needs a tensor-core precision (bf16x3 / bf16) and the flagship layer sizes");
  TW_CHECK_ARG(c->num_coupling_layers <= 16 && c->num_transformer_layers <= 8, "too many layers for the training tape");
  TW_CHECK_ARG(c->num_coupling_layers % 2 == 0 && c->num_coupling_layers >= 2, "Real NVP should have an even number of coupling layers");
  TW_CHECK_ARG(B >= 0 && V >= 1 && V <= 128, "bad sizes (training path: at most 128 atoms per sample)");
  TW_CHECK_ARG(B * V < (1LL << 24), "too many tokens for one training call");
  return TW_OK;
}

// ============================================================================================
// backward workspace
struct BwdBuffers {
  float* dz[2];          // gradient w.r.t. the flow state (coords, velocs)
  float* dst[2];         // gradient w.r.t. the network outputs (s, t)  [M,3]
  float* dyA[2];         // [M,128] gradient ping
  float* dyB[2];         // [M,128] gradient pong
  float* dr[2];          // [M,128] LayerNorm-input gradient
  float* wide0[2];       // [M,F] fp32 (re-computed pre-activations)
  float* wide1[2];       // [M,F] fp32 (gradient w.r.t. the wide activation)
  uint8_t* img_x[2];     // [M,128] image of a saved activation
  uint8_t* img_d[2];     // [M,128] image of a gradient
  uint8_t* img_w0[2];    // [M,F] image (activation)
  uint8_t* img_w1[2];    // [M,F] image (pre-activation gradient)
  uint8_t* img_g[2];     // [M,H*128] image (transposed mixing of the gradient)
  uint8_t* img_u;        // [M,64] image of the conditioner input (shared by both networks)
  float* du[2];          // [M,64]
  float* dwc[2];         // [128, H*128]
  float* sgrad;          // [B,H,V,V] gradient w.r.t. the attention scores, summed over layers and networks (learnable lengthscales)
  float* dxc;            // [M,3] gradient w.r.t. the centred conditioning coordinates (conditioner inputs + scores)
  float* dxv;            // [M,3] gradient w.r.t. the conditioning velocities
  // `local` attention: recomputed q | k | v and attention output, their gradients, and the weight images packed per layer
  float* lqkv[2];        // [M, 3 H D]
  float* ldqkv[2];       // [M, 3 H D]
  float* latt[2];        // [M, H D]
  float* ldatt[2];       // [M, H D]
  uint8_t* img_dq[2];    // image of ldqkv
  uint8_t* img_wq[2];    // image of qkv_proj.weight [3 H D, 128]
  uint8_t* img_wo[2];    // image of output_proj.weight [128, H D]
};

static size_t carve_bwd(const tw_flow_config* c, int64_t B, int V, void* base, size_t cap, BwdBuffers* out) {
  Arena ar(base, cap);
  const int64_t M = B * V;
  const int F = c->dim_feedforward > 256 ? c->dim_feedforward : 256, H = c->num_heads;
  BwdBuffers b{};
  for (int i = 0; i < 2; i++) {
    b.dz[i] = ar.take<float>(M * 3);
    b.dst[i] = ar.take<float>(M * 3);
    b.dyA[i] = ar.take<float>(M * 128);
    b.dyB[i] = ar.take<float>(M * 128);
    b.dr[i] = ar.take<float>(M * 128);
    b.wide0[i] = ar.take<float>(M * F);
    b.wide1[i] = ar.take<float>(M * F);
    b.du[i] = ar.take<float>(M * 64);
    b.dwc[i] = ar.take<float>((size_t)128 * H * 128);
  }
  auto take_img = [&](int n_ct) {
    ar.off = align_up(ar.off, 1024);
    return ar.take<uint8_t>(plain_img_bytes(M, n_ct));
  };
  for (int i = 0; i < 2; i++) {
    b.img_x[i] = take_img(2);
    b.img_d[i] = take_img(2);
    b.img_w0[i] = take_img(F / 64);
    b.img_w1[i] = take_img(F / 64);
    b.img_g[i] = take_img(H * 2);
  }
  b.img_u = take_img(1);
  if (c->attention_type == TW_ATTENTION_LOCAL) {
    const int HD = H * 128;
    for (int i = 0; i < 2; i++) {
      b.lqkv[i] = ar.take<float>(M * 3 * HD);
      b.ldqkv[i] = ar.take<float>(M * 3 * HD);
      b.latt[i] = ar.take<float>(M * HD);
      b.ldatt[i] = ar.take<float>(M * HD);
      b.img_dq[i] = take_img(3 * HD / 64);
      ar.off = align_up(ar.off, 1024);
      b.img_wq[i] = ar.take<uint8_t>(plain_img_bytes(3 * HD, 2));
      b.img_wo[i] = ar.take<uint8_t>(plain_img_bytes(128, HD / 64));
    }
  }
  b.sgrad = ar.take<float>((size_t)B * H * V * V);
  b.dxc = ar.take<float>(M * 3);
  b.dxv = ar.take<float>(M * 3);
  if (out) *out = b;
  return align_up(ar.off, 1024);
}

// gradient table: same order as the parameter table; entries may be NULL (parameter frozen)
struct GradView {
  ParamView pv;  // index arithmetic only
  void* const* g;
  float* at(int i) const { return (float*)g[i]; }
  float* embed() const { return at(0); }
  float* log_scale_c() const { return at(1); }
  float* log_scale_v() const { return at(2); }
  float* in_w(int k, int net, int i) const { return at(pv.net_base(k, net) + 2 * i); }
  float* in_b(int k, int net, int i) const { return at(pv.net_base(k, net) + 2 * i + 1); }
  float* enc(int k, int net, int t, int j) const { return at(pv.net_base(k, net) + pv.per_mlp() + 11 * t + j); }
  float* out_w(int k, int net, int i) const { return at(pv.net_base(k, net) + pv.per_mlp() + 11 * pv.c->num_transformer_layers + 2 * i); }
  float* out_b(int k, int net, int i) const { return at(pv.net_base(k, net) + pv.per_mlp() + 11 * pv.c->num_transformer_layers + 2 * i + 1); }
};

struct BwdCtx {
  const tw_flow_config* c;
  ParamView pv;
  GradView gv;
  TcLayout L;
  const uint8_t* packed;
  Tape tp;
  BwdBuffers b;
  const int64_t* atom_types;
  const float* x_velocs;
  const uint8_t* mask;
  int64_t B, M;
  int V;
  int tiles;  // token tiles
  float* dls;  // [H] gradient w.r.t. the lengthscales of the pass (NULL: not requested)
  bool want_inputs;  // gradients w.r.t. the conditioning state / the target requested
  cudaStream_t st;
};

static int pack_act(BwdCtx& x, float* const X[2], uint8_t* const img[2], int C, float* const colsum[2]) {
  PackArgs a{};
  for (int s = 0; s < 2; s++) a.X[s] = X[s], a.img[s] = img[s], a.colsum[s] = colsum ? colsum[s] : nullptr;
  a.M = x.M, a.C = C, a.ld = C, a.n_ct = C / 64;
  k_pack_act<<<dim3(C / 64, x.tiles, 2), 256, 0, x.st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// operand images of two row-major [rows, C] matrices (one per network)
static int pack_mat(cudaStream_t st, const float* const X[2], uint8_t* const img[2], int64_t rows, int C) {
  PackArgs a{};
  for (int s = 0; s < 2; s++) a.X[s] = X[s], a.img[s] = img[s], a.colsum[s] = nullptr;
  a.M = rows, a.C = C, a.ld = C, a.n_ct = C / 64;
  k_pack_act<<<dim3(C / 64, (unsigned)((rows + 127) / 128), 2), 256, 0, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// `local` attention, forward half shared by the taped pass and the backward's recomputation (local_self_attention.py:46-119):
// qkv = x Wqkv^T on the generic tcgen05 GEMM, masked-softmax attention per (sample, head) on CUDA cores -> att [M, H D].
// Leaves the images of x and of Wqkv in img_x / img_wq.
static int local_attention_qkv_att(const tw_flow_config* c, const ParamView& pv, int k, int t, float* const x[2], uint8_t* const img_x[2],
                                   uint8_t* const img_wq[2], float* const qkv[2], float* const att[2], const float* xc,
                                   const uint8_t* mask, int64_t B, int V, cudaStream_t st) {
  const int H = c->num_heads, HD = H * 128;
  const int64_t M = B * V;
  const int tiles = (int)((M + 127) / 128);
  const float* xin[2] = {x[0], x[1]};
  TW_TRY(pack_mat(st, xin, img_x, M, 128));
  const float* wq[2] = {pv.enc(k, 0, t, 0), pv.enc(k, 1, t, 0)};
  TW_TRY(pack_mat(st, wq, img_wq, 3 * HD, 128));
  GemmArgs g{};
  g.mode = GEMM_NT, g.bn = 128, g.splits = 1;
  for (int s = 0; s < 2; s++) g.A[s] = plain_img(img_x[s], 2), g.B[s] = plain_img(img_wq[s], 2), g.C[s] = qkv[s];
  g.ldc = 3 * HD, g.rows = (int)M, g.cols = 3 * HD, g.tiles_m = tiles, g.tiles_n = 3 * HD / 128, g.KB = 2;
  TW_TRY(launch_gemm(c, g, st));
  return launch_local_attn(qkv[0], qkv[1], att[0], att[1], 2, B, B, V, H, 128, xc, mask, c->max_radius, st);
}

static GemmArgs gemm_base(BwdCtx& x, int mode) {
  GemmArgs g{};
  g.mode = mode;
  g.bn = 128;
  g.splits = 1;
  return g;
}

// number of token-range splits for a weight-gradient GEMM with `out_tiles` output tiles
static int wgrad_splits(const BwdCtx& x, int out_tiles) {
  int s = 74 / (out_tiles > 0 ? out_tiles : 1);
  if (s < 1) s = 1;
  const int kb = x.tiles * 2;
  if (s > kb) s = kb;
  return s;
}

// ---- one conditioner pair (scale + shift network) of coupling layer k, backward -------------------------
// In: b.dst[s] = gradient w.r.t. the outputs of network s.  Out: gradient of the conditioner input added to
// dz_other and the embedding gradient; every parameter gradient of both networks accumulated.
static int conditioner_bwd(BwdCtx& x, int k, float* dz_other, const float* z_other_in) {
  const tw_flow_config* c = x.c;
  const int T = c->num_transformer_layers, F = c->dim_feedforward, H = c->num_heads, hid = c->mlp_hidden_dims[0], E = c->atom_embedding_dim;
  BwdBuffers& b = x.b;
  const TcLayout& L = x.L;
  const uint8_t* wbase[2] = {x.packed + L.net_offset(k, 0), x.packed + L.net_offset(k, 1)};

  // ------------------------------------------------------------------ out_mlp
  {
    float* h3[2] = {x.tp.net[k][0].h[T], x.tp.net[k][1].h[T]};
    TW_TRY(pack_act(x, h3, b.img_x, 128, nullptr));
    GemmArgs g = gemm_base(x, GEMM_NT);  // pre3 = h3 W3^T + b3
    for (int s = 0; s < 2; s++) {
      g.A[s] = plain_img(b.img_x[s], 2);
      g.B[s] = ImgRef{wbase[s] + L.out_w1, 16384u, (uint32_t)(2 * hid * 128), 0u, (uint32_t)(hid * 128)};
      g.C[s] = b.wide0[s], g.bias[s] = x.pv.out_b(k, s, 0);
    }
    g.ldc = hid, g.rows = (int)x.M, g.cols = hid, g.tiles_m = x.tiles, g.tiles_n = hid / 128, g.KB = 2;
    TW_TRY(launch_gemm(c, g, x.st));
    OutLastArgs o{};
    for (int s = 0; s < 2; s++) {
      o.pre[s] = b.wide0[s], o.dst[s] = b.dst[s], o.w4[s] = x.pv.out_w(k, s, 1), o.img_dpre[s] = b.img_w1[s];
      o.dw4[s] = x.gv.out_w(k, s, 1), o.db4[s] = x.gv.out_b(k, s, 1), o.db3[s] = x.gv.out_b(k, s, 0);
    }
    o.M = x.M, o.hid = hid, o.n_ct = hid / 64;
    k_out_last_bwd<<<dim3(hid / 64, x.tiles, 2), 256, 0, x.st>>>(o);
    TW_LAUNCH_CHECK();
    GemmArgs w = gemm_base(x, GEMM_TN);  // dW3 [hid,128] += dpre3^T h3
    for (int s = 0; s < 2; s++) w.A[s] = plain_img(b.img_w1[s], hid / 64), w.B[s] = plain_img(b.img_x[s], 2), w.C[s] = x.gv.out_w(k, s, 0);
    w.ldc = 128, w.rows = hid, w.cols = 128, w.tiles_m = hid / 128, w.tiles_n = 1, w.KB = x.tiles * 2, w.splits = wgrad_splits(x, hid / 128);
    TW_TRY(launch_gemm(c, w, x.st));
    GemmArgs d = gemm_base(x, GEMM_NN);  // dh3 = dpre3 W3
    for (int s = 0; s < 2; s++) {
      d.A[s] = plain_img(b.img_w1[s], hid / 64);
      d.B[s] = ImgRef{wbase[s] + L.out_w1, 16384u, (uint32_t)(2 * hid * 128), 0u, (uint32_t)(hid * 128)};
      d.C[s] = b.dyA[s];
    }
    d.ldc = 128, d.rows = (int)x.M, d.cols = 128, d.tiles_m = x.tiles, d.tiles_n = 1, d.KB = hid / 64;
    TW_TRY(launch_gemm(c, d, x.st));
  }

  // ------------------------------------------------------------------ encoder layers, last to first
  for (int t = T - 1; t >= 0; t--) {
    const uint8_t* eb[2] = {wbase[0] + L.enc0 + (size_t)t * L.enc_stride, wbase[1] + L.enc0 + (size_t)t * L.enc_stride};
    // LN2 backward: dyA -> dr (+ image), dgamma2, dbeta2, db2
    {
      LnBwdArgs a{};
      for (int s = 0; s < 2; s++) {
        a.pre[s] = x.tp.net[k][s].r2[t], a.dy[s] = b.dyA[s], a.gamma[s] = x.pv.enc(k, s, t, 9), a.dr[s] = b.dr[s], a.img_dr[s] = b.img_d[s];
        a.dgamma[s] = x.gv.enc(k, s, t, 9), a.dbeta[s] = x.gv.enc(k, s, t, 10), a.dbias[s] = x.gv.enc(k, s, t, 6);
      }
      a.M = x.M, a.eps = c->layer_norm_eps;
      k_ln_bwd<<<dim3(x.tiles * 4, 2), 256, 0, x.st>>>(a);
      TW_LAUNCH_CHECK();
    }
    // FFN weight images of this layer: W1 [F,128] row tile c at c*128K (+kb*16K), lo +32K; W2 [128,F] col tile b
    ImgRef w1[2], w2[2];
    for (int s = 0; s < 2; s++) {
      w1[s] = ImgRef{eb[s] + L.enc_ffn, 4u * 32768u, 16384u, 0u, 32768u};
      w2[s] = ImgRef{eb[s] + L.enc_ffn + 2 * 32768, 0u, 16384u, 4u * 32768u, 32768u};
    }
    float* y1[2] = {x.tp.net[k][0].y1[t], x.tp.net[k][1].y1[t]};
    TW_TRY(pack_act(x, y1, b.img_x, 128, nullptr));
    {
      FfnBwdPreArgs f{};  // act = relu(y1 W1^T + b1), dpre = (dr W2) where pre > 0, db1: one launch, no [M, F] fp32 matrix
      for (int s = 0; s < 2; s++) {
        f.Ay[s] = plain_img(b.img_x[s], 2), f.Ad[s] = plain_img(b.img_d[s], 2), f.W1[s] = w1[s], f.W2[s] = w2[s];
        f.b1[s] = x.pv.enc(k, s, t, 4), f.img_act[s] = b.img_w0[s], f.img_dpre[s] = b.img_w1[s], f.db1[s] = x.gv.enc(k, s, t, 4);
      }
      f.rows = (int)x.M, f.tiles_m = x.tiles, f.tiles_n = F / 128, f.n_ct = F / 64;
      TW_TRY(launch_ffn_bwd_pre(c, f, x.st));
      GemmArgs wa = gemm_base(x, GEMM_TN);  // dW2 [128,F] += dr^T hid
      for (int s = 0; s < 2; s++) wa.A[s] = plain_img(b.img_d[s], 2), wa.B[s] = plain_img(b.img_w0[s], F / 64), wa.C[s] = x.gv.enc(k, s, t, 5);
      wa.ldc = F, wa.rows = 128, wa.cols = F, wa.tiles_m = 1, wa.tiles_n = F / 128, wa.KB = x.tiles * 2, wa.splits = wgrad_splits(x, F / 128);
      TW_TRY(launch_gemm(c, wa, x.st));
      GemmArgs wb = gemm_base(x, GEMM_TN);  // dW1 [F,128] += dpre^T y1
      for (int s = 0; s < 2; s++) wb.A[s] = plain_img(b.img_w1[s], F / 64), wb.B[s] = plain_img(b.img_x[s], 2), wb.C[s] = x.gv.enc(k, s, t, 3);
      wb.ldc = 128, wb.rows = F, wb.cols = 128, wb.tiles_m = F / 128, wb.tiles_n = 1, wb.KB = x.tiles * 2, wb.splits = wgrad_splits(x, F / 128);
      TW_TRY(launch_gemm(c, wb, x.st));
      GemmArgs e = gemm_base(x, GEMM_NN);  // dy1 = dr + dpre W1
      for (int s = 0; s < 2; s++) e.A[s] = plain_img(b.img_w1[s], F / 64), e.B[s] = w1[s], e.C[s] = b.dyB[s], e.resid[s] = b.dr[s];
      e.ldc = 128, e.ldr = 128, e.rows = (int)x.M, e.cols = 128, e.tiles_m = x.tiles, e.tiles_n = 1, e.KB = F / 64;
      TW_TRY(launch_gemm(c, e, x.st));
    }
    // LN1 backward: dyB -> dr (+ image), dgamma1, dbeta1
    {
      LnBwdArgs a{};
      for (int s = 0; s < 2; s++) {
        a.pre[s] = x.tp.net[k][s].r1[t], a.dy[s] = b.dyB[s], a.gamma[s] = x.pv.enc(k, s, t, 7), a.dr[s] = b.dr[s], a.img_dr[s] = b.img_d[s];
        a.dgamma[s] = x.gv.enc(k, s, t, 7), a.dbeta[s] = x.gv.enc(k, s, t, 8), a.dbias[s] = nullptr;
      }
      a.M = x.M, a.eps = c->layer_norm_eps;
      k_ln_bwd<<<dim3(x.tiles * 4, 2), 256, 0, x.st>>>(a);
      TW_LAUNCH_CHECK();
    }
    if (x.pv.local()) {
      // local attention: r1 = x + att(x Wqkv^T) Wo^T (local_self_attention.py:46-119).  qkv and att are recomputed from the layer
      // input; projections and their weight / input gradients on the generic tcgen05 GEMM, the attention core on CUDA cores.
      const int HD = H * 128;
      float* hin[2] = {x.tp.net[k][0].h[t], x.tp.net[k][1].h[t]};
      TW_TRY(local_attention_qkv_att(c, x.pv, k, t, hin, b.img_x, b.img_wq, b.lqkv, b.latt, x.tp.xc, x.mask, x.B, x.V, x.st));
      const float* att_c[2] = {b.latt[0], b.latt[1]};
      TW_TRY(pack_mat(x.st, att_c, b.img_g, x.M, HD));
      GemmArgs w = gemm_base(x, GEMM_TN);  // dWo [128, H D] += dr^T att
      for (int s = 0; s < 2; s++) w.A[s] = plain_img(b.img_d[s], 2), w.B[s] = plain_img(b.img_g[s], HD / 64), w.C[s] = x.gv.enc(k, s, t, 2);
      w.ldc = HD, w.rows = 128, w.cols = HD, w.tiles_m = 1, w.tiles_n = HD / 128, w.KB = x.tiles * 2, w.splits = wgrad_splits(x, HD / 128);
      TW_TRY(launch_gemm(c, w, x.st));
      const float* wo[2] = {x.pv.enc(k, 0, t, 2), x.pv.enc(k, 1, t, 2)};
      TW_TRY(pack_mat(x.st, wo, b.img_wo, 128, HD));
      GemmArgs d = gemm_base(x, GEMM_NN);  // datt [M, H D] = dr Wo
      for (int s = 0; s < 2; s++) d.A[s] = plain_img(b.img_d[s], 2), d.B[s] = plain_img(b.img_wo[s], HD / 64), d.C[s] = b.ldatt[s];
      d.ldc = HD, d.rows = (int)x.M, d.cols = HD, d.tiles_m = x.tiles, d.tiles_n = HD / 128, d.KB = 2;
      TW_TRY(launch_gemm(c, d, x.st));
      TW_TRY(launch_local_attn_bwd(b.lqkv[0], b.lqkv[1], b.ldatt[0], b.ldatt[1], b.ldqkv[0], b.ldqkv[1], 2, x.B, x.B, x.V, H, 128, x.tp.xc,
                                   x.mask, c->max_radius, x.st));
      const float* dq_c[2] = {b.ldqkv[0], b.ldqkv[1]};
      TW_TRY(pack_mat(x.st, dq_c, b.img_dq, x.M, 3 * HD));
      GemmArgs wq = gemm_base(x, GEMM_TN);  // dWqkv [3 H D, 128] += dqkv^T x
      for (int s = 0; s < 2; s++) wq.A[s] = plain_img(b.img_dq[s], 3 * HD / 64), wq.B[s] = plain_img(b.img_x[s], 2), wq.C[s] = x.gv.enc(k, s, t, 0);
      wq.ldc = 128, wq.rows = 3 * HD, wq.cols = 128, wq.tiles_m = 3 * HD / 128, wq.tiles_n = 1, wq.KB = x.tiles * 2, wq.splits = wgrad_splits(x, 3 * HD / 128);
      TW_TRY(launch_gemm(c, wq, x.st));
      GemmArgs e = gemm_base(x, GEMM_NN);  // dx = dr + dqkv Wqkv
      for (int s = 0; s < 2; s++) e.A[s] = plain_img(b.img_dq[s], 3 * HD / 64), e.B[s] = plain_img(b.img_wq[s], 2), e.C[s] = b.dyA[s], e.resid[s] = b.dr[s];
      e.ldc = 128, e.ldr = 128, e.rows = (int)x.M, e.cols = 128, e.tiles_m = x.tiles, e.tiles_n = 1, e.KB = 3 * HD / 64;
      TW_TRY(launch_gemm(c, e, x.st));
      continue;
    }
    // attention: r1 = x + sum_h W_c,h (A_h x)
    if (x.dls || x.want_inputs) {  // S_h += (dr W_c,h) x^T; the [M, F] scratch of the FFN block is free here
      GemmArgs q = gemm_base(x, GEMM_NN);  // Q [M, H*128] = dr W_c
      for (int s = 0; s < 2; s++) q.A[s] = plain_img(b.img_d[s], 2), q.B[s] = plain_img(eb[s] + L.enc_wc, H * 2), q.C[s] = b.wide0[s];
      q.ldc = H * 128, q.rows = (int)x.M, q.cols = H * 128, q.tiles_m = x.tiles, q.tiles_n = H, q.KB = 2;
      TW_TRY(launch_gemm(c, q, x.st));
      ScoreGradArgs sa{};
      for (int s = 0; s < 2; s++) sa.q[s] = b.wide0[s], sa.x[s] = x.tp.net[k][s].h[t];
      sa.S = b.sgrad;
      k_score_grad<<<dim3((unsigned)x.B, H), 128, ((size_t)x.V * 129 + 4 * 128) * sizeof(float), x.st>>>(sa, x.V, H, 0, 2);
      TW_LAUNCH_CHECK();
    }
    {
      const size_t wc_bytes = (size_t)128 * H * 128 * sizeof(float);
      for (int s = 0; s < 2; s++) TW_CUDA(cudaMemsetAsync(b.dwc[s], 0, wc_bytes, x.st));
      GemmArgs w = gemm_base(x, GEMM_TN);  // dWc [128, H*128] = dr^T mixed
      for (int s = 0; s < 2; s++) w.A[s] = plain_img(b.img_d[s], 2), w.B[s] = plain_img(x.tp.net[k][s].mixed[t], H * 2), w.C[s] = b.dwc[s];
      w.ldc = H * 128, w.rows = 128, w.cols = H * 128, w.tiles_m = 1, w.tiles_n = H, w.KB = x.tiles * 2, w.splits = wgrad_splits(x, H);
      TW_TRY(launch_gemm(c, w, x.st));
      WcChainArgs ch{};
      for (int s = 0; s < 2; s++)
        ch.dwc[s] = b.dwc[s], ch.wo[s] = x.pv.enc(k, s, t, 2), ch.wv[s] = x.pv.enc(k, s, t, 0), ch.dwo[s] = x.gv.enc(k, s, t, 2), ch.dwv[s] = x.gv.enc(k, s, t, 0);
      static DeviceOnce chain_attr;
      if (!chain_attr.done()) {
        TW_CUDA(cudaFuncSetAttribute(k_wc_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, kWcChainSmem));
        chain_attr.mark();
      }
      k_wc_chain<<<dim3(H, 8, 4), 256, kWcChainSmem, x.st>>>(ch, H);
      TW_LAUNCH_CHECK();
      // G_h = A_h^T dr  (transposed score images), then dx = dr + sum_h G_h W_c,h
      if (x.pv.chebyshev()) {
        // per network: rebuild this layer's scores (fp32 + transposed images), G_h, and the coefficient gradient
        GemmArgs q = gemm_base(x, GEMM_NN);  // Q [M, H*128] = dr W_c (both networks)
        for (int s = 0; s < 2; s++) q.A[s] = plain_img(b.img_d[s], 2), q.B[s] = plain_img(eb[s] + L.enc_wc, H * 2), q.C[s] = b.wide0[s];
        q.ldc = H * 128, q.rows = (int)x.M, q.cols = H * 128, q.tiles_m = x.tiles, q.tiles_n = H, q.KB = 2;
        TW_TRY(launch_gemm(c, q, x.st));
        for (int net = 0; net < 2; net++) {
          const float* coef = x.pv.cheb(k, net, t);
          TW_TRY(launch_scores(x.tp.xc, x.mask, x.pv.enc(0, 0, 0, 1), x.B, x.V, H, x.tp.scores, x.st, coef, c->cheb_order, c->force_asymptotic_zero));
          TW_TRY(tc_scores_images(c, x.tp.scores, x.B, x.V, x.tp.scores_img_t, 1, x.st));
          float* const dr1[2] = {b.dr[net], b.dr[net]};
          uint8_t* const g1[2] = {b.img_g[net], b.img_g[net]};
          TW_TRY(tc_mix(c, dr1, g1, x.tp.scores_img_t, x.B, x.B, x.V, x.st, 1));
          float* dcoef = x.gv.at(x.pv.regular() + (k * 2 + net) * c->num_transformer_layers + t);
          if (dcoef) {
            TW_CUDA(cudaMemsetAsync(b.sgrad, 0, (size_t)x.B * H * x.V * x.V * sizeof(float), x.st));
            ScoreGradArgs sa{};
            for (int s = 0; s < 2; s++) sa.q[s] = b.wide0[s], sa.x[s] = x.tp.net[k][s].h[t];
            sa.S = b.sgrad;
            k_score_grad<<<dim3((unsigned)x.B, H), 128, ((size_t)x.V * 129 + 4 * 128) * sizeof(float), x.st>>>(sa, x.V, H, net, net + 1);
            TW_LAUNCH_CHECK();
            k_cheb_grad<<<(unsigned)x.B, 256, 0, x.st>>>(x.tp.xc, x.mask, x.pv.enc(0, 0, 0, 1), coef, c->cheb_order, c->force_asymptotic_zero,
                                                        b.sgrad, x.V, H, dcoef);
            TW_LAUNCH_CHECK();
          }
        }
      } else {
        TW_TRY(tc_mix(c, b.dr, b.img_g, x.tp.scores_img_t, x.B, x.B, x.V, x.st));
      }
      GemmArgs d = gemm_base(x, GEMM_NN_HEADED);
      for (int s = 0; s < 2; s++) {
        d.A[s] = plain_img(b.img_g[s], H * 2);
        d.B[s] = plain_img(eb[s] + L.enc_wc, H * 2);
        d.C[s] = b.dyA[s], d.resid[s] = b.dr[s];
      }
      d.ldc = 128, d.ldr = 128, d.rows = (int)x.M, d.cols = 128, d.tiles_m = x.tiles, d.tiles_n = 1, d.KB = H * 2;
      TW_TRY(launch_gemm(c, d, x.st));
    }
  }

  // ------------------------------------------------------------------ in_mlp (dyA = gradient w.r.t. h0)
  {
    k_features_img<<<x.tiles * 4, 256, 0, x.st>>>(x.pv.embed(), x.atom_types, x.tp.xc, x.x_velocs, z_other_in, x.M, x.V, E, c->num_atom_types, b.img_u);
    TW_LAUNCH_CHECK();
    ImgRef w1[2], w2[2];
    for (int s = 0; s < 2; s++) {
      w1[s] = ImgRef{wbase[s] + L.in_w1, 16384u, 0u, 0u, (uint32_t)(hid * 128)};  // [hid x 64]: row tile tr at tr*16K
      w2[s] = plain_img(wbase[s] + L.in_w2, hid / 64);                            // [128 x hid]
    }
    GemmArgs g = gemm_base(x, GEMM_NT);  // pre1 = u W1^T + b1
    for (int s = 0; s < 2; s++) g.A[s] = plain_img(b.img_u, 1), g.B[s] = w1[s], g.C[s] = b.wide0[s], g.bias[s] = x.pv.in_b(k, s, 0);
    g.ldc = hid, g.rows = (int)x.M, g.cols = hid, g.tiles_m = x.tiles, g.tiles_n = hid / 128, g.KB = 1;
    TW_TRY(launch_gemm(c, g, x.st));
    float* db2[2] = {x.gv.in_b(k, 0, 1), x.gv.in_b(k, 1, 1)};
    TW_TRY(pack_act(x, b.dyA, b.img_d, 128, db2));
    GemmArgs d = gemm_base(x, GEMM_NN);  // dact1 = dh0 W2
    for (int s = 0; s < 2; s++) d.A[s] = plain_img(b.img_d[s], 2), d.B[s] = w2[s], d.C[s] = b.wide1[s];
    d.ldc = hid, d.rows = (int)x.M, d.cols = hid, d.tiles_m = x.tiles, d.tiles_n = hid / 128, d.KB = 2;
    TW_TRY(launch_gemm(c, d, x.st));
    ActBwdArgs r{};
    for (int s = 0; s < 2; s++)
      r.pre[s] = b.wide0[s], r.dact[s] = b.wide1[s], r.img_act[s] = b.img_w0[s], r.img_dpre[s] = b.img_w1[s], r.dbias[s] = x.gv.in_b(k, s, 0);
    r.M = x.M, r.C = hid, r.n_ct = hid / 64;
    k_act_bwd<ACT_SILU><<<dim3(hid / 64, x.tiles, 2), 256, 0, x.st>>>(r);
    TW_LAUNCH_CHECK();
    GemmArgs wa = gemm_base(x, GEMM_TN);  // dW2 [128,hid] += dh0^T act1
    for (int s = 0; s < 2; s++) wa.A[s] = plain_img(b.img_d[s], 2), wa.B[s] = plain_img(b.img_w0[s], hid / 64), wa.C[s] = x.gv.in_w(k, s, 1);
    wa.ldc = hid, wa.rows = 128, wa.cols = hid, wa.tiles_m = 1, wa.tiles_n = hid / 128, wa.KB = x.tiles * 2, wa.splits = wgrad_splits(x, hid / 128);
    TW_TRY(launch_gemm(c, wa, x.st));
    GemmArgs wb = gemm_base(x, GEMM_TN);  // dW1 [hid, E+9] += dpre1^T u
    for (int s = 0; s < 2; s++) wb.A[s] = plain_img(b.img_w1[s], hid / 64), wb.B[s] = plain_img(b.img_u, 1), wb.C[s] = x.gv.in_w(k, s, 0);
    wb.bn = 64, wb.ldc = E + 9, wb.rows = hid, wb.cols = E + 9, wb.tiles_m = hid / 128, wb.tiles_n = 1, wb.KB = x.tiles * 2, wb.splits = wgrad_splits(x, hid / 128);
    TW_TRY(launch_gemm(c, wb, x.st));
    GemmArgs e = gemm_base(x, GEMM_NN);  // du = dpre1 W1  [M,64]
    for (int s = 0; s < 2; s++) e.A[s] = plain_img(b.img_w1[s], hid / 64), e.B[s] = w1[s], e.C[s] = b.du[s];
    e.bn = 64, e.ldc = 64, e.rows = (int)x.M, e.cols = 64, e.tiles_m = x.tiles, e.tiles_n = 1, e.KB = hid / 64;
    TW_TRY(launch_gemm(c, e, x.st));
    k_du_scatter<<<x.tiles * 4, 256, c->num_atom_types * E * sizeof(float), x.st>>>(b.du[0], b.du[1], x.atom_types, x.M, E, c->num_atom_types,
                                                                             dz_other, x.gv.embed(), x.want_inputs ? b.dxc : nullptr,
                                                                             x.want_inputs ? b.dxv : nullptr);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

}  // namespace tw

using namespace tw;

extern "C" {

// Debug / test hook: one generic image GEMM on fp32 row-major inputs (packed to operand images first).
//   mode 0 NT: C[ar,br] = A B^T   1 NN: C[ar,bc] = A B   2 headed NN: C[ar,128] = sum_h A[:,h] B[:,h]   3 TN: C[ac,bc] += A^T B
int tw_debug_gemm(int mode, int precision, const float* A, int a_rows, int a_cols, const float* B, int b_rows, int b_cols, float* C,
                  int bn, int splits, void* workspace, size_t workspace_bytes, void* stream) {
  TW_CHECK_ARG(A && B && C && workspace, "NULL pointer");
  TW_CHECK_ARG(a_cols % 64 == 0 && b_cols % 64 == 0, "columns must be multiples of 64");
  TW_CHECK_ARG(precision == TW_PRECISION_BF16X3 || precision == TW_PRECISION_BF16, "tensor-core precisions only");
  tw_flow_config cfg{};
  cfg.precision = precision;
  cudaStream_t st = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  ar.off = align_up(ar.off, 1024);
  uint8_t* ia = ar.take<uint8_t>(plain_img_bytes(a_rows, a_cols / 64));
  ar.off = align_up(ar.off, 1024);
  uint8_t* ib = ar.take<uint8_t>(plain_img_bytes(b_rows, b_cols / 64));
  if (!ar.ok()) return fail(TW_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, ar.off);
  for (int which = 0; which < 2; which++) {
    PackArgs p{};
    p.X[0] = which ? B : A, p.img[0] = which ? ib : ia;
    p.M = which ? b_rows : a_rows, p.C = which ? b_cols : a_cols, p.ld = p.C, p.n_ct = p.C / 64;
    k_pack_act<<<dim3(p.n_ct, (unsigned)((p.M + 127) / 128), 1), 256, 0, st>>>(p);
    TW_LAUNCH_CHECK();
  }
  GemmArgs g{};
  g.mode = mode, g.bn = bn, g.splits = splits, g.nets = 1;
  g.A[0] = plain_img(ia, a_cols / 64), g.B[0] = plain_img(ib, b_cols / 64), g.C[0] = C;
  const int tiles_a = (a_rows + 127) / 128;
  if (mode == GEMM_NT) {
    TW_CHECK_ARG(a_cols == b_cols, "NT: inner dimensions differ");
    g.rows = a_rows, g.cols = b_rows, g.ldc = b_rows, g.tiles_m = tiles_a, g.tiles_n = (b_rows + bn - 1) / bn, g.KB = a_cols / 64;
  } else if (mode == GEMM_NN) {
    TW_CHECK_ARG(a_cols == b_rows, "NN: inner dimensions differ");
    g.rows = a_rows, g.cols = b_cols, g.ldc = b_cols, g.tiles_m = tiles_a, g.tiles_n = b_cols / bn, g.KB = a_cols / 64;
  } else if (mode == GEMM_NN_HEADED) {
    TW_CHECK_ARG(a_cols == b_cols && b_rows == 128, "headed: A [M,H*128], B [128,H*128]");
    g.rows = a_rows, g.cols = 128, g.ldc = 128, g.tiles_m = tiles_a, g.tiles_n = 1, g.KB = a_cols / 64;
  } else {
    TW_CHECK_ARG(a_rows == b_rows && a_cols % 128 == 0, "TN: token counts differ / A columns not a multiple of 128");
    g.rows = a_cols, g.cols = b_cols, g.ldc = b_cols, g.tiles_m = a_cols / 128, g.tiles_n = b_cols / bn, g.KB = tiles_a * 2;
  }
  return launch_gemm(&cfg, g, st);
}

int tw_flow_train_bytes(const tw_flow_config* cfg, int64_t B, int64_t V, size_t* tape_bytes, size_t* workspace_bytes) {
  void* dummy = (void*)cfg;
  TW_TRY(check_train(cfg, &dummy, B, V));
  if (tape_bytes) *tape_bytes = carve_tape(cfg, B, (int)V, nullptr, 0, nullptr) + 1024;
  if (workspace_bytes) *workspace_bytes = carve_bwd(cfg, B, (int)V, nullptr, 0, nullptr) + 1024;
  return TW_OK;
}

// conditioner pair of coupling layer k on `z_other`, every layer boundary written to the tape; (s, t) end up in net[k][.].st
static int taped_conditioner(const tw_flow_config* cfg, const ParamView& pv, const void* packed_weights, Tape& tp, int k,
                             const float* z_other, const int64_t* atom_types, const float* x_velocs, const uint8_t* mask, int64_t B,
                             int64_t V, cudaStream_t st) {
  const int T = cfg->num_transformer_layers;
  const int64_t M = B * V;
  TcScratch tc{};
  tc.packed = (const uint8_t*)packed_weights;
  tc.scores_img = tp.scores_img;
  float* h0[2] = {tp.net[k][0].h[0], tp.net[k][1].h[0]};
  TW_TRY(tc_in_mlp(cfg, pv, k, tc, atom_types, tp.xc, x_velocs, z_other, h0, B, B, (int)V, st));
  for (int t = 0; t < T; t++) {
    float* hin[2] = {tp.net[k][0].h[t], tp.net[k][1].h[t]};
    float* y1[2] = {tp.net[k][0].y1[t], tp.net[k][1].y1[t]};
    float* r1[2] = {tp.net[k][0].r1[t], tp.net[k][1].r1[t]};
    float* r2[2] = {tp.net[k][0].r2[t], tp.net[k][1].r2[t]};
    float* hout[2] = {tp.net[k][0].h[t + 1], tp.net[k][1].h[t + 1]};
    tc.mixed_img[0] = tp.net[k][0].mixed[t], tc.mixed_img[1] = tp.net[k][1].mixed[t];
    if (M % 128 && !pv.local()) {  // rows past the last token of the tail tile are read by the weight-gradient GEMM: keep them zero
      const size_t tile_bytes = tc_mixed_img_bytes(cfg, 128);
      for (int s = 0; s < 2; s++) TW_CUDA(cudaMemsetAsync(tc.mixed_img[s] + (size_t)(M / 128) * tile_bytes, 0, tile_bytes, st));
    }
    if (pv.local()) {
      // qkv projection (tcgen05 GEMM) -> masked softmax attention within max_radius (CUDA cores) -> output projection + residual
      // (tcgen05 GEMM) -> LayerNorm 1.  Nothing of the attention is taped: the backward recomputes qkv and att from the layer input.
      const int HD = cfg->num_heads * 128;
      TW_TRY(local_attention_qkv_att(cfg, pv, k, t, hin, tp.img_x, tp.img_wq, tp.qkv, tp.att, tp.xc, mask, B, (int)V, st));
      const float* att_c[2] = {tp.att[0], tp.att[1]};
      TW_TRY(pack_mat(st, att_c, tp.img_att, M, HD));
      const float* wo[2] = {pv.enc(k, 0, t, 2), pv.enc(k, 1, t, 2)};
      TW_TRY(pack_mat(st, wo, tp.img_wo, 128, HD));
      GemmArgs g{};
      g.mode = GEMM_NT, g.bn = 128, g.splits = 1;
      for (int s = 0; s < 2; s++)
        g.A[s] = plain_img(tp.img_att[s], HD / 64), g.B[s] = plain_img(tp.img_wo[s], HD / 64), g.C[s] = r1[s], g.resid[s] = hin[s];
      g.ldc = 128, g.ldr = 128, g.rows = (int)M, g.cols = 128, g.tiles_m = (int)((M + 127) / 128), g.tiles_n = 1, g.KB = HD / 64;
      TW_TRY(launch_gemm(cfg, g, st));
      for (int s = 0; s < 2; s++) TW_CUDA(cudaMemcpyAsync(y1[s], r1[s], (size_t)M * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      TW_TRY(launch_layernorm(y1[0], y1[1], pv.enc(k, 0, t, 7), pv.enc(k, 1, t, 7), pv.enc(k, 0, t, 8), pv.enc(k, 1, t, 8), 2, M, 128,
                              cfg->layer_norm_eps, st));
    } else if (pv.chebyshev()) {
      // every attention layer of every network has its own basis function, hence its own scores (kernel_attention.py:333-335):
      // the forward images are rebuilt in place per (layer, network); the backward rebuilds the transposed ones the same way
      for (int net = 0; net < 2; net++) {
        TW_TRY(tc_begin_pass_direct(cfg, tc, tp.xc, mask, pv.enc(0, 0, 0, 1), B, (int)V, st, pv.cheb(k, net, t), false));
        TW_TRY(tc_attention_layer(cfg, pv, k, t, tc, hin, y1, B, B, (int)V, st, r1, net));
      }
    } else {
      TW_TRY(tc_attention_layer(cfg, pv, k, t, tc, hin, y1, B, B, (int)V, st, r1));
    }
    TW_TRY(tc_ffn_layer(cfg, pv, k, t, tc, y1, hout, M, st, r2));
  }
  float* hl[2] = {tp.net[k][0].h[T], tp.net[k][1].h[T]};
  float* so[2] = {tp.net[k][0].st, tp.net[k][1].st};
  TW_TRY(tc_out_mlp(cfg, pv, k, tc, hl, so, M, st));
  return TW_OK;
}

static int begin_taped_pass(const tw_flow_config* cfg, const ParamView& pv, Tape& tp, const float* x_coords, const uint8_t* mask,
                            int64_t B, int64_t V, cudaStream_t st) {
  TW_TRY(launch_prep(x_coords, mask, B, (int)V, tp.xc, tp.com, st));
  if (!pv.chebyshev() && !pv.local()) {  // (chebyshev_kernel: the scores belong to an attention layer, not to the pass -- taped_conditioner;
                                           //  local: no position-only scores at all)
    TW_TRY(launch_scores(tp.xc, mask, pv.enc(0, 0, 0, 1), B, (int)V, cfg->num_heads, tp.scores, st));
    TW_TRY(tc_scores_images(cfg, tp.scores, B, (int)V, tp.scores_img, 0, st));
    TW_TRY(tc_scores_images(cfg, tp.scores, B, (int)V, tp.scores_img_t, 1, st));
  }
  TW_CUDA(cudaMemsetAsync(tp.delta, 0, B * sizeof(float), st));
  return TW_OK;
}

int tw_flow_log_likelihood_train(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types,
                                 const float* x_coords, const float* x_velocs, const float* y_coords, const float* y_velocs,
                                 const uint8_t* mask, int64_t B, int64_t V, int32_t flags, float* out_log_prob,
                                 const void* packed_weights, void* tape, size_t tape_bytes, void* stream) {
  TW_TRY(check_train(cfg, params, B, V));
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(atom_types && x_coords && x_velocs && y_coords && y_velocs && mask && out_log_prob, "NULL pointer");
  TW_CHECK_ARG(packed_weights && ((uintptr_t)packed_weights & 1023) == 0, "packed_weights missing or not 1024-byte aligned");
  TW_CHECK_ARG(tape && ((uintptr_t)tape & 1023) == 0, "tape missing or not 1024-byte aligned");
  Tape tp;
  const size_t need = carve_tape(cfg, B, (int)V, tape, tape_bytes, &tp);
  if (need > tape_bytes) return fail(TW_ERR_WORKSPACE, "tape %zu < %zu", tape_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  ParamView pv{cfg, params};
  const int L = cfg->num_coupling_layers;
  const int64_t M = B * V, cnt = M * 3;
  TW_TRY(begin_taped_pass(cfg, pv, tp, x_coords, mask, B, V, st));
  if (flags & TW_FLOW_DISPLACEMENT_TARGET)
    TW_TRY(launch_sub(y_coords, x_coords, cnt, tp.z[0][0], st));
  else
    TW_CUDA(cudaMemcpyAsync(tp.z[0][0], y_coords, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TW_CUDA(cudaMemcpyAsync(tp.z[0][1], y_velocs, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int k = 0; k < L; k++) {
    const bool pos = (k % 2) == cfg->position_layer_index_mod_2;
    TW_TRY(taped_conditioner(cfg, pv, packed_weights, tp, k, pos ? tp.z[k][1] : tp.z[k][0], atom_types, x_velocs, mask, B, V, st));
    // next state: copy both halves, then transform the target half in place
    TW_CUDA(cudaMemcpyAsync(tp.z[k + 1][0], tp.z[k][0], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TW_CUDA(cudaMemcpyAsync(tp.z[k + 1][1], tp.z[k][1], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TW_TRY(launch_coupling(tp.net[k][0].st, tp.net[k][1].st, pos ? tp.z[k + 1][0] : tp.z[k + 1][1], mask, tp.delta, B, B, (int)V, 0, nullptr,
                           nullptr, st));
  }
  TW_TRY(launch_prior(tp.z[L][0], tp.z[L][1], mask, pv.log_scale_c(), pv.log_scale_v(), tp.delta, -1.f, B, B, (int)V, out_log_prob, st));
  return TW_OK;
}

// Sampling direction with a tape (conditional_sample_with_logp under autograd: the energy-based losses, losses.py:396-664):
// z[L] = the latent draws, layers L-1 .. 0 in reverse mode; the tape ends up holding exactly what a density pass on the
// resulting y would record (layer k's conditioner reads the half that layer k leaves unchanged), so the backward shares
// conditioner_bwd.  out_delta = sum of log-scales (log p(y|x) = prior(z) + delta; the caller adds the prior term).
int tw_flow_sample_train(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types, const float* x_coords,
                         const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V, int32_t flags, const float* z_coords,
                         const float* z_velocs, float* out_y_coords, float* out_y_velocs, float* out_delta,
                         const void* packed_weights, void* tape, size_t tape_bytes, void* stream) {
  TW_TRY(check_train(cfg, params, B, V));
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(atom_types && x_coords && x_velocs && z_coords && z_velocs && mask && out_y_coords && out_y_velocs && out_delta, "NULL pointer");
  TW_CHECK_ARG(packed_weights && ((uintptr_t)packed_weights & 1023) == 0, "packed_weights missing or not 1024-byte aligned");
  TW_CHECK_ARG(tape && ((uintptr_t)tape & 1023) == 0, "tape missing or not 1024-byte aligned");
  Tape tp;
  const size_t need = carve_tape(cfg, B, (int)V, tape, tape_bytes, &tp);
  if (need > tape_bytes) return fail(TW_ERR_WORKSPACE, "tape %zu < %zu", tape_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  ParamView pv{cfg, params};
  const int L = cfg->num_coupling_layers;
  const int64_t cnt = B * V * 3;
  TW_TRY(begin_taped_pass(cfg, pv, tp, x_coords, mask, B, V, st));
  TW_CUDA(cudaMemcpyAsync(tp.z[L][0], z_coords, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TW_CUDA(cudaMemcpyAsync(tp.z[L][1], z_velocs, cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int k = L - 1; k >= 0; k--) {
    const bool pos = (k % 2) == cfg->position_layer_index_mod_2;
    TW_TRY(taped_conditioner(cfg, pv, packed_weights, tp, k, pos ? tp.z[k + 1][1] : tp.z[k + 1][0], atom_types, x_velocs, mask, B, V, st));
    TW_CUDA(cudaMemcpyAsync(tp.z[k][0], tp.z[k + 1][0], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TW_CUDA(cudaMemcpyAsync(tp.z[k][1], tp.z[k + 1][1], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TW_TRY(launch_coupling(tp.net[k][0].st, tp.net[k][1].st, pos ? tp.z[k][0] : tp.z[k][1], mask, tp.delta, B, B, (int)V, 1, nullptr, nullptr, st));
  }
  if (flags & TW_FLOW_DISPLACEMENT_TARGET)
    TW_TRY(launch_uncentre(tp.xc, tp.com, tp.z[0][0], B, B, (int)V, out_y_coords, st));
  else
    TW_CUDA(cudaMemcpyAsync(out_y_coords, tp.z[0][0], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TW_CUDA(cudaMemcpyAsync(out_y_velocs, tp.z[0][1], cnt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  TW_CUDA(cudaMemcpyAsync(out_delta, tp.delta, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return TW_OK;
}

// shared set-up of the two backward entry points
static int begin_backward(BwdCtx& x, const tw_flow_config* cfg, const void* const* params, void* const* grads, const int64_t* atom_types,
                          const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V, const void* packed_weights, void* tape,
                          size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  TW_CHECK_ARG(grads && atom_types && x_velocs && mask, "NULL pointer");
  TW_CHECK_ARG(packed_weights && ((uintptr_t)packed_weights & 1023) == 0, "packed_weights missing or not 1024-byte aligned");
  TW_CHECK_ARG(tape && ((uintptr_t)tape & 1023) == 0 && workspace && ((uintptr_t)workspace & 1023) == 0, "tape / workspace missing or not 1024-byte aligned");
  x.c = cfg;
  x.pv = ParamView{cfg, params};
  x.gv = GradView{ParamView{cfg, nullptr}, grads};
  for (int i = 0; i < x.pv.total(); i++) {
    const bool is_ls = i >= 3 && ((i - 3) % x.pv.per_net()) >= x.pv.per_mlp() && ((i - 3) % x.pv.per_net()) < x.pv.per_mlp() + 11 * cfg->num_transformer_layers &&
                       (((i - 3) % x.pv.per_net()) - x.pv.per_mlp()) % 11 == 1;
    const bool is_cheb = i >= x.pv.regular();  // trailing cheb_coeffs section: NULL = frozen
    TW_CHECK_ARG(is_ls || is_cheb || i == 1 || i == 2 || grads[i] != nullptr, "gradient table entry %d is NULL", i);
  }
  x.L = TcLayout::make(cfg);
  x.packed = (const uint8_t*)packed_weights;
  size_t need = carve_tape(cfg, B, (int)V, tape, tape_bytes, &x.tp);
  if (need > tape_bytes) return fail(TW_ERR_WORKSPACE, "tape %zu < %zu", tape_bytes, need);
  need = carve_bwd(cfg, B, (int)V, workspace, workspace_bytes, &x.b);
  if (need > workspace_bytes) return fail(TW_ERR_WORKSPACE, "workspace %zu < %zu", workspace_bytes, need);
  x.atom_types = atom_types, x.x_velocs = x_velocs, x.mask = mask;
  x.B = B, x.V = (int)V, x.M = B * V, x.tiles = (int)((x.M + 127) / 128);
  x.st = (cudaStream_t)stream;
  // lengthscale gradient (learnable_kernel): requested by a non-NULL entry for the lengthscales of chain[0].scale.layer[0]
  x.dls = x.pv.local() ? nullptr : x.gv.enc(0, 0, 0, 1);  // (local: that slot is unused; the radius mask has no gradient)
  if (x.pv.chebyshev()) {
    TW_CHECK_ARG(!x.dls && !x.want_inputs, "chebyshev_kernel: gradients w.r.t. lengthscales / conditioning state are not built");
    static DeviceOnce attr_cheb;
    if (!attr_cheb.done()) {
      TW_CUDA(cudaFuncSetAttribute(k_score_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 129 + 4 * 128) * (int)sizeof(float)));
      attr_cheb.mark();
    }
  }
  if (x.want_inputs) {
    TW_CUDA(cudaMemsetAsync(x.b.dxc, 0, (size_t)B * V * 3 * sizeof(float), x.st));
    TW_CUDA(cudaMemsetAsync(x.b.dxv, 0, (size_t)B * V * 3 * sizeof(float), x.st));
  }
  if ((x.dls || x.want_inputs) && !x.pv.local()) {
    TW_CHECK_ARG(cfg->num_heads * 128 <= (cfg->dim_feedforward > 256 ? cfg->dim_feedforward : 256),
                 "lengthscale gradient: H * 128 exceeds the [M, dim_feedforward] scratch");
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
      TW_CUDA(cudaFuncSetAttribute(k_score_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 129 + 4 * 128) * (int)sizeof(float)));
      attr_done.mark();
    }
    TW_CUDA(cudaMemsetAsync(x.b.sgrad, 0, (size_t)B * cfg->num_heads * V * V * sizeof(float), x.st));
  }
  return TW_OK;
}

static int finish_backward(BwdCtx& x) {
  if (x.want_inputs && !x.pv.local()) {
    k_score_coord_grad<<<(unsigned)x.B, 256, (size_t)x.V * 3 * sizeof(float), x.st>>>(x.tp.xc, x.mask, x.pv.enc(0, 0, 0, 1), x.b.sgrad, x.V,
                                                                                   x.c->num_heads, x.b.dxc);
    TW_LAUNCH_CHECK();
  }
  if (x.dls) {
    k_ls_grad<<<(unsigned)x.B, 256, 0, x.st>>>(x.tp.xc, x.mask, x.pv.enc(0, 0, 0, 1), x.b.sgrad, x.V, x.c->num_heads, x.dls);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

int tw_flow_log_likelihood_backward(const tw_flow_config* cfg, const void* const* params, void* const* grads,
                                    const int64_t* atom_types, const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V,
                                    const float* grad_log_prob, const void* packed_weights, void* tape, size_t tape_bytes,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  return tw_flow_log_likelihood_backward_inputs(cfg, params, grads, atom_types, x_velocs, mask, B, V, grad_log_prob, packed_weights, tape,
                                                tape_bytes, workspace, workspace_bytes, nullptr, nullptr, nullptr, nullptr, stream);
}

int tw_flow_log_likelihood_backward_inputs(const tw_flow_config* cfg, const void* const* params, void* const* grads,
                                           const int64_t* atom_types, const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V,
                                           const float* grad_log_prob, const void* packed_weights, void* tape, size_t tape_bytes,
                                           void* workspace, size_t workspace_bytes, float* out_grad_xc, float* out_grad_x_velocs,
                                           float* out_grad_z0_coords, float* out_grad_z0_velocs, void* stream) {
  TW_TRY(check_train(cfg, params, B, V));
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(grad_log_prob, "NULL pointer");
  const bool want_inputs = out_grad_xc || out_grad_x_velocs;
  TW_CHECK_ARG(!want_inputs || (out_grad_xc && out_grad_x_velocs), "conditioning gradients come as a pair (coords, velocs)");
  BwdCtx x{};
  x.want_inputs = want_inputs;
  TW_TRY(begin_backward(x, cfg, params, grads, atom_types, x_velocs, mask, B, V, packed_weights, tape, tape_bytes, workspace, workspace_bytes, stream));
  const int L = cfg->num_coupling_layers;
  k_prior_bwd<<<(unsigned)B, 128, 0, x.st>>>(x.tp.z[L][0], x.tp.z[L][1], mask, x.pv.log_scale_c(), x.pv.log_scale_v(), grad_log_prob, (int)V,
                                            x.b.dz[0], x.b.dz[1], x.gv.log_scale_c(), x.gv.log_scale_v());
  TW_LAUNCH_CHECK();
  for (int k = L - 1; k >= 0; k--) {
    const bool pos = (k % 2) == cfg->position_layer_index_mod_2;
    const int tgt = pos ? 0 : 1, oth = 1 - tgt;
    k_coupling_bwd<<<(unsigned)B, 128, 0, x.st>>>(x.tp.net[k][0].st, x.tp.z[k][tgt], x.b.dz[tgt], mask, grad_log_prob, (int)V, x.b.dst[0], x.b.dst[1]);
    TW_LAUNCH_CHECK();
    TW_TRY(conditioner_bwd(x, k, x.b.dz[oth], x.tp.z[k][oth]));
  }
  TW_TRY(finish_backward(x));
  const size_t bytes = (size_t)B * V * 3 * sizeof(float);
  if (want_inputs) {
    TW_CUDA(cudaMemcpyAsync(out_grad_xc, x.b.dxc, bytes, cudaMemcpyDeviceToDevice, x.st));
    TW_CUDA(cudaMemcpyAsync(out_grad_x_velocs, x.b.dxv, bytes, cudaMemcpyDeviceToDevice, x.st));
  }
  if (out_grad_z0_coords) TW_CUDA(cudaMemcpyAsync(out_grad_z0_coords, x.b.dz[0], bytes, cudaMemcpyDeviceToDevice, x.st));
  if (out_grad_z0_velocs) TW_CUDA(cudaMemcpyAsync(out_grad_z0_velocs, x.b.dz[1], bytes, cudaMemcpyDeviceToDevice, x.st));
  return TW_OK;
}

// Backward of tw_flow_sample_train: given d/dy_coords, d/dy_velocs [B,V,3] and d/d(delta) [B], accumulates the parameter
// gradients into `grads` and returns d/dz_coords, d/dz_velocs (the latent draws).  The layers are walked in the order
// 0 .. L-1, the reverse of the order the sampling pass applied them.
int tw_flow_sample_backward(const tw_flow_config* cfg, const void* const* params, void* const* grads, const int64_t* atom_types,
                            const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V, const float* grad_y_coords,
                            const float* grad_y_velocs, const float* grad_delta, const void* packed_weights, void* tape,
                            size_t tape_bytes, void* workspace, size_t workspace_bytes, float* out_grad_z_coords,
                            float* out_grad_z_velocs, void* stream) {
  TW_TRY(check_train(cfg, params, B, V));
  if (B == 0) return TW_OK;
  TW_CHECK_ARG(grad_y_coords && grad_y_velocs && grad_delta && out_grad_z_coords && out_grad_z_velocs, "NULL pointer");
  BwdCtx x{};
  TW_TRY(begin_backward(x, cfg, params, grads, atom_types, x_velocs, mask, B, V, packed_weights, tape, tape_bytes, workspace, workspace_bytes, stream));
  const int L = cfg->num_coupling_layers;
  const size_t bytes = (size_t)B * V * 3 * sizeof(float);
  TW_CUDA(cudaMemcpyAsync(x.b.dz[0], grad_y_coords, bytes, cudaMemcpyDeviceToDevice, x.st));  // y = (x +) z[0]
  TW_CUDA(cudaMemcpyAsync(x.b.dz[1], grad_y_velocs, bytes, cudaMemcpyDeviceToDevice, x.st));
  for (int k = 0; k < L; k++) {
    const bool pos = (k % 2) == cfg->position_layer_index_mod_2;
    const int tgt = pos ? 0 : 1, oth = 1 - tgt;
    k_coupling_rev_bwd<<<(unsigned)B, 128, 0, x.st>>>(x.tp.net[k][0].st, x.tp.z[k][tgt], x.b.dz[tgt], mask, grad_delta, (int)V, x.b.dst[0], x.b.dst[1]);
    TW_LAUNCH_CHECK();
    TW_TRY(conditioner_bwd(x, k, x.b.dz[oth], x.tp.z[k][oth]));
  }
  TW_CUDA(cudaMemcpyAsync(out_grad_z_coords, x.b.dz[0], bytes, cudaMemcpyDeviceToDevice, x.st));
  TW_CUDA(cudaMemcpyAsync(out_grad_z_velocs, x.b.dz[1], bytes, cudaMemcpyDeviceToDevice, x.st));
  return finish_backward(x);
}

}  // extern "C"

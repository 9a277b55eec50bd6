// Fused kernel-attention encoder sub-layer, FEATURE-MAJOR form (inference path):
//
//     out = LayerNorm1( x + sum_h A_h (x W_c,h^T) ),     W_c,h = W_o,h W_v,h   (pre-multiplied at pack time)
//
// (custom_attention_encoder.py:102-108, kernel_self_attention.py:29-48, kernel_attention.py:124-214: values are projected
// FIRST and mixed per sample afterwards, the order the reference itself uses.)  Features live on the 128 TMEM lanes and the
// tokens of a group of G samples on the MMA N axis, so no MMA row is padding whatever the atom count:
//
//   P(h):  PT[f_out, t]  = W_c,h[f_out, f_in] (A: packed weight image, smem, SW128 K-major)
//                          * X[t, f_in]       (B: the group's x rows as bf16 hi/lo K-major SW128 tiles, N = G*VP tokens)
//   conversion: PT fp32 -> bf16 hi | lo IN PLACE in TMEM (two column blocks)                      [epilogue warps]
//   M(h):  DT[f_out, i] += PT[f_out, j] (A: TMEM) * A_h[i, j] (B: the [VP x VP] K-major score image)   per sample, N = VP
//   drain: DT -> staging[token][feature] (transposition through conflict-free 4-byte shared-memory stores)
//   LayerNorm: one warp per token row, residual x re-read from global (L2), warp-shuffle statistics, coalesced float4 stores
//
// The previous fused kernel (k_attn_fused, flow_tc.cu) mixed first with the TOKENS of one sample on the lanes: 65 of 128 MMA
// rows real at 65 atoms, 39 M128xN128 MMAs per (sample, head).  Here a (sample, head) costs 24 MMAs of N = G*VP / G (projection)
// + 3 VP/16 MMAs of N = VP (mixing): 39 MMAs of N = 80 at 65 atoms -- 0.625 of the tensor work, and atom counts up to 128.
//
// Warp roles (480 threads): 0 W_c producer, 14 score-image producer (bulk async copies; two lanes of ONE warp starve each other
// while one of them spins on a full ring), 1 MMA issuer, 2-5 x-tile builders + LayerNorm, 6-13 two epilogue groups (column halves).  TMEM: PT0 | PT1 | DT0 | DT1, 128 columns each.
// MMA issue order over a global head counter g: P(g), M(g-1) -- the tensor pipe executes in issue order, so PT[g & 1] is not
// overwritten before M(g - 2) has read it, and the conversion of PT(g) overlaps M(g-1) + P(g+1).
#include <stdlib.h>

#include "flow_tc.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

constexpr int kFmThreads = 480;
constexpr int kFmWcStage = 16384;  // the hi or the lo image of one [128 out x 64 in] K block of W_c,h
constexpr uint32_t FM_PT = 0, FM_DT = 256;

struct FmArgs {
  const float* x[2];
  float* out[2];
  const uint8_t* scores_img;
  const uint8_t* wc[2];
  const float* gamma[2];
  const float* beta[2];
  int64_t n, n_cond;
  int V, VP, H, G;  // G samples per group, N = G * VP tokens on the MMA N axis
  int wc_stages, sc_stages;
  float eps;
  long long* trace;
  int exp;  // timing experiments (TW_FM_EXP, bring-up only; results are wrong when set)
};

struct FmSmem {
  uint32_t N, xb_bytes, sc_unit, stg_bytes, wc_stages, sc_stages;
  __host__ __device__ FmSmem(int V, int VP, int G, int wcs, int scs) {
    N = (uint32_t)(G * VP);
    xb_bytes = N * 512u;                       // hi kb0 | hi kb1 | lo kb0 | lo kb1, each [N x 128 B]
    sc_unit = (uint32_t)(2 * VP * VP * 2);     // hi | lo image of one (sample, head)
    stg_bytes = (uint32_t)(G * V) * 512u;      // [token][128 fp32]
    wc_stages = (uint32_t)wcs, sc_stages = (uint32_t)scs;
  }
  __host__ __device__ uint32_t xb(int b) const { return (uint32_t)b * xb_bytes; }
  __host__ __device__ uint32_t wc() const { return 2 * xb_bytes; }
  __host__ __device__ uint32_t sc() const { return wc() + wc_stages * kFmWcStage; }
  __host__ __device__ uint32_t stg() const { return sc() + sc_stages * sc_unit; }
  __host__ __device__ uint32_t vec() const { return stg() + stg_bytes; }
  __host__ __device__ uint32_t bars() const { return vec(); }
  __host__ __device__ uint32_t total() const { return bars() + 384; }
};

// mbarrier wait with a watchdog: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void fm_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("k_attn_fm: barrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fm_epi_bar() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

template <int kSplit>
__global__ void __launch_bounds__(kFmThreads, 1) k_attn_fm(FmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // no static shared memory in this kernel: the dynamic window is 1024-byte aligned (checked below)
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V, VP = a.VP, H = a.H, G = a.G;
  const FmSmem L(V, VP, G, a.wc_stages, a.sc_stages);
  const int N = (int)L.N;
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  constexpr int kParts = kSplit == 3 ? 2 : 1;
  const int64_t n_groups_total = (a.n + G - 1) / G;
  const int64_t my_groups = ((int64_t)blockIdx.x < n_groups_total) ? (n_groups_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto group_of = [&](int64_t it) -> int64_t { return blockIdx.x + it * gridDim.x; };
  auto samples_in = [&](int64_t grp) -> int { int64_t r = a.n - grp * G; return (int)(r < G ? r : G); };

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars());
  uint64_t* wc_full = bars;                 // [8]
  uint64_t* wc_empty = wc_full + 8;         // [8]
  uint64_t* sc_full = wc_empty + 8;         // [4]
  uint64_t* sc_empty = sc_full + 4;         // [4]
  uint64_t* xb_full = sc_empty + 4;         // [2] 128 arrivals: x tiles of a group written
  uint64_t* xb_free = xb_full + 2;          // [2] commit: the last projection MMA that reads the tiles retired
  uint64_t* pt_full = xb_free + 2;          // [2] commit: P(g) retired
  uint64_t* h_full = pt_full + 2;           // [2] 256 arrivals: PT(g) converted in place
  uint64_t* dt_full = h_full + 2;           // [2] commit: last M of the group retired
  uint64_t* dt_free = dt_full + 2;          // [2] 256 arrivals: DT drained
  uint64_t* stg_full = dt_free + 2;         // 256 arrivals: staging rows written
  uint64_t* stg_free = stg_full + 1;        // 128 arrivals: LayerNorm read the staging rows
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_free + 1);

  if (tid == 0) {
    for (int i = 0; i < 8; i++) mbar_init(&wc_full[i], 1), mbar_init(&wc_empty[i], 1);
    for (int i = 0; i < 4; i++) mbar_init(&sc_full[i], 1), mbar_init(&sc_empty[i], 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&xb_full[i], 128), mbar_init(&xb_free[i], 1), mbar_init(&pt_full[i], 1), mbar_init(&h_full[i], 256);
      mbar_init(&dt_full[i], 1), mbar_init(&dt_free[i], 256);
    }
    mbar_init(stg_full, 256), mbar_init(stg_free, 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ksteps = VP / 16;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define FM_TRACE(role, ev, item)                                                          \
  if (tr_on && tr_n < 1024) {                                                             \
    a.trace[((role) * 1024 + tr_n) * 2] = (long long)(ev) | ((long long)(item) << 8);     \
    a.trace[((role) * 1024 + tr_n) * 2 + 1] = clock64();                                  \
    tr_n++;                                                                               \
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: W_c units
    if (lane == 0) {
      uint32_t ws = 0, wp = 0;
      const int64_t total_heads = my_groups * H;
      for (int64_t g = 0, h = 0; g < total_heads; g++, h = (h + 1 == H ? 0 : h + 1)) {
        for (int u = 0; u < 2 * kParts; u++) {  // K block 0: hi, lo; K block 1: hi, lo
          const int kb = u / kParts, part = u % kParts;
          fm_wait(&wc_empty[ws], wp ^ 1);
          mbar_arrive_expect_tx(&wc_full[ws], (uint32_t)kFmWcStage);
          bulk_g2s(smem + L.wc() + ws * kFmWcStage, a.wc[net] + (size_t)(h * 2 + kb) * 32768 + part * 16384, kFmWcStage, &wc_full[ws]);
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ------------------------------------------------------------------ producer: score images, one unit per (sample, head)
    if (lane == 0) {
      uint32_t ss = 0, sp = 0;
      for (int64_t it = 0; it < my_groups; it++) {
        const int64_t grp = group_of(it);
        const int ns = samples_in(grp);
        for (int h = 0; h < H; h++) {
          for (int s = 0; s < ns; s++) {
            const int64_t n = grp * G + s;
            const uint8_t* src = a.scores_img + ((size_t)(a.n_cond == a.n ? n : n % a.n_cond) * H + h) * (2 * (size_t)mat_bytes);
            fm_wait(&sc_empty[ss], sp ^ 1);
            mbar_arrive_expect_tx(&sc_full[ss], kParts * mat_bytes);
            bulk_g2s(smem + L.sc() + ss * L.sc_unit, src, kParts * mat_bytes, &sc_full[ss]);
            if (++ss == L.sc_stages) ss = 0, sp ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t ws = 0, wp = 0, ss = 0, sp = 0;
    uint32_t ph_xb = 0, ph_h = 0, ph_dtfree = 0;  // bit b = phase of buffer b
    const uint32_t idescP = make_idesc_bf16(128, (uint32_t)N, 0, 0);
    const uint32_t idescM = make_idesc_bf16(128, (uint32_t)VP, 0, 0);
    const uint32_t sc_sbo = (uint32_t)(VP >> 3) * 128;
    const uint32_t blk = (uint32_t)N * 128u;  // one [N x 64] K block of the x tiles

    auto issue_P = [&](int64_t g, int64_t it, int h) {  // PT[g & 1] = W_c,h X^T  (24 MMAs with the bf16x3 split)
      const uint32_t d = tmem + FM_PT + (uint32_t)(g & 1) * 128;
      const uint32_t xt = smem_u32(smem + L.xb((int)(it & 1)));
      for (int kb = 0; kb < 2; kb++) {
        const uint32_t x_hi = xt + kb * blk, x_lo = xt + 2 * blk + kb * blk;
        if (!(a.exp & 1)) fm_wait(&wc_full[ws], wp);  // hi image of this K block
        tc_fence_after();
        if (kb == 0) { FM_TRACE(0, 3, g); }
        if (elect_one()) {
          const uint32_t w = smem_u32(smem + L.wc() + ws * kFmWcStage);
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(w + k * 32), desc_kmajor_sw128(x_hi + k * 32), idescP, (kb | k) != 0);
          if (kSplit == 3) {
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(w + k * 32), desc_kmajor_sw128(x_lo + k * 32), idescP, 1);
          }
          mma_commit(&wc_empty[ws]);
          if (kSplit != 3 && kb == 1) {
            mma_commit(&pt_full[g & 1]);
            if (h == H - 1) mma_commit(&xb_free[it & 1]);
          }
        }
        __syncwarp();
        if (++ws == L.wc_stages) ws = 0, wp ^= 1;
        if (kSplit == 3) {
          if (!(a.exp & 1)) fm_wait(&wc_full[ws], wp);  // lo image
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w = smem_u32(smem + L.wc() + ws * kFmWcStage);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(w + k * 32), desc_kmajor_sw128(x_hi + k * 32), idescP, 1);
            mma_commit(&wc_empty[ws]);
            if (kb == 1) {
              mma_commit(&pt_full[g & 1]);
              if (h == H - 1) mma_commit(&xb_free[it & 1]);
            }
          }
          __syncwarp();
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
        }
      }
    };
    auto issue_M = [&](int64_t g, int64_t it, int h) {  // DT[it & 1][:, sample s] += PT(g)[:, sample s] A_h(s)^T
      const int b = (int)(g & 1), db = (int)(it & 1);
      const int ns = samples_in(group_of(it));
      if (h == 0 && it >= 2) {  // the accumulator was last used two groups ago: drained?
        fm_wait(&dt_free[db], (ph_dtfree >> db) & 1u);
        ph_dtfree ^= 1u << db;
      }
      if (!(a.exp & 2)) fm_wait(&h_full[b], (ph_h >> b) & 1u);
      ph_h ^= 1u << b;
      tc_fence_after();
      FM_TRACE(0, 4, g);
      for (int s = 0; s < ns; s++) {
        if (!(a.exp & 2)) fm_wait(&sc_full[ss], sp);
        tc_fence_after();
        if (s == 0) { FM_TRACE(0, 5, g); }
        if (elect_one()) {
          const uint32_t s_hi = smem_u32(smem + L.sc() + ss * L.sc_unit), s_lo = s_hi + mat_bytes;
          const uint32_t d = tmem + FM_DT + (uint32_t)db * 128 + (uint32_t)(s * VP);
          const uint32_t p_hi = tmem + FM_PT + (uint32_t)b * 128 + (uint32_t)(s * (VP >> 1)), p_lo = p_hi + (uint32_t)(N >> 1);
          for (int k = 0; k < ksteps; k++)
            mma_ts(d, p_hi + k * 8, make_smem_desc(s_hi + k * 256, 128, sc_sbo, LAYOUT_NONE), idescM, (h | k) != 0);
          if (kSplit == 3) {
            for (int k = 0; k < ksteps; k++) mma_ts(d, p_lo + k * 8, make_smem_desc(s_hi + k * 256, 128, sc_sbo, LAYOUT_NONE), idescM, 1);
            for (int k = 0; k < ksteps; k++) mma_ts(d, p_hi + k * 8, make_smem_desc(s_lo + k * 256, 128, sc_sbo, LAYOUT_NONE), idescM, 1);
          }
          mma_commit(&sc_empty[ss]);
          if (h == H - 1 && s == ns - 1) mma_commit(&dt_full[db]);
        }
        __syncwarp();
        if (++ss == L.sc_stages) ss = 0, sp ^= 1;
      }
    };

    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      fm_wait(&xb_full[it & 1], (ph_xb >> (it & 1)) & 1u);
      ph_xb ^= 1u << (it & 1);
      tc_fence_after();
      for (int h = 0; h < H; h++, g++) {
        FM_TRACE(0, 0, g);
        issue_P(g, it, h);
        FM_TRACE(0, 1, g);
        if (g >= 1) issue_M(g - 1, h > 0 ? it : it - 1, h > 0 ? h - 1 : H - 1);
        FM_TRACE(0, 2, g);
      }
    }
    if (g > 0) issue_M(g - 1, my_groups - 1, H - 1);
  } else if (warp < 6) {
    // ------------------------------------------------------------------ x-tile builders + LayerNorm (128 threads)
    const int lt = tid - 64;          // 0..127
    const int lw = lt >> 5;           // 0..3
    const int hw = lane >> 4;         // half-warp: row parity
    const int c = lane & 15;          // 16-byte chunk of a row's bf16 image = 8 features
    uint32_t ph_free = 0, ph_stg = 0;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma[net]) + lane), bt = __ldg(reinterpret_cast<const float4*>(a.beta[net]) + lane);

    auto build_tiles = [&](int64_t it) {  // group it -> xb[it & 1]: bf16 hi/lo K-major SW128 rows of the group's tokens
      const int b = (int)(it & 1);
      if (it >= 2) {
        fm_wait(&xb_free[b], (ph_free >> b) & 1u);
        ph_free ^= 1u << b;
      }
      const int64_t grp = group_of(it);
      const int ns = samples_in(grp);
      uint8_t* tile = smem + L.xb(b);
      const uint32_t blk = (uint32_t)N * 128u;
      const int kb = c >> 3, cc = c & 7;
      for (int r0 = 0; r0 < ((a.exp & 8) ? 0 : N); r0 += 32) {  // 8 rows per pass of the 4 warps, 4 passes in flight
        float4 v[4][2];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const int r = r0 + p * 8 + lw * 2 + hw;
          const int s = r / VP, at = r - s * VP;
          v[p][0] = v[p][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < N && s < ns && at < V) {
            const float4* src = reinterpret_cast<const float4*>(a.x[net] + ((grp * G + s) * V + at) * 128 + c * 8);
            v[p][0] = __ldg(src), v[p][1] = __ldg(src + 1);
          }
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const int r = r0 + p * 8 + lw * 2 + hw;
          if (r < N) {
            uint32_t hi[4], lo[4];
            split2(v[p][0].x, v[p][0].y, hi[0], lo[0]);
            split2(v[p][0].z, v[p][0].w, hi[1], lo[1]);
            split2(v[p][1].x, v[p][1].y, hi[2], lo[2]);
            split2(v[p][1].z, v[p][1].w, hi[3], lo[3]);
            const uint32_t off = kb * blk + (uint32_t)r * 128u + (((uint32_t)cc ^ ((uint32_t)r & 7u)) << 4);
            *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (kSplit == 3) *reinterpret_cast<uint4*>(tile + 2 * blk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&xb_full[b]);
    };
    auto layer_norm = [&](int64_t it) {  // staging rows of group it: + x, LayerNorm, store
      fm_wait(stg_full, ph_stg);
      ph_stg ^= 1;
      const int64_t grp = group_of(it);
      const int rows = samples_in(grp) * V;
      const float* xg = a.x[net] + grp * G * V * 128;
      float* og = a.out[net] + grp * G * V * 128;
      const uint8_t* stg = smem + L.stg();
      for (int r0 = lw; r0 < ((a.exp & 8) ? 0 : rows); r0 += 16) {  // 4 rows per warp in flight
        float4 xv[4], sv[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const int r = r0 + 4 * p;
          if (r < rows) {
            xv[p] = __ldg(reinterpret_cast<const float4*>(xg + (size_t)r * 128) + lane);
            sv[p] = *reinterpret_cast<const float4*>(stg + (size_t)r * 512 + lane * 16);
          }
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const int r = r0 + 4 * p;
          if (r < rows) {  // (warp-uniform)
            const float y0 = xv[p].x + sv[p].x, y1 = xv[p].y + sv[p].y, y2 = xv[p].z + sv[p].z, y3 = xv[p].w + sv[p].w;
            float sum = (y0 + y1) + (y2 + y3);
            float sq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, y3 * y3)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              sum += __shfl_xor_sync(0xffffffffu, sum, o);
              sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            const float mean = sum * (1.f / 128.f);
            const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
            const float rstd = 1.0f / sqrtf(var + a.eps);
            float4 o4;
            o4.x = (y0 - mean) * rstd * gm.x + bt.x;
            o4.y = (y1 - mean) * rstd * gm.y + bt.y;
            o4.z = (y2 - mean) * rstd * gm.z + bt.z;
            o4.w = (y3 - mean) * rstd * gm.w + bt.w;
            *(reinterpret_cast<float4*>(og + (size_t)r * 128) + lane) = o4;
          }
        }
      }
      mbar_arrive(stg_free);
    };

    if (my_groups > 0) build_tiles(0);
    for (int64_t it = 0; it < my_groups; it++) {
      if (it + 1 < my_groups) build_tiles(it + 1);
      if (it >= 1) layer_norm(it - 1);
    }
    if (my_groups > 0) layer_norm(my_groups - 1);
  } else {
    // ------------------------------------------------------------------ epilogue groups (column halves of the N tokens)
    const int q = warp & 3;                 // TMEM lane quarter of this warp (warps 6..13: 6&3 = 2, ...)
    const int e = (warp - 6) >> 2;          // column half
    const int f = q * 32 + lane;            // feature = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int half = N >> 1;                // fp32 columns per group (multiple of 8)
    const int nchunk = half >> 3;           // chunks of 8 columns (<= 8)
    uint32_t ph_pt = 0, ph_dt = 0, ph_stgfree = 0;
    float* stg = reinterpret_cast<float*>(smem + L.stg());

    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 1);
        fm_wait(&pt_full[b], (ph_pt >> b) & 1u);
        ph_pt ^= 1u << b;
        tc_fence_after();
        if (q == 2 && e == 0) { FM_TRACE(1, 0, g); }
        const uint32_t base = tmem + lane_base + FM_PT + (uint32_t)b * 128;
        uint32_t r[8][8];
        const int nchunk_c = (a.exp & 4) ? 0 : nchunk;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i < nchunk_c) tmem_ld8(base + (uint32_t)(e * half + 8 * i), r[i]);
        tmem_ld_wait();
        fm_epi_bar();  // both halves have read their fp32 columns: the in-place writes below may cross into the other half
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i < nchunk_c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; j++) split2(__uint_as_float(r[i][2 * j]), __uint_as_float(r[i][2 * j + 1]), hi[j], lo[j]);
            const uint32_t col = (uint32_t)((e * half + 8 * i) >> 1);  // packed column of tokens (8i, 8i+1)
            tmem_st4(base + col, hi[0], hi[1], hi[2], hi[3]);
            if (kSplit == 3) tmem_st4(base + (uint32_t)half + col, lo[0], lo[1], lo[2], lo[3]);
          }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&h_full[b]);
        if (q == 2 && e == 0) { FM_TRACE(1, 1, g); }
        if (h == 0 && it > 0) {
          // ---- drain the previous group's accumulator into the staging rows (transposition: lanes write consecutive features)
          const int64_t pit = it - 1;
          const int db = (int)(pit & 1);
          fm_wait(&dt_full[db], (ph_dt >> db) & 1u);
          ph_dt ^= 1u << db;
          tc_fence_after();
          if (pit >= 1) {
            fm_wait(stg_free, ph_stgfree);
            ph_stgfree ^= 1;
          }
          const int ns = samples_in(group_of(pit));
          const uint32_t dbase = tmem + lane_base + FM_DT + (uint32_t)db * 128;
#pragma unroll 1
          for (int i = 0; i < nchunk; i++) {
            uint32_t v[8];
            tmem_ld8(dbase + (uint32_t)(e * half + 8 * i), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const int t = e * half + 8 * i + j;
              const int s = t / VP, at = t - s * VP;
              if (s < ns && at < V) stg[(size_t)(s * V + at) * 128 + f] = __uint_as_float(v[j]);
            }
          }
          tc_fence_before();
          mbar_arrive(&dt_free[db]);
          mbar_arrive(stg_full);
          if (q == 2 && e == 0) { FM_TRACE(1, 2, g); }
        }
      }
    }
    if (my_groups > 0) {  // drain of the last group
      const int64_t pit = my_groups - 1;
      const int db = (int)(pit & 1);
      fm_wait(&dt_full[db], (ph_dt >> db) & 1u);
      tc_fence_after();
      if (pit >= 1) fm_wait(stg_free, ph_stgfree);
      const int ns = samples_in(group_of(pit));
      const uint32_t dbase = tmem + lane_base + FM_DT + (uint32_t)db * 128;
#pragma unroll 1
      for (int i = 0; i < nchunk; i++) {
        uint32_t v[8];
        tmem_ld8(dbase + (uint32_t)(e * half + 8 * i), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int t = e * half + 8 * i + j;
          const int s = t / VP, at = t - s * VP;
          if (s < ns && at < V) stg[(size_t)(s * V + at) * 128 + f] = __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      mbar_arrive(stg_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef FM_TRACE
}

// -------------------------------------------------------------------------------------------- host side
static long long* g_fm_trace = nullptr;
void tc_set_fm_trace(long long* buf) { g_fm_trace = buf; }

// Group size and ring depths for an atom count; false if the kernel's buffers do not fit (the caller falls back).
static bool fm_plan(int V, int VP, int64_t n, int* G, int* wcs, int* scs, int* smem_bytes) {
  if (VP > 128 || VP < 16) return false;
  int g = 128 / VP;                       // N = G * VP <= 128 TMEM columns per buffer
  while (g > 1 && g * VP * 512 * 2 > 96 * 1024) g--;  // both x-tile buffers within 96 KB
  if (g < 1) g = 1;
  if ((int64_t)g > n) g = (int)(n < 1 ? 1 : n);
  for (int w = 4; w >= 3; w--)
    for (int s = (g > 1 ? 4 : 2); s >= 2; s--) {
      const int total = (int)FmSmem(V, VP, g, w, s).total();
      if (total <= 232448) {
        *G = g, *wcs = w, *scs = s, *smem_bytes = total;
        return true;
      }
    }
  return false;
}

bool tc_attn_fm_supported(int V, int64_t n) {
  static int use = -1;
  if (use < 0) {
    const char* e = getenv("TW_ATTN_FM");  // bring-up switch: 0 = the token-major fused kernel / two-kernel form
    use = e ? atoi(e) : 1;
  }
  int G, w, s, b;
  return use && n >= 1 && fm_plan(V, (V + 15) / 16 * 16, n, &G, &w, &s, &b);
}

int tc_attn_fm(const tw_flow_config* c, const float* const x[2], float* const out[2], const uint8_t* const wc[2],
               const float* const gamma[2], const float* const beta[2], const uint8_t* scores_img, int64_t n, int64_t n_cond, int V,
               int nets, cudaStream_t st) {
  const int VP = (V + 15) / 16 * 16;
  FmArgs a{};
  int smem_bytes = 0;
  TW_CHECK_ARG(fm_plan(V, VP, n, &a.G, &a.wc_stages, &a.sc_stages, &smem_bytes), "feature-major attention: atom count out of range");
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_attn_fm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    TW_CUDA(cudaFuncSetAttribute(k_attn_fm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done.mark();
  }
  for (int s = 0; s < 2; s++) a.x[s] = x[s], a.out[s] = out[s], a.wc[s] = wc[s], a.gamma[s] = gamma[s], a.beta[s] = beta[s];
  a.scores_img = scores_img;
  a.n = n, a.n_cond = n_cond, a.V = V, a.VP = VP, a.H = c->num_heads, a.eps = c->layer_norm_eps;
  a.trace = g_fm_trace;
  {
    const char* e = getenv("TW_FM_EXP");
    a.exp = e ? atoi(e) : 0;
  }
  const int per_net = nets == 1 ? 148 : 74;
  const int64_t groups = (n + a.G - 1) / a.G;
  dim3 grid((unsigned)(groups < per_net ? groups : per_net), nets);
  if (c->precision == TW_PRECISION_BF16X3)
    k_attn_fm<3><<<grid, kFmThreads, smem_bytes, st>>>(a);
  else
    k_attn_fm<1><<<grid, kFmThreads, smem_bytes, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // namespace tw

// Fused kernel-attention encoder sub-layer, feature-major, DEEP pipeline (groups of at most 80 tokens):
//
//     out = LayerNorm1( x + sum_h A_h (x W_c,h^T) ),     W_c,h = W_o,h W_v,h   (pre-multiplied at pack time)
//
// Same arithmetic and operand layouts as k_attn_fm (attn_fm.cu: projection P(h) with the 128 output features on the TMEM lanes and
// the tokens on the MMA N axis, conversion fp32 -> bf16 hi | lo in place in TMEM, mixing M(h) with the A operand from TMEM).  What
// k_attn_fm's in-kernel trace showed is a kernel bound by its hand-overs, not by the tensor pipe (41 % busy): with N = 160 tokens
// TMEM holds two projection buffers and ONE accumulator, shared memory ONE set of x tiles, so the chain P -> conversion -> M -> P(+2),
// the tile rebuild and the accumulator drain all sit on the critical path.  Here a group is at most 80 tokens (one sample at 65..80
// atoms), which buys, in the same TMEM and shared memory:
//     4 projection buffers   (the P issuer runs up to four heads ahead of the mixing),
//     2 accumulators         (the drain of group g overlaps the mixing of group g + 1),
//     2 sets of x tiles      (the rebuild for group g + 2 overlaps groups g and g + 1; rows come in with plain loads),
// at the price of N = 80 instead of 160 in the SS-form projection MMAs (52.9 instead of 40.5 cycles per sample and MMA: +19 % tensor
// work) and twice the W_c traffic per sample (L2: 22 % -> about 50 % of its peak).
//
// Warp roles (640 threads): 0 producer (W_c ring, score-image ring: bulk async copies), 1 MMA issuer for the projections, 2 MMA issuer
// for the mixing, 3 idle, 4-11 two epilogue groups (column halves of the N tokens: conversion, accumulator drain), 12-19 x-tile
// builders + LayerNorm.  TMEM: PT0 | PT1 | PT2 | PT3 | DT0 | DT1, N <= 80 columns each.
#include <stdlib.h>

#include "flow_tc.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

constexpr int kF3Threads = 640;
constexpr int kF3WcStage = 32768;  // the hi and the lo image (16 KB each) of one [128 out x 64 in] K block of W_c,h
constexpr int kF3MaxN = 80;        // tokens of a group: 6 N <= 512 TMEM columns
constexpr int kF3Pt = 4, kF3Dt = 2, kF3Xb = 2;

struct F3Args {
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
  int dbg;  // bring-up: 1 = every CTA loads whole W_c units itself (no multicast), pair release protocol unchanged
};

struct F3Smem {
  uint32_t N, xb_bytes, sc_unit, wc_stages, sc_stages;
  __host__ __device__ F3Smem(int VP, int G, int wcs, int scs) {
    N = (uint32_t)(G * VP);
    xb_bytes = N * 512u;                       // hi kb0 | hi kb1 | lo kb0 | lo kb1, each [N x 128 B]
    sc_unit = (uint32_t)(2 * VP * VP * 2);     // hi | lo image of one (sample, head)
    wc_stages = (uint32_t)wcs, sc_stages = (uint32_t)scs;
  }
  __host__ __device__ uint32_t xb(int i) const { return (uint32_t)i * xb_bytes; }
  __host__ __device__ uint32_t wc() const { return kF3Xb * xb_bytes; }
  __host__ __device__ uint32_t sc() const { return wc() + wc_stages * kF3WcStage; }
  __host__ __device__ uint32_t bars() const { return sc() + sc_stages * sc_unit; }
  __host__ __device__ uint32_t total() const { return bars() + 512; }
};

// mbarrier wait with a watchdog: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void f3_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // (no printf: a call in the wait loop makes the compiler save live registers around it)
  }
}
__device__ __forceinline__ void f3_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void f3_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void f3_epi_bar() { asm volatile("bar.sync 3, 256;" ::: "memory"); }
// all previously issued MMAs of this thread complete -> arrive on the mbarrier at this offset in EVERY CTA of the pair
__device__ __forceinline__ void f3_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// bulk copy global -> the same shared-memory offset in both CTAs of the pair; each destination's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void f3_bulk_g2s_pair(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"((uint16_t)3), "l"(policy)
      : "memory");
}

template <int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF3Threads, 1) k_attn_fm3(F3Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // no static shared memory in this kernel: the dynamic window is 1024-byte aligned (checked below)
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V, VP = a.VP, H = a.H, G = a.G;
  const F3Smem L(VP, G, a.wc_stages, a.sc_stages);
  const int N = (int)L.N;
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  constexpr int kParts = kSplit == 3 ? 2 : 1;
  const int64_t n_groups_total = (a.n + G - 1) / G;
  const int64_t my_groups = ((int64_t)blockIdx.x < n_groups_total) ? (n_groups_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // The two CTAs of a cluster share ONE W_c stream (each loads half of every ring unit and multicasts it to both: half the
  // L2 -> SM traffic of the dominant operand), so both walk the same number of (group, head) steps: a CTA that has one group less
  // than the longest schedule consumes and releases that group's units without issuing MMAs.
  const int64_t w_iters = (n_groups_total + gridDim.x - 1) / gridDim.x;
  const uint32_t rank = cluster_ctarank();
  auto group_of = [&](int64_t it) -> int64_t { return blockIdx.x + it * gridDim.x; };
  auto samples_in = [&](int64_t grp) -> int { int64_t r = a.n - grp * G; return (int)(r < G ? r : G); };
  const uint32_t TM_DT = (uint32_t)(kF3Pt * N);  // PT0..PT3 | DT0 | DT1

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars());
  uint64_t* wc_full = bars;                 // [8]
  uint64_t* wc_empty = wc_full + 8;         // [8]
  uint64_t* sc_full = wc_empty + 8;         // [4]
  uint64_t* sc_empty = sc_full + 4;         // [4]
  uint64_t* xb_full = sc_empty + 4;         // [2] 256 arrivals: x tiles of a group written
  uint64_t* xb_free = xb_full + 2;          // [2] commit: the last projection MMA of a group retired
  uint64_t* pt_full = xb_free + 2;          // [4] commit: P(g) retired
  uint64_t* h_full = pt_full + 4;           // [4] 256 arrivals: PT(g) converted in place
  uint64_t* pt_free = h_full + 4;           // [4] commit: M(g) retired
  uint64_t* dt_full = pt_free + 4;          // [2] commit: last M of the group retired
  uint64_t* dt_free = dt_full + 2;          // [2] 256 arrivals: DT read out
  uint64_t* rows_out = dt_free + 2;         // [2] 256 arrivals: the group's pre-LayerNorm rows are in global memory
  uint64_t* ln_done = rows_out + 2;         // [2] 256 arrivals: LayerNorm of the group finished (keeps rows_out one phase ahead at most)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_done + 2);

  if (tid == 0) {
    for (int i = 0; i < 8; i++) mbar_init(&wc_full[i], 1), mbar_init(&wc_empty[i], 2);  // (released by both CTAs of the pair)
    for (int i = 0; i < 4; i++) mbar_init(&sc_full[i], 1), mbar_init(&sc_empty[i], 1), mbar_init(&pt_full[i], 1), mbar_init(&h_full[i], 256), mbar_init(&pt_free[i], 1);
    for (int i = 0; i < 2; i++)
      mbar_init(&xb_full[i], 256), mbar_init(&xb_free[i], 1), mbar_init(&dt_full[i], 1), mbar_init(&dt_free[i], 256), mbar_init(&rows_out[i], 256),
          mbar_init(&ln_done[i], 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (uint32_t i = tid * 16; i < kF3Xb * L.xb_bytes; i += kF3Threads * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);  // padding rows stay zero
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any multicast copy / commit
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ksteps = VP / 16;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define F3_TRACE(role, ev, item)                                                          \
  if (tr_on && tr_n < 1024) {                                                             \
    a.trace[((role) * 1024 + tr_n) * 2] = (long long)(ev) | ((long long)(item) << 8);     \
    a.trace[((role) * 1024 + tr_n) * 2 + 1] = clock64();                                  \
    tr_n++;                                                                               \
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: ONE thread feeds the W_c and score-image rings
    if (lane == 0) {
      uint32_t ws = 0, wp = 0;
      int64_t w_left = w_iters * H * 2;  // per head: K block 0, K block 1 (hi | lo images)
      int w_h = 0, w_u = 0;
      uint32_t ss = 0, sp = 0;
      int64_t s_it = 0;
      int s_h = 0, s_s = 0;
      bool s_done = my_groups == 0;
      long long t_idle = 0;
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      while (w_left > 0 || !s_done) {
        bool progress = false;
        if (w_left > 0 && mbar_test_wait(&wc_empty[ws], wp ^ 1)) {
          // each CTA of the pair copies half of the unit into both shared memories: with the 3-way split rank 0 the hi image and
          // rank 1 the lo image (16 KB each), otherwise rows [64 rank, 64 rank + 64) of the hi image
          const uint32_t unit = kSplit == 3 ? 32768u : 16384u, half_b = unit / 2;
          const uint8_t* src = a.wc[net] + (size_t)(w_h * 2 + w_u) * 32768;
          mbar_arrive_expect_tx(&wc_full[ws], unit);  // (own half + the peer's half)
          f3_bulk_g2s_pair(smem + L.wc() + ws * kF3WcStage + rank * half_b, src + rank * half_b, half_b, &wc_full[ws], pol_keep);
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
          if (++w_u == 2) {
            w_u = 0;
            if (++w_h == H) w_h = 0;
          }
          w_left--;
          progress = true;
        }
        if (!s_done && mbar_test_wait(&sc_empty[ss], sp ^ 1)) {
          const int64_t grp = group_of(s_it);
          const int64_t n = grp * G + s_s;
          const uint8_t* src = a.scores_img + ((size_t)(a.n_cond == a.n ? n : n % a.n_cond) * H + s_h) * (2 * (size_t)mat_bytes);
          mbar_arrive_expect_tx(&sc_full[ss], kParts * mat_bytes);
          bulk_g2s_hint(smem + L.sc() + ss * L.sc_unit, src, kParts * mat_bytes, &sc_full[ss], a.n_cond == a.n ? pol_stream : pol_keep);
          if (++ss == L.sc_stages) ss = 0, sp ^= 1;
          if (++s_s == samples_in(grp)) {
            s_s = 0;
            if (++s_h == H) {
              s_h = 0;
              if (++s_it == my_groups) s_done = true;
            }
          }
          progress = true;
        }
        if (progress) {
          t_idle = 0;
        } else {
          if (t_idle == 0) t_idle = clock64();
          else if (clock64() - t_idle > 4000000000LL) __trap();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer 1: projections  PT[g & 3] = W_c,h X^T
    uint32_t ws = 0, wp = 0;
    const uint32_t idescP = make_idesc_bf16(128, (uint32_t)N, 0, 0);
    const uint32_t blk = (uint32_t)N * 128u;  // one [N x 64] K block of the x tiles
    const uint32_t wring = smem_u32(smem + L.wc());
    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      const int xbuf = (int)(it & 1);
      f3_wait(&xb_full[xbuf], (uint32_t)((it >> 1) & 1));
      const uint32_t xt = smem_u32(smem + L.xb(xbuf));
      const uint64_t xd_hi0 = desc_kmajor_sw128(xt), xd_hi1 = desc_kmajor_sw128(xt + blk);
      const uint64_t xd_lo0 = desc_kmajor_sw128(xt + 2 * blk), xd_lo1 = desc_kmajor_sw128(xt + 3 * blk);
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 3);
        if (g >= kF3Pt) f3_wait(&pt_free[b], (uint32_t)(((g >> 2) - 1) & 1));
        F3_TRACE(0, 0, g);
        const uint32_t d = tmem + (uint32_t)b * (uint32_t)N;
#pragma unroll
        for (int kb = 0; kb < 2; kb++) {  // one ring unit per K block: W hi | W lo
          const uint64_t xh = kb ? xd_hi1 : xd_hi0, xl = kb ? xd_lo1 : xd_lo0;
          f3_wait(&wc_full[ws], wp);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t wd = desc_kmajor_sw128(wring + ws * kF3WcStage), wl = desc_kmajor_sw128(wring + ws * kF3WcStage + 16384);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(d, wd + 2 * k, xh + 2 * k, idescP, (kb | k) != 0);
            if (kSplit == 3) {
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, wd + 2 * k, xl + 2 * k, idescP, 1);
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, wl + 2 * k, xh + 2 * k, idescP, 1);
            }
            f3_commit_pair(&wc_empty[ws]);
            if (kb == 1) {
              mma_commit(&pt_full[b]);
              if (h == H - 1) mma_commit(&xb_free[xbuf]);
            }
          }
          __syncwarp();
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
        }
        F3_TRACE(0, 1, g);
      }
    }
    for (int64_t u = (w_iters - my_groups) * H * 2; u > 0; u--) {  // the peer's extra group: release its W_c units unused
      f3_wait(&wc_full[ws], wp);
      if (elect_one()) mbar_arrive_cluster(&wc_empty[ws], 0), mbar_arrive_cluster(&wc_empty[ws], 1);
      __syncwarp();
      if (++ws == L.wc_stages) ws = 0, wp ^= 1;
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer 2: mixing  DT[it & 1][:, sample s] += PT(g)[:, sample s] A_h(s)^T
    uint32_t ss = 0, sp = 0;
    const uint32_t idescM = make_idesc_bf16(128, (uint32_t)VP, 0, 0);
    const uint32_t sc_sbo = (uint32_t)(VP >> 3) * 128;
    const uint32_t sring = smem_u32(smem + L.sc());
    const uint64_t sd0 = make_smem_desc(0, 128, sc_sbo, LAYOUT_NONE);  // + (address >> 4) of the image, + 16 per K step
    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      const int ns = samples_in(group_of(it));
      const int db = (int)(it & 1);
      if (it >= kF3Dt) f3_wait(&dt_free[db], (uint32_t)(((it >> 1) - 1) & 1));  // the accumulator of group it - 2 has been read out
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 3);
        F3_TRACE(2, 0, g);
        f3_wait(&h_full[b], (uint32_t)((g >> 2) & 1));
        tc_fence_after();
        F3_TRACE(2, 1, g);
        for (int s = 0; s < ns; s++) {
          f3_wait(&sc_full[ss], sp);
          tc_fence_after();
          if (s == 0) { F3_TRACE(2, 2, g); }
          if (elect_one()) {
            const uint32_t s_addr = sring + ss * L.sc_unit;
            const uint64_t s_hi = sd0 + (uint64_t)((s_addr >> 4) & 0x3FFFu), s_lo = s_hi + (uint64_t)(mat_bytes >> 4);  // (14-bit address field: the shared window of cluster rank 1 has higher bits set)
            const uint32_t d = tmem + TM_DT + (uint32_t)(db * N) + (uint32_t)(s * VP);
            const uint32_t p_hi = tmem + (uint32_t)b * (uint32_t)N + (uint32_t)(s * (VP >> 1)), p_lo = p_hi + (uint32_t)(N >> 1);
#pragma unroll
            for (int k = 0; k < 5; k++)
              if (k < ksteps) mma_ts(d, p_hi + k * 8, s_hi + 16 * k, idescM, (h | k) != 0);
            if (kSplit == 3) {
#pragma unroll
              for (int k = 0; k < 5; k++)
                if (k < ksteps) mma_ts(d, p_lo + k * 8, s_hi + 16 * k, idescM, 1);
#pragma unroll
              for (int k = 0; k < 5; k++)
                if (k < ksteps) mma_ts(d, p_hi + k * 8, s_lo + 16 * k, idescM, 1);
            }
            mma_commit(&sc_empty[ss]);
            if (s == ns - 1) {
              mma_commit(&pt_free[b]);
              if (h == H - 1) mma_commit(&dt_full[db]);
            }
          }
          __syncwarp();
          if (++ss == L.sc_stages) ss = 0, sp ^= 1;
        }
        F3_TRACE(2, 3, g);
      }
    }
  } else if (warp == 3) {
    // (idle)
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ x-tile builders + LayerNorm (256 threads)
    const int lt = tid - 384;         // 0..255
    const int lw = lt >> 5;           // 0..7
    const int c = lt & 15;            // 16-byte chunk of a row's bf16 image = 8 features
    const int row0 = lt >> 4;         // 0..15: this thread's rows are row0 + 16 k, k = 0..4 (N <= 80 rows)
    const uint32_t blk = (uint32_t)N * 128u;
    const uint32_t cc = (uint32_t)(c & 7);

    auto build_tiles = [&](int64_t it) {  // group it -> bf16 hi/lo K-major SW128 rows of the group's tokens in tile set it & 1
      const int64_t grp = group_of(it);
      const int rows = samples_in(grp) * V;
      uint8_t* tile = smem + L.xb((int)(it & 1)) + (uint32_t)(c >> 3) * blk;
      const float* xg = a.x[net] + (grp * G * V + row0) * 128 + c * 8;
      if (lw == 0) { F3_TRACE(3, 0, it); }
      float4 v[5][2];
#pragma unroll
      for (int k = 0; k < 5; k++)  // (requested before the tile set is free: registers, not shared memory, hold them)
        if (row0 + 16 * k < rows) {
          const float4* src = reinterpret_cast<const float4*>(xg + (size_t)k * 16 * 128);
          v[k][0] = __ldg(src), v[k][1] = __ldg(src + 1);
        }
      if (it >= kF3Xb) f3_wait(&xb_free[it & 1], (uint32_t)(((it >> 1) - 1) & 1));  // the projections of group it - 2 have retired
      if (lw == 0) { F3_TRACE(3, 1, it); }
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const int gr = row0 + 16 * k;
        if (gr < rows) {
          const int s = gr / V, at = gr - s * V;
          const uint32_t r = (uint32_t)(s * VP + at);
          uint32_t hi[4], lo[4];
          split2(v[k][0].x, v[k][0].y, hi[0], lo[0]);
          split2(v[k][0].z, v[k][0].w, hi[1], lo[1]);
          split2(v[k][1].x, v[k][1].y, hi[2], lo[2]);
          split2(v[k][1].z, v[k][1].w, hi[3], lo[3]);
          const uint32_t off = r * 128u + ((cc ^ (r & 7u)) << 4);
          *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (kSplit == 3) *reinterpret_cast<uint4*>(tile + 2 * blk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&xb_full[it & 1]);
      if (lw == 0) { F3_TRACE(3, 2, it); }
    };
    // LayerNorm: half a warp per token row (8 features per lane: 4 shuffle stages); pre-LayerNorm rows come back through L2
    auto layer_norm = [&](int64_t it) {
      const int64_t grp = group_of(it);
      const int rows = samples_in(grp) * V;
      f3_wait(&rows_out[it & 1], (uint32_t)((it >> 1) & 1));
      const int hwr = lane >> 4, c16 = lane & 15;
      const float4* gp = reinterpret_cast<const float4*>(a.gamma[net]) + 2 * c16;
      const float4* bp = reinterpret_cast<const float4*>(a.beta[net]) + 2 * c16;
      const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
      const float4* xg = reinterpret_cast<const float4*>(a.x[net] + grp * G * V * 128) + 2 * c16;
      float4* og = reinterpret_cast<float4*>(a.out[net] + grp * G * V * 128) + 2 * c16;
      constexpr int kRows = 3;  // row pairs of a warp in flight
      for (int rb = 2 * lw; rb < rows; rb += 16 * kRows) {  // warp lw: rows 2 lw + {0, 1} + 16 j (warp-uniform trip count: shuffles inside)
        const int r0 = rb + hwr;
        float4 xv[kRows][2], sv[kRows][2];
#pragma unroll
        for (int j = 0; j < kRows; j++) {
          const int r = r0 + 16 * j;
          if (r < rows) {
            xv[j][0] = __ldg(xg + (size_t)r * 32), xv[j][1] = __ldg(xg + (size_t)r * 32 + 1);
            sv[j][0] = __ldcg(og + (size_t)r * 32), sv[j][1] = __ldcg(og + (size_t)r * 32 + 1);  // (L2, not the read-only path)
          }
        }
#pragma unroll
        for (int j = 0; j < kRows; j++) {
          const int r = r0 + 16 * j;
          const bool ok = r < rows;
          float y[8];
          y[0] = xv[j][0].x + sv[j][0].x, y[1] = xv[j][0].y + sv[j][0].y, y[2] = xv[j][0].z + sv[j][0].z, y[3] = xv[j][0].w + sv[j][0].w;
          y[4] = xv[j][1].x + sv[j][1].x, y[5] = xv[j][1].y + sv[j][1].y, y[6] = xv[j][1].z + sv[j][1].z, y[7] = xv[j][1].w + sv[j][1].w;
          if (!ok) {
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = 0.f;
          }
          float sum = ((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]));
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++) sq = fmaf(y[i], y[i], sq);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {  // (stays inside the half warp)
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          const float mean = sum * (1.f / 128.f);
          const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
          const float rstd = 1.0f / sqrtf(var + a.eps);
          if (ok) {
            float4 o0, o1;
            o0.x = (y[0] - mean) * rstd * g0.x + b0.x, o0.y = (y[1] - mean) * rstd * g0.y + b0.y;
            o0.z = (y[2] - mean) * rstd * g0.z + b0.z, o0.w = (y[3] - mean) * rstd * g0.w + b0.w;
            o1.x = (y[4] - mean) * rstd * g1.x + b1.x, o1.y = (y[5] - mean) * rstd * g1.y + b1.y;
            o1.z = (y[6] - mean) * rstd * g1.z + b1.z, o1.w = (y[7] - mean) * rstd * g1.w + b1.w;
            og[(size_t)r * 32] = o0, og[(size_t)r * 32 + 1] = o1;
          }
        }
      }
      mbar_arrive(&ln_done[it & 1]);
      if (lw == 0) { F3_TRACE(3, 3, it); }
    };

    if (my_groups > 0) build_tiles(0);
    if (my_groups > 1) build_tiles(1);
    for (int64_t it = 0; it < my_groups; it++) {  // (the tile set of group it is free before its rows are out: build first)
      if (it + 2 < my_groups) build_tiles(it + 2);
      layer_norm(it);
      if (lw == 0) { F3_TRACE(3, 4, it); }
    }
  } else {
    // ------------------------------------------------------------------ epilogue groups (column halves of the N tokens), warps 4..11
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int e = (warp - 4) >> 2;          // column half
    const int f = q * 32 + lane;            // feature = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int half = N >> 1;                // fp32 columns per group (multiple of 8, <= 40)
    const int nchunk = half >> 3;           // chunks of 8 columns (<= 5)
    uint32_t r[5][8];
    const int col0 = e * half;
    const int s00 = col0 / VP, at00 = col0 - s00 * VP;  // start of this thread's column range as (sample, atom)

    auto drain = [&](int64_t pit) {  // accumulator of group pit -> registers (DT handed back at once) -> pre-LayerNorm rows in `out`
      const int db = (int)(pit & 1);
      f3_wait(&dt_full[db], (uint32_t)((pit >> 1) & 1));
      tc_fence_after();
      const uint32_t dbase = tmem + lane_base + TM_DT + (uint32_t)(db * N) + (uint32_t)col0;
#pragma unroll
      for (int i = 0; i < 5; i++)
        if (i < nchunk) f3_ld8(dbase + (uint32_t)(8 * i), r[i]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&dt_free[db]);
      const int64_t grp = group_of(pit);
      const int ns = samples_in(grp);
      float* og = a.out[net] + grp * G * V * 128 + f;
      int s = s00, at = at00;  // 4-byte stores, 128 contiguous bytes per warp; no division in the loop
#pragma unroll
      for (int i = 0; i < 5; i++)
        if (i < nchunk) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (s < ns && at < V) og[(size_t)(s * V + at) * 128] = __uint_as_float(r[i][j]);
            if (++at == VP) at = 0, s++;
          }
        }
      __threadfence();  // the LayerNorm warps read these rows through L2
      if (pit >= 2) f3_wait(&ln_done[db], (uint32_t)(((pit >> 1) - 1) & 1));  // rows_out[db] may be one phase ahead of its reader at most
      mbar_arrive(&rows_out[db]);
      if (q == 0 && e == 0) { F3_TRACE(1, 2, pit); }
    };

    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 3);
        f3_wait(&pt_full[b], (uint32_t)((g >> 2) & 1));
        tc_fence_after();
        if (q == 0 && e == 0) { F3_TRACE(1, 0, g); }
        const uint32_t base = tmem + lane_base + (uint32_t)b * (uint32_t)N;
#pragma unroll
        for (int i = 0; i < 5; i++)
          if (i < nchunk) f3_ld8(base + (uint32_t)(col0 + 8 * i), r[i]);
        tmem_ld_wait();
        f3_epi_bar();  // both halves have read their fp32 columns: the in-place writes below may cross into the other half
#pragma unroll
        for (int i = 0; i < 5; i++)
          if (i < nchunk) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; j++) split2(__uint_as_float(r[i][2 * j]), __uint_as_float(r[i][2 * j + 1]), hi[j], lo[j]);
            const uint32_t col = (uint32_t)((col0 + 8 * i) >> 1);  // packed column of tokens (8i, 8i+1)
            f3_st4(base + col, hi[0], hi[1], hi[2], hi[3]);
            if (kSplit == 3) f3_st4(base + (uint32_t)half + col, lo[0], lo[1], lo[2], lo[3]);
          }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&h_full[b]);
        if (q == 0 && e == 0) { F3_TRACE(1, 1, g); }
        if (h == 0 && it > 0) drain(it - 1);  // (two accumulators: the mixing of this group does not wait for it)
      }
    }
    if (my_groups > 0) drain(my_groups - 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into this CTA's ring / arrive on its barriers until here
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef F3_TRACE
}

// -------------------------------------------------------------------------------------------- host side
long long* tc_get_fm_trace();

// Group size and ring depths for an atom count; false if a group of whole samples does not fit 80 tokens.
static bool f3_plan(int VP, int64_t n, int* G, int* wcs, int* scs, int* smem_bytes) {
  if (VP > kF3MaxN || VP < 16) return false;
  int g = kF3MaxN / VP;
  if ((int64_t)g > n) g = (int)(n < 1 ? 1 : n);
  for (int w = 3; w >= 2; w--)  // (a head's W_c is 2 units of 32 KB: the ring holds 1.5 heads)
    for (int s = 4; s >= (g > 2 ? g : 2); s--) {
      const int total = (int)F3Smem(VP, g, w, s).total();
      if (total <= 232448) {
        *G = g, *wcs = w, *scs = s, *smem_bytes = total;
        return true;
      }
    }
  return false;
}

bool tc_attn_fm3_supported(int V, int64_t n) {
  static int use = -1;
  if (use < 0) {
    const char* e = getenv("TW_ATTN_FM");  // bring-up switch: 2 = the two-buffer kernel of attn_fm.cu for every atom count
    use = e ? atoi(e) : 3;
  }
  int G, w, s, b;
  return use >= 3 && n >= 1 && f3_plan((V + 15) / 16 * 16, n, &G, &w, &s, &b);
}

int tc_attn_fm3(const tw_flow_config* c, const float* const x[2], float* const out[2], const uint8_t* const wc[2],
                const float* const gamma[2], const float* const beta[2], const uint8_t* scores_img, int64_t n, int64_t n_cond, int V,
                int nets, cudaStream_t st) {
  const int VP = (V + 15) / 16 * 16;
  F3Args a{};
  int smem_bytes = 0;
  TW_CHECK_ARG(f3_plan(VP, n, &a.G, &a.wc_stages, &a.sc_stages, &smem_bytes), "feature-major attention (deep pipeline): atom count out of range");
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_attn_fm3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    TW_CUDA(cudaFuncSetAttribute(k_attn_fm3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done.mark();
  }
  for (int s = 0; s < 2; s++) a.x[s] = x[s], a.out[s] = out[s], a.wc[s] = wc[s], a.gamma[s] = gamma[s], a.beta[s] = beta[s];
  a.scores_img = scores_img;
  a.n = n, a.n_cond = n_cond, a.V = V, a.VP = VP, a.H = c->num_heads, a.eps = c->layer_norm_eps;
  a.trace = tc_get_fm_trace();
  {
    const char* e = getenv("TW_FM3_DBG");
    a.dbg = e ? atoi(e) : 0;
  }
  const int per_net = nets == 1 ? 148 : 74;
  const int64_t groups = (n + a.G - 1) / a.G;
  dim3 grid((unsigned)(groups < per_net ? (groups + 1) / 2 * 2 : per_net), nets);  // CTA pairs (clusters of two along x)
  if (c->precision == TW_PRECISION_BF16X3)
    k_attn_fm3<3><<<grid, kF3Threads, smem_bytes, st>>>(a);
  else
    k_attn_fm3<1><<<grid, kF3Threads, smem_bytes, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // namespace tw

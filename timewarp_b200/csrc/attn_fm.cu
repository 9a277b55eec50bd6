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
//   drain: DT -> pre-LayerNorm rows in `out` (the transposition is free: lane f stores feature f, 128 contiguous bytes per warp)
//   LayerNorm: one warp per token row re-reads the row (L2) and the residual x, warp-shuffle statistics, coalesced float4 stores
//
// The previous fused kernel (k_attn_fused, flow_tc.cu) mixed first with the TOKENS of one sample on the lanes: 65 of 128 MMA
// rows real at 65 atoms, 39 M128xN128 MMAs per (sample, head).  Here a (sample, head) costs 24 MMAs of N = G*VP / G (projection)
// + 3 VP/16 MMAs of N = VP (mixing): 39 MMAs of N = 80 at 65 atoms -- 0.625 of the tensor work, and atom counts up to 128.
//
// Warp roles (480 threads): 0 producer (one thread feeds the W_c, score-image and staging rings with bulk async copies),
// 1 MMA issuer for the projections, 14 MMA issuer for the mixing, 2-5 x-tile builders + LayerNorm (both read raw fp32 rows
// from the staging ring), 6-13 two epilogue groups (column halves of the N tokens).
// TMEM: PT0 | PT1 | DT, N = G VP <= 160 columns each.  PT is double buffered (the conversion of PT(g) overlaps P(g+1) and
// M(g-1)); the x tiles and DT are single buffers whose hand-overs are covered by the other issuer's queued work.
#include <stdlib.h>

#include "flow_tc.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

constexpr int kFmThreads = 480;
constexpr int kFmWcStage = 16384;  // the hi or the lo image of one [128 out x 64 in] K block of W_c,h
constexpr int kFmXsStages = 4, kFmXsRows = 16, kFmXsChunk = kFmXsRows * 512;  // staging ring of raw x rows (bulk-copied)
constexpr int kFmMaxN = 160;       // tokens of a group on the MMA N axis: TMEM holds PT0 | PT1 | DT, N columns each (3 N <= 512)

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
};

struct FmSmem {
  uint32_t N, xb_bytes, sc_unit, wc_stages, sc_stages;
  __host__ __device__ FmSmem(int VP, int G, int wcs, int scs) {
    N = (uint32_t)(G * VP);
    xb_bytes = N * 512u;                       // hi kb0 | hi kb1 | lo kb0 | lo kb1, each [N x 128 B]
    sc_unit = (uint32_t)(2 * VP * VP * 2);     // hi | lo image of one (sample, head)
    wc_stages = (uint32_t)wcs, sc_stages = (uint32_t)scs;
  }
  __host__ __device__ uint32_t xb() const { return 0; }
  __host__ __device__ uint32_t wc() const { return xb_bytes; }
  __host__ __device__ uint32_t sc() const { return wc() + wc_stages * kFmWcStage; }
  __host__ __device__ uint32_t xs() const { return sc() + sc_stages * sc_unit; }  // raw fp32 x rows: kFmXsStages chunks of 16 rows
  __host__ __device__ uint32_t bars() const { return xs() + kFmXsStages * kFmXsChunk; }
  __host__ __device__ uint32_t total() const { return bars() + 512; }
};

// mbarrier wait with a watchdog: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void fm_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // (no printf: a call in the wait loop makes the compiler save live registers around it)
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
  const FmSmem L(VP, G, a.wc_stages, a.sc_stages);
  const int N = (int)L.N;
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  constexpr int kParts = kSplit == 3 ? 2 : 1;
  const int64_t n_groups_total = (a.n + G - 1) / G;
  const int64_t my_groups = ((int64_t)blockIdx.x < n_groups_total) ? (n_groups_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto group_of = [&](int64_t it) -> int64_t { return blockIdx.x + it * gridDim.x; };
  auto samples_in = [&](int64_t grp) -> int { int64_t r = a.n - grp * G; return (int)(r < G ? r : G); };
  const uint32_t TM_PT = 0, TM_DT = 2u * (uint32_t)N;  // PT0 | PT1 | DT

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars());
  uint64_t* wc_full = bars;                 // [8]
  uint64_t* wc_empty = wc_full + 8;         // [8]
  uint64_t* sc_full = wc_empty + 8;         // [4]
  uint64_t* sc_empty = sc_full + 4;         // [4]
  uint64_t* xb_full = sc_empty + 4;         // 128 arrivals: x tiles of a group written
  uint64_t* xb_free = xb_full + 1;          // commit: the last projection MMA of a group retired (the tiles may be rebuilt)
  uint64_t* pt_full = xb_free + 1;          // [2] commit: P(g) retired
  uint64_t* h_full = pt_full + 2;           // [2] 256 arrivals: PT(g) converted in place
  uint64_t* dt_full = h_full + 2;           // commit: last M of the group retired
  uint64_t* dt_free = dt_full + 1;          // 256 arrivals: DT read out (pre-LayerNorm rows on their way to global memory)
  uint64_t* rows_out = dt_free + 1;         // 256 arrivals: the group's pre-LayerNorm rows are in global memory
  uint64_t* pt_free = rows_out + 1;         // [2] commit: M(g) retired, PT[g & 1] may be overwritten by P(g + 2)
  uint64_t* xs_full = pt_free + 2;          // [4] raw x rows of a chunk landed (tx bytes)
  uint64_t* xs_empty = xs_full + 4;         // [4] 128 arrivals: chunk converted
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xs_empty + 4);

  if (tid == 0) {
    for (int i = 0; i < 8; i++) mbar_init(&wc_full[i], 1), mbar_init(&wc_empty[i], 1);
    for (int i = 0; i < 4; i++) mbar_init(&sc_full[i], 1), mbar_init(&sc_empty[i], 1), mbar_init(&xs_full[i], 1), mbar_init(&xs_empty[i], 128);
    for (int i = 0; i < 2; i++) mbar_init(&pt_full[i], 1), mbar_init(&h_full[i], 256), mbar_init(&pt_free[i], 1);
    mbar_init(xb_full, 128), mbar_init(xb_free, 1), mbar_init(dt_full, 1), mbar_init(dt_free, 256), mbar_init(rows_out, 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (uint32_t i = tid * 16; i < L.xb_bytes; i += kFmThreads * 16) *reinterpret_cast<uint4*>(smem + L.xb() + i) = make_uint4(0, 0, 0, 0);  // padding rows stay zero
  fence_proxy_async_smem();
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
    // ------------------------------------------------------------------ producer: ONE thread feeds the three rings
    // (W_c units, score images, staging chunks) with bulk async copies, polling their "empty" barriers without blocking:
    // separate spinning lanes of one warp starve each other, and a blocked W_c producer must not hold back the other rings.
    if (lane == 0) {
      // W_c: per head K block 0 hi, lo, K block 1 hi, lo
      uint32_t ws = 0, wp = 0;
      int64_t w_left = my_groups * H * 2 * kParts;
      int w_h = 0, w_u = 0;
      // score images: one unit per (group, head, sample)
      uint32_t ss = 0, sp = 0;
      int64_t s_it = 0;
      int s_h = 0, s_s = 0;
      bool s_done = my_groups == 0;
      // staging ring: per group the raw x rows (build), then the LayerNorm inputs of the previous group (x rows + pre-LN rows)
      uint32_t xs = 0, xp = 0, ph_rows = 0;
      int64_t x_it = 0;           // group whose chunks are being pushed
      int x_phase = 0;            // 0 build chunks of x_it, 1 wait for rows_out of x_it - 1, 2 LayerNorm chunks of x_it - 1
      int x_r0 = 0, x_half = 0;
      bool x_done = my_groups == 0;
      long long t_idle = 0;
      // the score images (157 MB per launch at 1024 x 65 atoms) are read once per launch: evict them first, so that the layer
      // input / output rows (read three times) and the weights stay L2-resident
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      while (w_left > 0 || !s_done || !x_done) {
        bool progress = false;
        if (w_left > 0 && mbar_test_wait(&wc_empty[ws], wp ^ 1)) {
          const int kb = w_u / kParts, part = w_u % kParts;
          mbar_arrive_expect_tx(&wc_full[ws], (uint32_t)kFmWcStage);
          bulk_g2s_hint(smem + L.wc() + ws * kFmWcStage, a.wc[net] + (size_t)(w_h * 2 + kb) * 32768 + part * 16384, kFmWcStage, &wc_full[ws], pol_keep);
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
          if (++w_u == 2 * kParts) {
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
          bulk_g2s_hint(smem + L.sc() + ss * L.sc_unit, src, kParts * mat_bytes, &sc_full[ss], pol_stream);
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
        if (!x_done) {
          if (x_phase == 1) {
            if (mbar_test_wait(rows_out, ph_rows)) {
              ph_rows ^= 1;
              asm volatile("fence.proxy.async;" ::: "memory");  // rows written with generic stores, read by the bulk copies
              x_phase = 2, x_r0 = 0, x_half = 0;
              progress = true;
            }
          } else if (mbar_test_wait(&xs_empty[xs], xp ^ 1)) {
            const int64_t grp = group_of(x_phase == 0 ? x_it : x_it - 1);
            const int rows = samples_in(grp) * V;
            const uint32_t bytes = (uint32_t)(rows - x_r0 < kFmXsRows ? rows - x_r0 : kFmXsRows) * 512u;
            const float* base = (x_phase == 2 && x_half == 1) ? a.out[net] : a.x[net];
            mbar_arrive_expect_tx(&xs_full[xs], bytes);
            bulk_g2s_hint(smem + L.xs() + xs * kFmXsChunk, base + (grp * G * V + x_r0) * 128, bytes, &xs_full[xs], pol_keep);
            if (++xs == kFmXsStages) xs = 0, xp ^= 1;
            if (x_phase == 0) {
              x_r0 += kFmXsRows;
              if (x_r0 >= rows) {  // build chunks of x_it done: LayerNorm inputs of the previous group next (if any)
                x_r0 = 0;
                if (x_it >= 1) x_phase = 1;
                else if (++x_it == my_groups) x_phase = 1;  // (single group: its LayerNorm inputs)
              }
            } else {
              if (++x_half == 2) {
                x_half = 0, x_r0 += kFmXsRows;
                if (x_r0 >= rows) {  // LayerNorm inputs of group x_it - 1 pushed
                  x_r0 = 0;
                  if (x_it == my_groups) x_done = true;          // that was the last group's LayerNorm
                  else if (++x_it == my_groups) x_phase = 1;      // no more builds: the last group's LayerNorm inputs
                  else x_phase = 0;
                }
              }
            }
            progress = true;
          }
        }
        if (progress) {
          t_idle = 0;
        } else {
          if (t_idle == 0) t_idle = clock64();
          else if (clock64() - t_idle > 4000000000LL) {
            printf("k_attn_fm: producer timeout (block %d,%d)\n", blockIdx.x, blockIdx.y);
            __trap();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer 1: projections  PT[g & 1] = W_c,h X^T
    // (Two issuing warps: the bookkeeping around an MMA block -- mbarrier polls, descriptor arithmetic, commits -- costs about as
    // many cycles as the block's tensor time at N = 80..160, so ONE issuer leaves the tensor pipe half idle.  The pipe takes
    // MMAs from both warps; every dependency between P and M goes through an mbarrier.)
    uint32_t ws = 0, wp = 0, ph_xb = 0, ph_ptfree = 0;
    const uint32_t idescP = make_idesc_bf16(128, (uint32_t)N, 0, 0);
    const uint32_t blk = (uint32_t)N * 128u;  // one [N x 64] K block of the x tiles
    const uint32_t xt = smem_u32(smem + L.xb());
    const uint32_t wring = smem_u32(smem + L.wc());
    // the x tiles never move: their descriptors are loop invariants
    const uint64_t xd_hi0 = desc_kmajor_sw128(xt), xd_hi1 = desc_kmajor_sw128(xt + blk);
    const uint64_t xd_lo0 = desc_kmajor_sw128(xt + 2 * blk), xd_lo1 = desc_kmajor_sw128(xt + 3 * blk);
    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      fm_wait(xb_full, ph_xb);
      ph_xb ^= 1;
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 1);
        if (g >= 2) {
          fm_wait(&pt_free[b], (ph_ptfree >> b) & 1u);
          ph_ptfree ^= 1u << b;
        }
        FM_TRACE(0, 0, g);
        const uint32_t d = tmem + TM_PT + (uint32_t)b * (uint32_t)N;
#pragma unroll
        for (int kb = 0; kb < 2; kb++) {
          const uint64_t xh = kb ? xd_hi1 : xd_hi0, xl = kb ? xd_lo1 : xd_lo0;
          fm_wait(&wc_full[ws], wp);  // hi image of this K block
          tc_fence_after();
          if (elect_one()) {
            const uint64_t wd = desc_kmajor_sw128(wring + ws * kFmWcStage);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(d, wd + 2 * k, xh + 2 * k, idescP, (kb | k) != 0);
            if (kSplit == 3) {
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, wd + 2 * k, xl + 2 * k, idescP, 1);
            }
            mma_commit(&wc_empty[ws]);
            if (kSplit != 3 && kb == 1) {
              mma_commit(&pt_full[b]);
              if (h == H - 1) mma_commit(xb_free);
            }
          }
          __syncwarp();
          if (++ws == L.wc_stages) ws = 0, wp ^= 1;
          if (kSplit == 3) {
            fm_wait(&wc_full[ws], wp);  // lo image
            tc_fence_after();
            if (elect_one()) {
              const uint64_t wd = desc_kmajor_sw128(wring + ws * kFmWcStage);
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, wd + 2 * k, xh + 2 * k, idescP, 1);
              mma_commit(&wc_empty[ws]);
              if (kb == 1) {
                mma_commit(&pt_full[b]);
                if (h == H - 1) mma_commit(xb_free);
              }
            }
            __syncwarp();
            if (++ws == L.wc_stages) ws = 0, wp ^= 1;
          }
        }
        FM_TRACE(0, 1, g);
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------------------------ MMA issuer 2: mixing  DT[:, sample s] += PT(g)[:, sample s] A_h(s)^T
    uint32_t ss = 0, sp = 0, ph_h = 0, ph_dtfree = 0;
    const uint32_t idescM = make_idesc_bf16(128, (uint32_t)VP, 0, 0);
    const uint32_t sc_sbo = (uint32_t)(VP >> 3) * 128;
    const uint32_t sring = smem_u32(smem + L.sc());
    const uint64_t sd0 = make_smem_desc(0, 128, sc_sbo, LAYOUT_NONE);  // + (address >> 4) of the image, + 16 per K step
    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      const int ns = samples_in(group_of(it));
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 1);
        if (h == 0 && it >= 1) {  // the (single) accumulator was last used by the previous group: read out?
          fm_wait(dt_free, ph_dtfree);
          ph_dtfree ^= 1;
        }
        FM_TRACE(2, 0, g);
        fm_wait(&h_full[b], (ph_h >> b) & 1u);
        ph_h ^= 1u << b;
        tc_fence_after();
        FM_TRACE(2, 1, g);
        for (int s = 0; s < ns; s++) {
          fm_wait(&sc_full[ss], sp);
          tc_fence_after();
          if (s == 0) { FM_TRACE(2, 2, g); }
          if (elect_one()) {
            const uint32_t s_addr = sring + ss * L.sc_unit;
            const uint64_t s_hi = sd0 + (uint64_t)((s_addr >> 4) & 0x3FFFu), s_lo = s_hi + (uint64_t)(mat_bytes >> 4);  // (14-bit address field: the shared window of cluster rank 1 has higher bits set)
            const uint32_t d = tmem + TM_DT + (uint32_t)(s * VP);
            const uint32_t p_hi = tmem + TM_PT + (uint32_t)b * (uint32_t)N + (uint32_t)(s * (VP >> 1)), p_lo = p_hi + (uint32_t)(N >> 1);
#pragma unroll
            for (int k = 0; k < 8; k++)
              if (k < ksteps) mma_ts(d, p_hi + k * 8, s_hi + 16 * k, idescM, (h | k) != 0);
            if (kSplit == 3) {
#pragma unroll
              for (int k = 0; k < 8; k++)
                if (k < ksteps) mma_ts(d, p_lo + k * 8, s_hi + 16 * k, idescM, 1);
#pragma unroll
              for (int k = 0; k < 8; k++)
                if (k < ksteps) mma_ts(d, p_hi + k * 8, s_lo + 16 * k, idescM, 1);
            }
            mma_commit(&sc_empty[ss]);
            if (s == ns - 1) {
              mma_commit(&pt_free[b]);
              if (h == H - 1) mma_commit(dt_full);
            }
          }
          __syncwarp();
          if (++ss == L.sc_stages) ss = 0, sp ^= 1;
        }
        FM_TRACE(2, 3, g);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ x-tile builders + LayerNorm (128 threads)
    const int lt = tid - 64;          // 0..127
    const int lw = lt >> 5;           // 0..3
    const int c = lane & 15;          // 16-byte chunk of a row's bf16 image = 8 features
    uint32_t ph_free = 0;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma[net]) + lane), bt = __ldg(reinterpret_cast<const float4*>(a.beta[net]) + lane);

    uint32_t xs = 0, xp = 0;
    auto build_tiles = [&](int64_t it) {  // group it -> bf16 hi/lo K-major SW128 rows of the group's tokens
      // The raw rows arrive through the staging ring (bulk copies issued ahead of time, no registers held across the latency):
      // per chunk of 16 rows, 128 threads x 2 tasks of one 16-byte image chunk (8 features of one row) each.
      const int rows = samples_in(group_of(it)) * V;
      uint8_t* tile = smem + L.xb();
      const uint32_t blk = (uint32_t)N * 128u;
      const int kb = c >> 3, cc = c & 7;
      if (lw == 0) { FM_TRACE(3, 0, it); }
      if (it >= 1) {
        fm_wait(xb_free, ph_free);
        ph_free ^= 1;
      }
      if (lw == 0) { FM_TRACE(3, 1, it); }
      for (int r0 = 0; r0 < rows; r0 += kFmXsRows) {
        fm_wait(&xs_full[xs], xp);
        const uint8_t* stg = smem + L.xs() + xs * kFmXsChunk;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int rl = (lt >> 4) + 8 * j;  // row inside the chunk
          const int gr = r0 + rl;
          if (gr < rows) {
            const float4 v0 = *reinterpret_cast<const float4*>(stg + rl * 512 + c * 32);
            const float4 v1 = *reinterpret_cast<const float4*>(stg + rl * 512 + c * 32 + 16);
            const int s = gr / V, at = gr - s * V;
            const uint32_t r = (uint32_t)(s * VP + at);
            uint32_t hi[4], lo[4];
            split2(v0.x, v0.y, hi[0], lo[0]);
            split2(v0.z, v0.w, hi[1], lo[1]);
            split2(v1.x, v1.y, hi[2], lo[2]);
            split2(v1.z, v1.w, hi[3], lo[3]);
            const uint32_t off = kb * blk + r * 128u + (((uint32_t)cc ^ (r & 7u)) << 4);
            *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (kSplit == 3) *reinterpret_cast<uint4*>(tile + 2 * blk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        mbar_arrive(&xs_empty[xs]);
        if (++xs == kFmXsStages) xs = 0, xp ^= 1;
      }
      fence_proxy_async_smem();
      mbar_arrive(xb_full);
      if (lw == 0) { FM_TRACE(3, 2, it); }
    };
    auto layer_norm = [&](int64_t it) {  // pre-LayerNorm rows of group it + residual x rows (staging ring) -> LayerNorm -> out
      const int64_t grp = group_of(it);
      const int rows = samples_in(grp) * V;
      float* og = a.out[net] + grp * G * V * 128;
      for (int r0 = 0; r0 < rows; r0 += kFmXsRows) {
        const uint32_t sa = xs;
        fm_wait(&xs_full[xs], xp);
        if (++xs == kFmXsStages) xs = 0, xp ^= 1;
        const uint32_t sb = xs;
        fm_wait(&xs_full[xs], xp);
        if (++xs == kFmXsStages) xs = 0, xp ^= 1;
        const uint8_t* xa = smem + L.xs() + sa * kFmXsChunk;
        const uint8_t* pa = smem + L.xs() + sb * kFmXsChunk;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int rl = lw * 4 + j;
          if (r0 + rl < rows) {  // (warp-uniform)
            const float4 xv = *reinterpret_cast<const float4*>(xa + rl * 512 + lane * 16);
            const float4 sv = *reinterpret_cast<const float4*>(pa + rl * 512 + lane * 16);
            const float y0 = xv.x + sv.x, y1 = xv.y + sv.y, y2 = xv.z + sv.z, y3 = xv.w + sv.w;
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
            *(reinterpret_cast<float4*>(og + (size_t)(r0 + rl) * 128) + lane) = o4;
          }
        }
        mbar_arrive(&xs_empty[sa]);
        mbar_arrive(&xs_empty[sb]);
      }
      if (lw == 0) { FM_TRACE(3, 3, it); }
    };

    for (int64_t it = 0; it < my_groups; it++) {
      build_tiles(it);
      if (it >= 1) layer_norm(it - 1);
      if (lw == 0) { FM_TRACE(3, 4, it); }
    }
    if (my_groups > 0) layer_norm(my_groups - 1);
  } else {
    // ------------------------------------------------------------------ epilogue groups (column halves of the N tokens)
    const int q = warp & 3;                 // TMEM lane quarter of this warp (warps 6..13: 6&3 = 2, ...)
    const int e = (warp - 6) >> 2;          // column half
    const int f = q * 32 + lane;            // feature = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int half = N >> 1;                // fp32 columns per group (multiple of 8)
    const int nchunk = half >> 3;           // chunks of 8 columns (<= 10)
    uint32_t ph_pt = 0, ph_dt = 0;

    uint32_t r[10][8];
    // start of this thread's column range as (sample, atom): one division per launch
    const int col0 = e * half;
    const int s00 = col0 / VP, at00 = col0 - s00 * VP;

    auto drain = [&](int64_t pit) {  // accumulator of group pit -> registers (DT handed back at once) -> pre-LayerNorm rows in `out`
      fm_wait(dt_full, ph_dt);
      ph_dt ^= 1;
      tc_fence_after();
      const uint32_t dbase = tmem + lane_base + TM_DT + (uint32_t)col0;
#pragma unroll
      for (int i = 0; i < 10; i++)
        if (i < nchunk) tmem_ld8(dbase + (uint32_t)(8 * i), r[i]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dt_free);
      const int64_t grp = group_of(pit);
      const int ns = samples_in(grp);
      float* og = a.out[net] + grp * G * V * 128 + f;
      int s = s00, at = at00;  // 4-byte stores, 128 contiguous bytes per warp; no division in the loop
#pragma unroll
      for (int i = 0; i < 10; i++)
        if (i < nchunk) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (s < ns && at < V) og[(size_t)(s * V + at) * 128] = __uint_as_float(r[i][j]);
            if (++at == VP) at = 0, s++;
          }
        }
      __threadfence();  // the bulk copies of the LayerNorm stage read these rows through L2
      mbar_arrive(rows_out);
      if (q == 2 && e == 0) { FM_TRACE(1, 2, pit); }
    };

    int64_t g = 0;
    for (int64_t it = 0; it < my_groups; it++) {
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 1);
        fm_wait(&pt_full[b], (ph_pt >> b) & 1u);
        ph_pt ^= 1u << b;
        tc_fence_after();
        if (q == 2 && e == 0) { FM_TRACE(1, 0, g); }
        const uint32_t base = tmem + lane_base + TM_PT + (uint32_t)b * (uint32_t)N;
#pragma unroll
        for (int i = 0; i < 10; i++)
          if (i < nchunk) tmem_ld8(base + (uint32_t)(col0 + 8 * i), r[i]);
        tmem_ld_wait();
        fm_epi_bar();  // both halves have read their fp32 columns: the in-place writes below may cross into the other half
#pragma unroll
        for (int i = 0; i < 10; i++)
          if (i < nchunk) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; j++) split2(__uint_as_float(r[i][2 * j]), __uint_as_float(r[i][2 * j + 1]), hi[j], lo[j]);
            const uint32_t col = (uint32_t)((col0 + 8 * i) >> 1);  // packed column of tokens (8i, 8i+1)
            tmem_st4(base + col, hi[0], hi[1], hi[2], hi[3]);
            if (kSplit == 3) tmem_st4(base + (uint32_t)half + col, lo[0], lo[1], lo[2], lo[3]);
          }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&h_full[b]);
        if (q == 2 && e == 0) { FM_TRACE(1, 1, g); }
        // The previous group's accumulator: its last mixing head was issued before this group's first projection retired, so
        // the first conversion of the new group goes first (the mixing issuer needs it next) and the drain follows.
        if (h == 0 && it > 0) drain(it - 1);
      }
    }
    if (my_groups > 0) drain(my_groups - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef FM_TRACE
}

// -------------------------------------------------------------------------------------------- host side
static long long* g_fm_trace = nullptr;
void tc_set_fm_trace(long long* buf) { g_fm_trace = buf; }
long long* tc_get_fm_trace() { return g_fm_trace; }

// Group size and ring depths for an atom count; false if the kernel's buffers do not fit (the caller falls back).
static bool fm_plan(int V, int VP, int64_t n, int* G, int* wcs, int* scs, int* smem_bytes) {
  (void)V;
  if (VP > 128 || VP < 16) return false;
  int g = kFmMaxN / VP;  // N = G * VP <= 160 TMEM columns per buffer (PT0 | PT1 | DT)
  if (g < 1) g = 1;
  if ((int64_t)g > n) g = (int)(n < 1 ? 1 : n);
  for (int w = 5; w >= 3; w--)
    for (int s = (g > 1 ? 4 : 3); s >= (g > 1 ? g : 1); s--) {  // at least one score image per sample of a group in flight
      const int total = (int)FmSmem(VP, g, w, s).total();
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

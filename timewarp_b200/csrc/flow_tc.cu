// Tensor-core kernels of the conditioner networks (sm_100a: tcgen05.mma + TMEM + bulk async copy).
//
// Arithmetic: every fp32 product a*w is evaluated as hi(a)*hi(w) + lo(a)*hi(w) + hi(a)*lo(w) with
// bf16 hi/lo parts and fp32 accumulation in TMEM ("bf16x3", ~16 significand bits per operand), or
// as hi(a)*hi(w) only ("bf16").  Orientation: tokens on the 128 TMEM lanes (MMA M), output
// features on the columns (MMA N), so bias/activation/LayerNorm run per thread without shuffles.
#include <stdlib.h>

#include "flow_tc.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

// Debug event trace (tw_debug_set_trace): CTA (0,0) records {event | item << 8, clock64} per warp role.
#define TW_TRACE(buf, on, cnt, role, ev, item)                                      \
  if ((on) && (cnt) < 1024) {                                                       \
    (buf)[((role) * 1024 + (cnt)) * 2] = (long long)(ev) | ((long long)(item) << 8); \
    (buf)[((role) * 1024 + (cnt)) * 2 + 1] = clock64();                             \
    (cnt)++;                                                                        \
  }

// ============================================================================================
// packed-weight layout
constexpr int kTileBytes128 = 128 * 128 * 2;  // [128 rows x 128 K] bf16 = two [128 x 64] swizzled K blocks
constexpr int kFfnChunk = 128;                // hidden units per FFN chunk

TcLayout TcLayout::make(const tw_flow_config* c) {
  TcLayout L{};
  L.D = c->d_model, L.F = c->dim_feedforward, L.H = c->num_heads, L.E = c->atom_embedding_dim;
  L.hid = c->mlp_hidden_dims[0], L.T = c->num_transformer_layers;
  size_t off = 0;
  L.in_w1 = off, off += (size_t)2 * L.hid * 128;                    // [hid x 64] hi + lo
  L.in_w2 = off, off += (size_t)2 * (L.hid / 64) * 128 * 128;       // hid/64 K blocks of [128 x 64], hi + lo
  L.out_w1 = off, off += (size_t)2 * 2 * L.hid * 128;               // 2 K blocks of [hid x 64], hi + lo
  L.enc0 = off;
  L.enc_wc = 0;
  size_t wc_bytes = (size_t)2 * (L.H * L.D / 64) * 128 * 128;       // H*D/64 K blocks of [128 x 64], hi + lo
  L.enc_ffn = wc_bytes;
  size_t ffn_bytes = (size_t)(L.F / kFfnChunk) * 4 * kTileBytes128; // per chunk: W1hi, W1lo, W2hi, W2lo
  L.enc_stride = wc_bytes + ffn_bytes;
  off += L.enc_stride * L.T;
  L.net_bytes = align_up(off, 1024);
  return L;
}

uint32_t tc_stage_mask() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("TW_TC_STAGES");
    cached = e ? (atoi(e) & (int)TC_IMPLEMENTED) : (int)TC_IMPLEMENTED;
  }
  return (uint32_t)cached;
}

bool tc_supported(const tw_flow_config* c) {
  return c->d_model == 128 && c->num_mlp_hidden == 1 && c->mlp_hidden_dims[0] == 256 && c->dim_feedforward % kFfnChunk == 0 &&
         c->dim_feedforward >= kFfnChunk && c->atom_embedding_dim + 9 <= 64 && c->num_heads >= 1;
}

size_t tc_packed_bytes(const tw_flow_config* c) {
  TcLayout L = TcLayout::make(c);
  return L.total(c) + (size_t)L.D * L.H * L.D * sizeof(float) + 1024;  // + fp32 scratch for W_c
}

// ---- all weight images in ONE launch ----------------------------------------------------------
// A training step re-packs every weight (the fp32 parameters changed): per-matrix launches are 240 kernels of ~12 us each
// (latency, not bandwidth: 144 MB read, 150 MB written) = 2.75 ms of a 19 ms step.  k_pack_all does the same work with one CTA
// per image tile; the job is decoded from blockIdx.x, the pointers travel in the kernel parameter block (groups of up to
// kPackGroup conditioner networks).  The W_c tiles are computed in place (fp64 accumulation, so that attention becomes sum_h W_c,h (A_h x)).
constexpr int kPackGroup = 16, kPackMaxT = 4;
struct PackGroup {
  const float* in_w1[kPackGroup];
  const float* in_w2[kPackGroup];
  const float* out_w1[kPackGroup];
  const float* wv[kPackGroup][kPackMaxT];
  const float* wo[kPackGroup][kPackMaxT];
  const float* w1[kPackGroup][kPackMaxT];
  const float* w2[kPackGroup][kPackMaxT];
  uint8_t* base[kPackGroup];
  int T, D, F, H, hid, Kin, local;
  int n_fixed, n_wc, n_w1, n_w2, per_net;
  unsigned long long in_w1_off, in_w2_off, out_w1_off, enc0, enc_stride, enc_wc, enc_ffn;
};

__device__ __forceinline__ void pack_tile_dev(const float* __restrict__ W, int ldw, int n_rows, int n_cols, int rows, int tr, int tk,
                                              uint8_t* hi, uint8_t* lo) {
  // whole tile inside the matrix, 16-byte aligned rows: one 16-byte chunk (8 columns) per thread -- two float4 loads, one uint4
  // store per part; eight neighbouring threads cover a 128-byte swizzled row (the 4-byte version below ran at a third of the rate)
  if ((ldw & 3) == 0 && (tr + 1) * rows <= n_rows && (tk + 1) * 64 <= n_cols && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    for (int e = threadIdx.x; e < rows * 8; e += blockDim.x) {
      const int r = e >> 3, ch = e & 7;
      const float4* src = reinterpret_cast<const float4*>(W + (size_t)(tr * rows + r) * ldw + tk * 64 + ch * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      uint4 h, l;
      split2(a.x, a.y, h.x, l.x), split2(a.z, a.w, h.y, l.y), split2(b.x, b.y, h.z, l.z), split2(b.z, b.w, h.w, l.w);
      const uint32_t off = sw128_offset(r, ch * 8, rows);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
    return;
  }
  for (int e = threadIdx.x; e < rows * 32; e += blockDim.x) {
    int r = e >> 5, kp = (e & 31) * 2;
    int gr = tr * rows + r, gc = tk * 64 + kp;
    float a = (gr < n_rows && gc < n_cols) ? W[(size_t)gr * ldw + gc] : 0.f;
    float b = (gr < n_rows && gc + 1 < n_cols) ? W[(size_t)gr * ldw + gc + 1] : 0.f;
    uint32_t h, l;
    split2(a, b, h, l);
    uint32_t off = sw128_offset(r, kp, rows);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
  }
}

__global__ void __launch_bounds__(256) k_pack_all(const __grid_constant__ PackGroup g) {
  const int nb = blockIdx.x / g.per_net;
  int j = blockIdx.x % g.per_net;
  uint8_t* base = g.base[nb];
  const int D = g.D, F = g.F, H = g.H, hid = g.hid;
  if (j < g.n_fixed) {
    if (j == 0) {  // in_mlp L1 [hid x Kin] -> one [hid x 64] tile: hi at 0, lo at hid*128
      pack_tile_dev(g.in_w1[nb], g.Kin, hid, g.Kin, hid, 0, 0, base + g.in_w1_off, base + g.in_w1_off + (size_t)hid * 128);
    } else if (j < 1 + hid / 64) {  // in_mlp L2 [D x hid] -> hid/64 K blocks [128 x 64]: block kb at kb*32 KB: hi 16 KB, lo 16 KB
      const int tk = j - 1;
      uint8_t* hi = base + g.in_w2_off + (size_t)tk * (2 * 128 * 128);
      pack_tile_dev(g.in_w2[nb], hid, D, hid, 128, 0, tk, hi, hi + 128 * 128);
    } else {  // out_mlp L1 [hid x D] -> D/64 K blocks of [hid x 64]: block kb at kb*(2*hid*128): hi, lo
      const int tk = j - 1 - hid / 64;
      uint8_t* hi = base + g.out_w1_off + (size_t)tk * (2 * hid * 128);
      pack_tile_dev(g.out_w1[nb], D, hid, D, hid, 0, tk, hi, hi + (size_t)hid * 128);
    }
    return;
  }
  j -= g.n_fixed;
  const int per_layer = g.n_wc + g.n_w1 + g.n_w2;
  const int t = j / per_layer;
  j %= per_layer;
  uint8_t* eb = base + g.enc0 + (size_t)t * g.enc_stride;
  if (j < g.n_wc) {
    // W_c[i, h*D + c] = sum_k W_o[i, h*D + k] W_v[h*D + k, c] (fp64 accumulate) -> K block kb = j of [D x H*D]:
    // [128 x 64] tile at kb*32 KB: hi 16 KB, lo 16 KB.  A warp shares the row i (broadcast W_o loads, coalesced W_v loads).
    const int kb = j;
    uint8_t* hi = eb + g.enc_wc + (size_t)kb * (2 * 128 * 128);
    uint8_t* lo = hi + 128 * 128;
    const float* __restrict__ Wo = g.wo[nb][t];
    const float* __restrict__ Wv = g.wv[nb][t];
    const int h = (kb * 64) / D, c0 = (kb * 64) % D;
    // four rows per thread at a time: eight independent fp64 chains and one W_v load for four rows (one row at a time left the
    // kernel waiting on the latency of two dependent DFMA chains per thread); same summation order per element
    const int kp = (threadIdx.x & 31) * 2;
    const float* wv = Wv + (size_t)(h * D) * D + c0 + kp;
    for (int r0 = threadIdx.x >> 5; r0 < 128; r0 += 32) {
      double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      const float* wo[4];
#pragma unroll
      for (int u = 0; u < 4; u++) wo[u] = Wo + (size_t)min(r0 + 8 * u, D - 1) * H * D + h * D;
      for (int k = 0; k < D; k++) {
        const float2 v = *reinterpret_cast<const float2*>(wv + (size_t)k * D);
        const double vx = (double)v.x, vy = (double)v.y;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const double o = (double)wo[u][k];
          acc[u][0] += o * vx, acc[u][1] += o * vy;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int r = r0 + 8 * u;
        uint32_t hh, ll;
        split2(r < D ? (float)acc[u][0] : 0.f, r < D ? (float)acc[u][1] : 0.f, hh, ll);
        const uint32_t off = sw128_offset(r, kp, 128);
        *reinterpret_cast<uint32_t*>(hi + off) = hh;
        *reinterpret_cast<uint32_t*>(lo + off) = ll;
      }
    }
    return;
  }
  j -= g.n_wc;
  if (j < g.n_w1) {
    // FFN chunk c: [W1hi | W1lo | W2hi | W2lo], each 32 KB = K blocks kb0, kb1 of [128 x 64];
    // W1 [F x D]: tile (tr = chunk, tk) at chunk*128KB + tk*16KB, lo +32KB
    const int tiles_k = D / 64, tr = j / tiles_k, tk = j % tiles_k;
    uint8_t* hi = eb + g.enc_ffn + (size_t)tr * (4 * kTileBytes128) + (size_t)tk * (128 * 128);
    pack_tile_dev(g.w1[nb][t], D, F, D, 128, tr, tk, hi, hi + kTileBytes128);
    return;
  }
  j -= g.n_w1;
  {  // W2 [D x F]: K block b (columns b*64..) -> chunk b/2, K block b%2 of the W2hi / W2lo tiles
    const int b = j;
    uint8_t* hi = eb + g.enc_ffn + 2 * kTileBytes128 + (size_t)(b >> 1) * 4 * kTileBytes128 + (b & 1) * (128 * 128);
    pack_tile_dev(g.w2[nb][t], F, D, F, 128, 0, b, hi, hi + kTileBytes128);
  }
}

int tc_pack_weights(const tw_flow_config* c, const ParamView& pv, uint8_t* packed, size_t bytes, cudaStream_t st) {
  TW_CHECK_ARG(tc_supported(c), "configuration not supported by the tensor-core path");
  TW_CHECK_ARG(packed && bytes >= tc_packed_bytes(c), "packed buffer too small");
  TW_CHECK_ARG(((uintptr_t)packed & 1023) == 0, "packed buffer must be 1024-byte aligned");
  TW_CHECK_ARG(c->num_transformer_layers <= kPackMaxT, "tensor-core weight packing supports at most %d encoder layers per network", kPackMaxT);
  TcLayout L = TcLayout::make(c);
  PackGroup g{};
  g.T = L.T, g.D = L.D, g.F = L.F, g.H = L.H, g.hid = L.hid, g.Kin = L.E + 9, g.local = pv.local() ? 1 : 0;
  g.n_fixed = 1 + L.hid / 64 + L.D / 64;
  g.n_wc = pv.local() ? 0 : L.H * L.D / 64;  // (local attention runs on the CUDA-core kernels straight from the fp32 weights)
  g.n_w1 = (L.F / 128) * (L.D / 64);
  g.n_w2 = L.F / 64;
  g.per_net = g.n_fixed + L.T * (g.n_wc + g.n_w1 + g.n_w2);
  g.in_w1_off = L.in_w1, g.in_w2_off = L.in_w2, g.out_w1_off = L.out_w1;
  g.enc0 = L.enc0, g.enc_stride = L.enc_stride, g.enc_wc = L.enc_wc, g.enc_ffn = L.enc_ffn;
  const int n_nets = c->num_coupling_layers * 2;
  for (int first = 0; first < n_nets; first += kPackGroup) {
    const int count = n_nets - first < kPackGroup ? n_nets - first : kPackGroup;
    for (int i = 0; i < count; i++) {
      const int k = (first + i) / 2, net = (first + i) % 2;
      g.base[i] = packed + L.net_offset(k, net);
      g.in_w1[i] = pv.in_w(k, net, 0), g.in_w2[i] = pv.in_w(k, net, 1), g.out_w1[i] = pv.out_w(k, net, 0);
      for (int t = 0; t < L.T; t++) {
        g.wv[i][t] = pv.enc(k, net, t, 0), g.wo[i][t] = pv.enc(k, net, t, 2);
        g.w1[i][t] = pv.enc(k, net, t, 3), g.w2[i][t] = pv.enc(k, net, t, 5);
      }
    }
    k_pack_all<<<(unsigned)(count * g.per_net), 256, 0, st>>>(g);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

// Row epilogue shared by the token-major kernels: v[128] (+ bias) + residual row -> LayerNorm -> global
__device__ __forceinline__ void ln_store_row(float (&v)[128], const float* __restrict__ bias, const float* __restrict__ resid_row,
                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                             float* __restrict__ out_row, float* __restrict__ pre_row = nullptr) {
  const float4* xr = reinterpret_cast<const float4*>(resid_row);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    float4 xv = __ldg(xr + j);
    if (bias) {
      float4 bv = *reinterpret_cast<const float4*>(bias + 4 * j);
      xv.x += bv.x, xv.y += bv.y, xv.z += bv.z, xv.w += bv.w;
    }
    v[4 * j + 0] += xv.x;
    v[4 * j + 1] += xv.y;
    v[4 * j + 2] += xv.z;
    v[4 * j + 3] += xv.w;
    sum += (v[4 * j] + v[4 * j + 1]) + (v[4 * j + 2] + v[4 * j + 3]);
    if (pre_row) reinterpret_cast<float4*>(pre_row)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  const float mean = sum * (1.f / 128.f);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 128; j++) {
    float d = v[j] - mean;
    sq = fmaf(d, d, sq);
  }
  const float rstd = 1.0f / sqrtf(sq * (1.f / 128.f) + eps);
  float4* orow = reinterpret_cast<float4*>(out_row);
#pragma unroll
  for (int j = 0; j < 32; j++) {
    float4 g4 = *reinterpret_cast<const float4*>(gamma + 4 * j);
    float4 b4 = *reinterpret_cast<const float4*>(beta + 4 * j);
    float4 o;
    o.x = (v[4 * j + 0] - mean) * rstd * g4.x + b4.x;
    o.y = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
    o.z = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
    o.w = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
    orow[j] = o;
  }
}

// ============================================================================================
// Fused FFN:  out = LayerNorm(x + W2 relu(W1 x + b1) + b2)         custom_attention_encoder.py:111-113
//
// Persistent CTAs (grid.x per network, grid.y = network), 128-token tiles; each CTA works through ONE stream
// of (tile, 128-hidden-unit chunk) items with no drain between tiles.  Warp roles:
//   warp 0   : bulk-copy producer -- streams 32 KB weight tiles (W1hi, W1lo, W2hi, W2lo per chunk) through a
//              shared-memory ring
//   warp 1   : one elected lane issues tcgen05.mma.  BOTH GEMMs take their A operand from TMEM (TS form: an smem A
//              operand costs ~32 extra cycles per MMA, tools/umma_timing.py):
//                  G1(g): DH[g&1]  = X W1c^T          A = bf16 hi/lo images of the x tile in TMEM
//                  G2(g): Y       += H_c W2c^T        A = H_c, written IN PLACE over DH[g&1] by the epilogue
//              issue order G1(g), G2(g-1): the tensor pipe executes in issue order, so DH[g&1] is not overwritten by
//              G1(g+2) before G2(g) has read it, and the chunk epilogue of g overlaps G2(g-1) + G1(g+1).
//   warps 2-5: group 0 -- chunk epilogue for hidden units [0,64) of every chunk (+b1, ReLU, hi/lo split, in place),
//              LayerNorm epilogue of the previous tile (Y already holds x + b2 + FFN(x))
//   warps 6-9: group 1 -- chunk epilogue for hidden units [64,128); stages the next x tile (fp32) in shared memory
//              while the chunks run, writes its hi/lo images to TMEM as soon as the last G1 of the tile has retired,
//              and initialises Y with x + b2 once the LayerNorm epilogue has drained the previous tile.
// The 128x2048 hidden activation never leaves the SM.  TMEM (512 columns): X hi 64 | X lo 64 | Y 128 | DH0 128 | DH1 128.
constexpr int kFfnStages = 4;
constexpr int kStageBytes = kTileBytes128;
constexpr int kFfnThreads = 320;
constexpr uint32_t TM_X = 0, TM_Y = 128, TM_DH = 256;
constexpr int kXsRow = 528;  // bytes per staged x row: 512 + 16 padding (conflict-free row-per-thread 16-byte reads)

struct FfnArgs {
  const float* x[2];       // [M,128] layer input (post-LN1)
  float* out[2];           // [M,128]
  float* pre[2];           // optional [M,128]: the pre-LayerNorm sum x + FFN(x) (saved for the training backward)
  const uint8_t* w[2];     // packed FFN chunks of this layer (per net)
  const float* b1[2];
  const float* b2[2];
  const float* gamma[2];
  const float* beta[2];
  int64_t M;
  int F;
  float eps;
  int dbg;  // TW_FFN_DBG experiments (timing only, results invalid): 1 no weight traffic, 2 no chunk-epilogue math, 4 hi-only MMAs
  long long* trace;  // optional event trace of CTA (0,0) (tw_debug_set_ffn_trace): [role][1024][2] = {event | item << 8, clock64}
  float* tail[2];    // pair kernel: split-tile partial sums [tile][rank][128][128] (nullptr: leftover tiles are not split)
  int* tail_cnt[2];  // arrival counters [tile][rank]
};

struct FfnSmem {
  static constexpr int RING = 0;
  static constexpr int XS = RING + kFfnStages * kStageBytes;       // staged fp32 x tile, 128 rows of kXsRow bytes
  static constexpr int B1 = XS + 128 * kXsRow;                     // F floats (<= 4096)
  static constexpr int VEC = B1 + 4096 * 4;                        // b2, gamma, beta: 3*128 floats
  static constexpr int BARS = VEC + 3 * 128 * 4;
  static constexpr int TOTAL = BARS + 256;
};

template <int kSplit>
__global__ void __launch_bounds__(kFfnThreads, 1) k_ffn_tc(FfnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_chunks = a.F / kFfnChunk;
  const int64_t n_tiles = (a.M + 127) / 128;
  const int64_t my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FfnSmem::BARS);
  uint64_t* full = bars;                      // [kFfnStages]
  uint64_t* empty = bars + kFfnStages;        // [kFfnStages]
  uint64_t* d1_full = bars + 2 * kFfnStages;  // [2]   G1 into DH[b] has retired
  uint64_t* h_full = d1_full + 2;             // [4]   (buffer b, K half hf) of H written: index b*2+hf, 128 arrivals
  uint64_t* x_full = h_full + 4;              // X images of the next tile are in TMEM (128 arrivals)
  uint64_t* x_free = x_full + 1;              // last G1 of the tile has retired
  uint64_t* y_full = x_free + 1;              // last G2 of the tile has retired
  uint64_t* y_free = y_full + 1;              // LayerNorm epilogue has read Y (128 arrivals)
  uint64_t* y_init = y_free + 1;              // Y = x + b2 of the next tile written (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_init + 1);

  float* b1s = reinterpret_cast<float*>(smem + FfnSmem::B1);
  float* vecs = reinterpret_cast<float*>(smem + FfnSmem::VEC);

  if (tid == 0) {
    for (int i = 0; i < kFfnStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) mbar_init(&d1_full[i], 1);
    for (int i = 0; i < 4; i++) mbar_init(&h_full[i], 128);
    mbar_init(x_full, 128);
    mbar_init(x_free, 1);
    mbar_init(y_full, 1);
    mbar_init(y_free, 128);
    mbar_init(y_init, 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < a.F; i += blockDim.x) b1s[i] = a.b1[net][i];
  for (int i = tid; i < 128; i += blockDim.x) {
    vecs[i] = a.b2[net][i];
    vecs[128 + i] = a.gamma[net][i];
    vecs[256 + i] = a.beta[net][i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t G = my_tiles * n_chunks;  // (tile, chunk) items of this CTA

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (whole warp, one elected lane issues)
    uint32_t stage = 0, phase = 0;
    int n_loaded = 0;
    const uint8_t* wbase = a.w[net];
    auto load = [&](int op, int chunk) {  // op 0: W1 tiles of the chunk, op 1: W2 tiles
      for (int part = 0; part < (kSplit == 3 ? 2 : 1); part++) {
        const uint8_t* src = wbase + (size_t)chunk * 4 * kTileBytes128 + (size_t)(op * 2 + part) * kTileBytes128;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          if ((a.dbg & 1) && n_loaded >= kFfnStages) {
            mbar_arrive(&full[stage]);
          } else {
            mbar_arrive_expect_tx(&full[stage], kStageBytes);
            bulk_g2s(smem + FfnSmem::RING + stage * kStageBytes, src, kStageBytes, &full[stage]);
          }
        }
        __syncwarp();
        n_loaded++;
        if (++stage == kFfnStages) stage = 0, phase ^= 1;
      }
    };
    for (int64_t it = 0; it < my_tiles; it++)  // consumption order: G1(g), G2(g-1); no 64-bit div/mod in the loops
      for (int c = 0; c < n_chunks; c++) {
        load(0, c);
        if (it > 0 || c > 0) load(1, c > 0 ? c - 1 : n_chunks - 1);
      }
    if (G > 0) load(1, n_chunks - 1);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one elected lane issues)
    uint32_t stage = 0, phase = 0;
    uint32_t ph_x = 0, ph_h = 0 /* bit b*2+hf */, ph_yinit = 0;
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t ring = smem_u32(smem + FfnSmem::RING);
    const uint32_t x_hi = tmem + TM_X, x_lo = tmem + TM_X + 64;

    auto issue_g2 = [&](int64_t gp, int j) {  // Y += H(gp) W2c^T; j = chunk index of item gp
      const int buf = (int)(gp & 1);
      const uint32_t st_hi = stage;
      mbar_wait(&full[stage], phase);
      if (++stage == kFfnStages) stage = 0, phase ^= 1;
      uint32_t st_lo = st_hi;
      if (kSplit == 3) {
        st_lo = stage;
        mbar_wait(&full[stage], phase);
        if (++stage == kFfnStages) stage = 0, phase ^= 1;
      }
      if (j == 0) {  // Y of this tile has been initialised with x + b2 (after the previous tile's Y was read out)
        mbar_wait(y_init, ph_yinit);
        ph_yinit ^= 1;
      }
      const uint32_t whi = ring + st_hi * kStageBytes, wlo = ring + st_lo * kStageBytes;
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        mbar_wait(&h_full[buf * 2 + hf], (ph_h >> (buf * 2 + hf)) & 1u);  // epilogue group hf has written its half of H
        ph_h ^= 1u << (buf * 2 + hf);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t h_hi = tmem + TM_DH + buf * 128 + hf * 64, h_lo = h_hi + 32;
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ts(tmem + TM_Y, h_hi + k * 8, desc_kmajor_sw128(whi + hf * 16384 + k * 32), idesc, 1);
          if (kSplit == 3 && !(a.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts(tmem + TM_Y, h_lo + k * 8, desc_kmajor_sw128(whi + hf * 16384 + k * 32), idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts(tmem + TM_Y, h_hi + k * 8, desc_kmajor_sw128(wlo + hf * 16384 + k * 32), idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        mma_commit(&empty[st_hi]);
        if (kSplit == 3) mma_commit(&empty[st_lo]);
        if (j == n_chunks - 1) mma_commit(y_full);
      }
      __syncwarp();
    };

    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; it++) {
      mbar_wait(x_full, ph_x);
      ph_x ^= 1;
      tc_fence_after();
      for (int c = 0; c < n_chunks; c++, g++) {
        const uint32_t d1 = tmem + TM_DH + (uint32_t)(g & 1) * 128;
        mbar_wait(&full[stage], phase);  // W1 hi tile: Xhi*W1hi, Xlo*W1hi
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wt = ring + stage * kStageBytes;
#pragma unroll
          for (int k = 0; k < 8; k++)
            mma_ts(d1, x_hi + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 16384 + (k & 3) * 32), idesc, k > 0);
          if (kSplit == 3 && !(a.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 8; k++) mma_ts(d1, x_lo + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 16384 + (k & 3) * 32), idesc, 1);
          }
          mma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kFfnStages) stage = 0, phase ^= 1;
        if (kSplit == 3) {  // W1 lo tile: Xhi*W1lo
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t wt = ring + stage * kStageBytes;
            if (!(a.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 8; k++) mma_ts(d1, x_hi + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 16384 + (k & 3) * 32), idesc, 1);
            }
            mma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == kFfnStages) stage = 0, phase ^= 1;
        }
        if (elect_one()) {
          mma_commit(&d1_full[g & 1]);
          if (c == n_chunks - 1) mma_commit(x_free);  // every read of the X images of this tile has retired
        }
        __syncwarp();
        if (g >= 1) issue_g2(g - 1, c > 0 ? c - 1 : n_chunks - 1);
      }
    }
    if (G > 0) issue_g2(G - 1, n_chunks - 1);
  } else {
    // ------------------------------------------------------------------ epilogue warps (2 groups x 128 threads)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hf = (warp - 2) >> 2;         // epilogue group = K half of H this thread produces
    const int row = q * 32 + lane;          // token row inside the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_d1 = 0 /* bit per buffer */, ph_xfree = 0, ph_y = 0, ph_yfree = 0;
    const float* x = a.x[net];
    float* out = a.out[net];
    uint8_t* xs = smem + FfnSmem::XS;
    const uint8_t* xs_row = xs + row * kXsRow;

    // ---- group 1 helpers: staging of an x tile, TMEM images, Y initialisation
    auto load_rows = [&](int64_t tile, int r0, float4 (&v)[4]) {  // warp q stages rows q*32 + r0 .. +3 (coalesced 512 B each)
      const int64_t row0 = tile * 128 + q * 32 + r0;
#pragma unroll
      for (int u = 0; u < 4; u++)
        v[u] = (row0 + u < a.M) ? __ldg(reinterpret_cast<const float4*>(x + (row0 + u) * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_rows = [&](int r0, const float4 (&v)[4]) {
#pragma unroll
      for (int u = 0; u < 4; u++) *reinterpret_cast<float4*>(xs + (q * 32 + r0 + u) * kXsRow + lane * 16) = v[u];
    };
    auto group1_sync = [&]() { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    auto x_to_tmem = [&]() {  // own row of the staged tile -> bf16 hi/lo A-operand images in TMEM
#pragma unroll 1
      for (int b = 0; b < 4; b++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 v = *reinterpret_cast<const float4*>(xs_row + b * 128 + j * 16);
          split2(v.x, v.y, hi[2 * j], lo[2 * j]);
          split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        tmem_st16(tmem + lane_base + TM_X + b * 16, hi);
        if (kSplit == 3) tmem_st16(tmem + lane_base + TM_X + 64 + b * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x_full);
    };
    auto init_y = [&]() {  // Y <- x + b2: the residual and the second bias ride in the accumulator
#pragma unroll 1
      for (int b = 0; b < 8; b++) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float4 v = *reinterpret_cast<const float4*>(xs_row + b * 64 + j * 16);
          const float4 bv = *reinterpret_cast<const float4*>(vecs + b * 16 + 4 * j);
          r[4 * j] = __float_as_uint(v.x + bv.x), r[4 * j + 1] = __float_as_uint(v.y + bv.y);
          r[4 * j + 2] = __float_as_uint(v.z + bv.z), r[4 * j + 3] = __float_as_uint(v.w + bv.w);
        }
        tmem_st16(tmem + lane_base + TM_Y + b * 16, r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(y_init);
    };
    // ---- group 0 helper: LayerNorm epilogue of a finished tile (two passes over Y in TMEM)
    auto layer_norm_tile = [&](int64_t tile) {
      mbar_wait(y_full, ph_y);
      ph_y ^= 1;
      tc_fence_after();
      const int64_t grow = tile * 128 + row;
      const bool valid = grow < a.M;
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int g4 = 0; g4 < 4; g4++) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + TM_Y + g4 * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float v = __uint_as_float(r[j]);
          sum += v;
          sq = fmaf(v, v, sq);
        }
      }
      const float mean = sum * (1.f / 128.f);
      const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
      const float rstd = 1.0f / sqrtf(var + a.eps);
      float4* orow = reinterpret_cast<float4*>(out + (valid ? grow : 0) * 128);
      float4* prow = a.pre[net] ? reinterpret_cast<float4*>(a.pre[net] + (valid ? grow : 0) * 128) : nullptr;
#pragma unroll 1
      for (int g4 = 0; g4 < 4; g4++) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + TM_Y + g4 * 32, r);
        tmem_ld_wait();
        if (g4 == 3) {  // Y has been read for the last time: the next tile may initialise it
          tc_fence_before();
          mbar_arrive(y_free);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 gm = *reinterpret_cast<const float4*>(vecs + 128 + g4 * 32 + 4 * j);
          const float4 bt = *reinterpret_cast<const float4*>(vecs + 256 + g4 * 32 + 4 * j);
          float4 p, o;
          p.x = __uint_as_float(r[4 * j]), p.y = __uint_as_float(r[4 * j + 1]);
          p.z = __uint_as_float(r[4 * j + 2]), p.w = __uint_as_float(r[4 * j + 3]);
          o.x = (p.x - mean) * rstd * gm.x + bt.x;
          o.y = (p.y - mean) * rstd * gm.y + bt.y;
          o.z = (p.z - mean) * rstd * gm.z + bt.z;
          o.w = (p.w - mean) * rstd * gm.w + bt.w;
          if (valid) orow[g4 * 8 + j] = o;
          if (valid && prow) prow[g4 * 8 + j] = p;
        }
      }
    };

    if (hf == 1 && my_tiles > 0) {  // first tile: stage, images, (Y is initialised after the first chunk epilogue)
      for (int r0 = 0; r0 < 32; r0 += 4) {
        float4 v[4];
        load_rows(blockIdx.x, r0, v);
        store_rows(r0, v);
      }
      group1_sync();
      x_to_tmem();
    }
    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; it++) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const bool has_next = it + 1 < my_tiles;
      int staged = 0;  // rows (per warp) of the next tile already staged
      float4 pf[4];
      bool pf_valid = false;
      for (int c = 0; c < n_chunks; c++, g++) {
        const int buf = (int)(g & 1);
        if (hf == 1 && n_chunks > 1 && c == n_chunks - 1 && has_next) {
          // the last G1 of this tile and the release of the X images retire together: publish the next tile's images
          // FIRST, so that its first GEMM is queued behind G2 of the previous chunk without a bubble
          while (staged < 32 || pf_valid) {
            if (pf_valid) store_rows(staged - 4, pf), pf_valid = false;
            if (staged < 32) load_rows(tile + gridDim.x, staged, pf), staged += 4, pf_valid = true;
          }
          group1_sync();
          mbar_wait(x_free, ph_xfree);
          ph_xfree ^= 1;
          tc_fence_after();
          x_to_tmem();
        }
        // ---- DH[buf][:, hf*64 .. +64): + b1, ReLU, hi/lo split, written back in place as the A operand of G2
        mbar_wait(&d1_full[buf], (ph_d1 >> buf) & 1u);
        ph_d1 ^= 1u << buf;
        tc_fence_after();
        if (!(a.dbg & 2)) {
          const float* bc = b1s + c * kFfnChunk + hf * 64;
          const uint32_t base = tmem + lane_base + TM_DH + buf * 128 + hf * 64;
          uint32_t r0[32], r1[32];
          tmem_ld32(base, r0);
          tmem_ld32(base + 32, r1);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 bb = *reinterpret_cast<const float2*>(bc + j);
            split2(fmaxf(__uint_as_float(r0[j]) + bb.x, 0.f), fmaxf(__uint_as_float(r0[j + 1]) + bb.y, 0.f), hi[j >> 1], lo[j >> 1]);
          }
          tmem_st16(base, hi);
          if (kSplit == 3) tmem_st16(base + 32, lo);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 bb = *reinterpret_cast<const float2*>(bc + 32 + j);
            split2(fmaxf(__uint_as_float(r1[j]) + bb.x, 0.f), fmaxf(__uint_as_float(r1[j + 1]) + bb.y, 0.f), hi[j >> 1], lo[j >> 1]);
          }
          tmem_st16(base + 16, hi);
          if (kSplit == 3) tmem_st16(base + 48, lo);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&h_full[buf * 2 + hf]);

        if (c == 0) {
          if (hf == 0) {
            if (it > 0) layer_norm_tile(tile - gridDim.x);
          } else {
            if (it > 0) {
              mbar_wait(y_free, ph_yfree);
              ph_yfree ^= 1;
              tc_fence_after();
            }
            init_y();
            group1_sync();  // every row of the staging buffer has been consumed: it may be refilled
          }
        }
        if (hf == 1 && has_next && c < n_chunks - 1) {  // stage 4 rows per warp per chunk, loads in flight across a chunk
          if (pf_valid) store_rows(staged - 4, pf), pf_valid = false;
          if (staged < 32) load_rows(tile + gridDim.x, staged, pf), staged += 4, pf_valid = true;
        }
      }
      if (n_chunks == 1 && hf == 1 && has_next) {  // single-chunk FFN: nothing overlapped, stage and publish here
        for (int r0 = 0; r0 < 32; r0 += 4) {
          float4 v[4];
          load_rows(tile + gridDim.x, r0, v);
          store_rows(r0, v);
        }
        group1_sync();
        mbar_wait(x_free, ph_xfree);
        ph_xfree ^= 1;
        tc_fence_after();
        x_to_tmem();
      }
    }
    if (hf == 0 && my_tiles > 0) layer_norm_tile(blockIdx.x + (my_tiles - 1) * gridDim.x);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ============================================================================================
// Fused FFN on CTA PAIRS (cta_group::2).  Same pipeline as k_ffn_tc, with one tcgen05.mma spanning the two SMs of a
// cluster: 256-token tiles (128 rows in each CTA's TMEM), the B operand (weight chunk) split over the pair -- each CTA
// streams only HALF of every weight tile (rows [64r, 64r+64) of each [128 x 64] K block), which halves the L2 -> SM
// traffic (at full tensor rate the single-CTA kernel needs ~10 TB/s of L2 reads, the measured cap) and makes the
// 128 KB ring hold two chunks instead of one.  Rank 0 issues every MMA; commits are multicast to both CTAs; the
// epilogue threads of both CTAs arrive on rank 0's barriers; rank 1's MMA warp relays its "weights landed" signals.
constexpr int kPairStages = 8;
constexpr int kPairThreads = 352;       // warp 0 weight producer, 1 MMA issuer / relay, 2-9 epilogue, 10 activation I/O
constexpr int kPairStageBytes = 16384;  // two [64 rows x 64 K] K blocks of one weight tile part
constexpr int kPairMaxF = 2048;         // dim_feedforward limit of the pair kernel (b1 lives in shared memory)

struct FfnPairSmem {
  static constexpr int RING = 0;
  static constexpr int XS = RING + kPairStages * kPairStageBytes;
  static constexpr int STAT = XS + 128 * kXsRow;                  // per-row partial (sum, sum of squares) of each group
  static constexpr int B1 = STAT + 2 * 128 * 8;                   // F floats (<= kPairMaxF)
  static constexpr int VEC = B1 + kPairMaxF * 4;
  static constexpr int BARS = VEC + 3 * 128 * 4;
  static constexpr int TOTAL = BARS + 512;
};
static_assert(FfnPairSmem::TOTAL + 1024 <= 232448, "pair FFN shared memory exceeds 227 KB");

template <int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) k_ffn_pair(FfnArgs a) {  // 11 warps: 3 on one SMSP -> 168 registers
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_chunks = a.F / kFfnChunk;
  const int64_t n_ptiles = (a.M + 255) / 256;  // 256-token tiles of the pair
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // Work schedule.  Every pair runs R = n_ptiles / n_pairs full tiles; the Lft leftover tiles (7.03 rounds would cost 8)
  // are SPLIT along the hidden dimension: S pairs each run n_chunks / S chunks of a leftover tile, add their partial
  // accumulator to a global scratch tile, and the last pair to arrive normalises and stores it.
  const int64_t R = n_ptiles / n_pairs;
  const int Lft = (int)(n_ptiles - R * n_pairs);
  int S = 1;
  if (a.tail[net] != nullptr && Lft > 0 && 2 * Lft <= (int)n_pairs && Lft <= kFfnTailTiles)
    for (int d = n_chunks; d >= 2; d--)
      if (n_chunks % d == 0 && (int64_t)d * Lft <= n_pairs) {
        S = d;
        break;
      }
  const bool split = S > 1;
  const int cps = n_chunks / S;  // chunks per part
  const bool has_partial = split && pair < (int64_t)Lft * S;
  const int part_l = has_partial ? (int)pair / S : 0, part_p = has_partial ? (int)pair % S : 0;
  const int pc0 = part_p * cps, pc1 = pc0 + cps;
  const int64_t my_regular = split ? R : ((pair < n_ptiles) ? (n_ptiles - pair + n_pairs - 1) / n_pairs : 0);
  const int64_t my_tiles = my_regular + (has_partial ? 1 : 0);  // tile iterations of this pair (the partial one is last)
  auto tile_index = [&](int64_t it) -> int64_t { return it < my_regular ? pair + it * n_pairs : R * n_pairs + part_l; };
  auto c_begin = [&](int64_t it) -> int { return it < my_regular ? 0 : pc0; };
  auto c_end = [&](int64_t it) -> int { return it < my_regular ? n_chunks : pc1; };
  int* tail_flag = reinterpret_cast<int*>(smem + FfnPairSmem::BARS + 448);  // 1: this CTA normalises the split tile
  constexpr int kParts = (kSplit == 3) ? 2 : 1;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define FFN_TRACE(role, ev, item)                                                   \
  if (tr_on && tr_n < 1024) {                                                       \
    a.trace[((role) * 1024 + tr_n) * 2] = (long long)(ev) | ((long long)(item) << 8); \
    a.trace[((role) * 1024 + tr_n) * 2 + 1] = clock64();                            \
    tr_n++;                                                                         \
  }

  if (tr_on && warp == 1) {  // kernel entry / exit stamps of CTA (0,0): slots 4..7 after the event rows
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * 2048 + 4] = clock64(), a.trace[3 * 2048 + 5] = gt;
  }
  if (a.trace != nullptr && tid == 0) {  // ... and of EVERY CTA (global timer): slots 8 + 2 * cta, + 1
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * 2048 + 8 + 2 * (blockIdx.y * gridDim.x + blockIdx.x)] = gt;
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FfnPairSmem::BARS);
  uint64_t* full = bars;                         // [8] own half of a weight tile part has landed
  uint64_t* peer_full = full + kPairStages;      // [8] (rank 0) rank 1's half has landed
  uint64_t* empty = peer_full + kPairStages;     // [8] multicast commit
  uint64_t* d1_full = empty + kPairStages;       // [2] multicast commit
  uint64_t* h_full = d1_full + 2;                // [4] (rank 0) 256 arrivals: both CTAs' epilogue group hf
  uint64_t* x_full = h_full + 4;                 // (rank 0) 256 arrivals
  uint64_t* x_free = x_full + 1;                 // multicast commit
  uint64_t* y_full = x_free + 1;                 // multicast commit
  uint64_t* y_free = y_full + 1;                 // (unused)
  uint64_t* y_init = y_free + 1;                 // (rank 0) 512 arrivals: every epilogue thread of both CTAs
  uint64_t* xs_full = y_init + 1;                // local: the staged x tile has landed (bulk copies, tx bytes)
  uint64_t* ln_staged = xs_full + 1;             // local, 256 arrivals: Y initialised from the staged rows and the previous
                                                 // tile's LayerNorm output parked in the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_staged + 1);

  float* b1s = reinterpret_cast<float*>(smem + FfnPairSmem::B1);
  float* vecs = reinterpret_cast<float*>(smem + FfnPairSmem::VEC);

  if (tid == 0) {
    for (int i = 0; i < kPairStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&peer_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) mbar_init(&d1_full[i], 1);
    for (int i = 0; i < 4; i++) mbar_init(&h_full[i], 256);
    mbar_init(x_full, 256);
    mbar_init(x_free, 1);
    mbar_init(y_full, 1);
    mbar_init(y_free, 128);
    mbar_init(y_init, 512);
    mbar_init(xs_full, 1);
    mbar_init(ln_staged, 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  for (int i = tid; i < a.F; i += blockDim.x) b1s[i] = a.b1[net][i];
  for (int i = tid; i < 128; i += blockDim.x) {
    vecs[i] = a.b2[net][i];
    vecs[128 + i] = a.gamma[net][i];
    vecs[256 + i] = a.beta[net][i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t G = my_regular * n_chunks + (has_partial ? cps : 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: this CTA's half of every weight tile part
    uint32_t stage = 0, phase = 0;
    const uint8_t* wbase = a.w[net] + (size_t)rank * 8192;
    auto load = [&](int op, int chunk) {
      for (int part = 0; part < kParts; part++) {
        const uint8_t* src = wbase + (size_t)chunk * 4 * kTileBytes128 + (size_t)(op * 2 + part) * kTileBytes128;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = smem + FfnPairSmem::RING + stage * kPairStageBytes;
          mbar_arrive_expect_tx(&full[stage], kPairStageBytes);
          bulk_g2s(dst, src, 8192, &full[stage]);
          bulk_g2s(dst + 8192, src + 16384, 8192, &full[stage]);
        }
        __syncwarp();
        if (++stage == kPairStages) stage = 0, phase ^= 1;
      }
    };
    bool have_prev = false;
    int prev_c = 0;
    for (int64_t it = 0; it < my_tiles; it++)  // consumption order: G1(item), G2(previous item)
      for (int c = c_begin(it); c < c_end(it); c++) {
        load(0, c);
        if (have_prev) load(1, prev_c);
        have_prev = true, prev_c = c;
      }
    if (have_prev) load(1, prev_c);
  } else if (warp == 10) {
    // ------------------------------------------------------------------ activation I/O warp
    // The staging buffer XS (128 padded fp32 rows) carries the x tile of the NEXT 256-token tile in (per-row bulk loads)
    // and, between the accumulator hand-over of a tile boundary and that refill, the LayerNorm output of the PREVIOUS tile
    // out (per-row bulk stores): the epilogue warps never touch global memory, and the 9.5 MB store burst of 148 CTAs
    // crossing a tile boundary in lock-step drains through the TMA queues without stalling them or the weight producer.
    uint32_t ph_staged = 0;
    auto tile_row0 = [&](int64_t it) -> int64_t { return (tile_index(it) * 2 + rank) * 128; };
    auto valid_rows = [&](int64_t row0) -> int {
      const int64_t left = a.M - row0;
      return left >= 128 ? 128 : (left > 0 ? (int)left : 0);
    };
    auto stage_x = [&](int64_t it) {
      const int64_t row0 = tile_row0(it);
      const int n_valid = valid_rows(row0);
      if (lane == 0) {
        mbar_arrive_expect_tx(xs_full, (uint32_t)n_valid * 512u);
        const float* src = a.x[net] + row0 * 128;
        for (int r = 0; r < n_valid; r++) bulk_g2s(smem + FfnPairSmem::XS + r * kXsRow, src + (size_t)r * 128, 512, xs_full);
      }
      __syncwarp();
    };
    auto store_out = [&](int64_t it) {  // parked LayerNorm rows of tile `it` -> global
      const int64_t row0 = tile_row0(it);
      const int n_valid = valid_rows(row0);
      if (lane == 0) {
        float* dst = a.out[net] + row0 * 128;
        for (int r = 0; r < n_valid; r++) bulk_s2g(dst + (size_t)r * 128, smem + FfnPairSmem::XS + r * kXsRow, 512);
        bulk_commit_group();
      }
      __syncwarp();
    };
    if (my_tiles > 0) stage_x(0);
    for (int64_t it = 0; it < my_tiles; it++) {
      mbar_wait(ln_staged, ph_staged);  // boundary of tile `it` done: Y initialised, LayerNorm(it - 1) parked
      ph_staged ^= 1;
      if (it > 0) store_out(it - 1);
      if (lane == 0) bulk_wait_group_read0();  // the parked rows have been read out of shared memory
      __syncwarp();
      if (it + 1 < my_tiles) {
        stage_x(it + 1);
      } else {  // nothing more to stage: just tell the epilogue warps that the buffer is free for the last LayerNorm
        if (lane == 0) mbar_arrive(xs_full);
        __syncwarp();
      }
    }
    if (my_tiles > 0) {
      mbar_wait(ln_staged, ph_staged);  // LayerNorm of the last tile parked (split tile: stored by the epilogue warps themselves)
      if (!has_partial) store_out(my_tiles - 1);
    }
    if (lane == 0) bulk_wait_group0();
    __syncwarp();
  } else if (warp == 1 && rank != 0) {
    // ------------------------------------------------------------------ rank 1: relay "my half has landed" to rank 0
    uint32_t stage = 0, phase = 0;
    const int64_t n_parts = G * 2 * kParts;
    for (int64_t n = 0; n < n_parts; n++) {
      mbar_wait(&full[stage], phase);
      if (elect_one()) mbar_arrive_cluster(&peer_full[stage], 0);
      __syncwarp();
      if (++stage == kPairStages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ rank 0: MMA issuer for the pair
    uint32_t stage = 0, phase = 0;
    uint32_t ph_x = 0, ph_h = 0 /* bit b*2+hf */, ph_yinit = 0;
    const uint32_t idesc = make_idesc_bf16(256, 128, 0, 0);
    const uint32_t ring = smem_u32(smem + FfnPairSmem::RING);
    const uint32_t x_hi = tmem + TM_X, x_lo = tmem + TM_X + 64;
    auto wait_stage = [&]() -> uint32_t {  // both halves of the next weight tile part are in shared memory
      mbar_wait(&full[stage], phase);
      mbar_wait(&peer_full[stage], phase);
      const uint32_t addr = ring + stage * kPairStageBytes;
      if (++stage == kPairStages) stage = 0, phase ^= 1;
      return addr;
    };
    auto stage_bar = [&](uint32_t addr) -> uint64_t* { return &empty[(addr - ring) / kPairStageBytes]; };

    auto issue_g2 = [&](int64_t gp, bool first, bool last) {  // Y += H(gp) W2c^T; first / last item of its tile iteration
      const int buf = (int)(gp & 1);
      const uint32_t whi = wait_stage();
      const uint32_t wlo = (kSplit == 3) ? wait_stage() : whi;
      FFN_TRACE(0, 3, gp);
      if (first) {
        mbar_wait(y_init, ph_yinit);
        ph_yinit ^= 1;
      }
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        mbar_wait(&h_full[buf * 2 + hf], (ph_h >> (buf * 2 + hf)) & 1u);
        ph_h ^= 1u << (buf * 2 + hf);
        tc_fence_after();
        FFN_TRACE(0, 4 + hf, gp);
        if (elect_one()) {
          const uint32_t h_hi = tmem + TM_DH + buf * 128 + hf * 64, h_lo = h_hi + 32;
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ts_pair(tmem + TM_Y, h_hi + k * 8, desc_kmajor_sw128(whi + hf * 8192 + k * 32), idesc, 1);
          if (kSplit == 3 && !(a.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts_pair(tmem + TM_Y, h_lo + k * 8, desc_kmajor_sw128(whi + hf * 8192 + k * 32), idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts_pair(tmem + TM_Y, h_hi + k * 8, desc_kmajor_sw128(wlo + hf * 8192 + k * 32), idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        mma_commit_pair(stage_bar(whi));
        if (kSplit == 3) mma_commit_pair(stage_bar(wlo));
        if (last) mma_commit_pair(y_full);
      }
      __syncwarp();
    };

    int64_t g = 0;
    if (tr_on) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      a.trace[3 * 2048] = clock64(), a.trace[3 * 2048 + 1] = gt;
    }
    bool prev_first = false, prev_last = false;
    for (int64_t it = 0; it < my_tiles; it++) {
      mbar_wait(x_full, ph_x);
      ph_x ^= 1;
      tc_fence_after();
      const int cb = c_begin(it), ce = c_end(it);
      for (int c = cb; c < ce; c++, g++) {
        const uint32_t d1 = tmem + TM_DH + (uint32_t)(g & 1) * 128;
        FFN_TRACE(0, 0, g);
        {
          const uint32_t wt = wait_stage();  // W1 hi: Xhi*W1hi, Xlo*W1hi
          FFN_TRACE(0, 1, g);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; k++)
              mma_ts_pair(d1, x_hi + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 8192 + (k & 3) * 32), idesc, k > 0);
            if (kSplit == 3 && !(a.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 8; k++) mma_ts_pair(d1, x_lo + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 8192 + (k & 3) * 32), idesc, 1);
            }
            mma_commit_pair(stage_bar(wt));
          }
          __syncwarp();
        }
        if (kSplit == 3) {  // W1 lo: Xhi*W1lo
          const uint32_t wt = wait_stage();
          tc_fence_after();
          if (elect_one()) {
            if (!(a.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 8; k++) mma_ts_pair(d1, x_hi + k * 8, desc_kmajor_sw128(wt + (k >> 2) * 8192 + (k & 3) * 32), idesc, 1);
            }
            mma_commit_pair(stage_bar(wt));
          }
          __syncwarp();
        }
        if (elect_one()) {
          mma_commit_pair(&d1_full[g & 1]);
          if (c == ce - 1) mma_commit_pair(x_free);
        }
        __syncwarp();
        FFN_TRACE(0, 2, g);
        if (g >= 1) issue_g2(g - 1, prev_first, prev_last);
        prev_first = (c == cb), prev_last = (c == ce - 1);
      }
    }
    if (G > 0) issue_g2(G - 1, prev_first, prev_last);
    if (tr_on) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      a.trace[3 * 2048 + 2] = clock64(), a.trace[3 * 2048 + 3] = gt;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2 groups x 128 threads), per CTA
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_d1 = 0, ph_xfree = 0, ph_y = 0;
    uint8_t* xs = smem + FfnPairSmem::XS;
    const uint8_t* xs_row = xs + row * kXsRow;
    auto tile_row0 = [&](int64_t it) -> int64_t { return (tile_index(it) * 2 + rank) * 128; };

    uint32_t ph_xsfull = 0;
    auto x_to_tmem = [&](bool row_valid) {  // staged row -> bf16 hi/lo A-operand images in TMEM (rows beyond M are zeros)
      mbar_wait(xs_full, ph_xsfull);
      ph_xsfull ^= 1;
#pragma unroll 1
      for (int b = 0; b < 4; b++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float4 v = *reinterpret_cast<const float4*>(xs_row + b * 128 + j * 16);
          if (!row_valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
          split2(v.x, v.y, hi[2 * j], lo[2 * j]);
          split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        tmem_st16(tmem + lane_base + TM_X + b * 16, hi);
        if (kSplit == 3) tmem_st16(tmem + lane_base + TM_X + 64 + b * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cluster(x_full, 0);
    };
    // ---- tile boundary.  Every epilogue thread owns 64 columns (its group's half) of its row of Y: it drains them into
    // registers as soon as the last G2 of the previous tile has retired and immediately re-initialises the same columns
    // with x + b2 of the new tile (the residual and the second bias ride in the accumulator), so the accumulator is
    // handed back to the MMA warp after ~1 k cycles; the LayerNorm of the drained tile then runs out of registers, with
    // the row statistics combined across the two groups through shared memory.
    auto drain_y = [&](uint32_t (&ya)[32], uint32_t (&yb)[32]) {
      mbar_wait(y_full, ph_y);
      ph_y ^= 1;
      tc_fence_after();
      tmem_ld32(tmem + lane_base + TM_Y + hf * 64, ya);
      tmem_ld32(tmem + lane_base + TM_Y + hf * 64 + 32, yb);
      tmem_ld_wait();
    };
    auto init_y = [&](bool row_valid, bool with_residual) {  // with_residual false: a later part of a split tile starts from 0
#pragma unroll 1
      for (int b = 0; b < 4; b++) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4 v = *reinterpret_cast<const float4*>(xs_row + hf * 256 + b * 64 + j * 16);
          if (!row_valid || !with_residual) v = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 bv = *reinterpret_cast<const float4*>(vecs + hf * 64 + b * 16 + 4 * j);
          if (!with_residual) bv = make_float4(0.f, 0.f, 0.f, 0.f);
          r[4 * j] = __float_as_uint(v.x + bv.x), r[4 * j + 1] = __float_as_uint(v.y + bv.y);
          r[4 * j + 2] = __float_as_uint(v.z + bv.z), r[4 * j + 3] = __float_as_uint(v.w + bv.w);
        }
        tmem_st16(tmem + lane_base + TM_Y + hf * 64 + b * 16, r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cluster(y_init, 0);
    };
    float2* stat = reinterpret_cast<float2*>(smem + FfnPairSmem::STAT);
    uint8_t* park = xs + row * kXsRow + hf * 256;  // this thread's 64 columns of its row in the staging buffer
    auto layer_norm_half = [&](int64_t row0t, const uint32_t (&ya)[32], const uint32_t (&yb)[32]) {
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const float v0 = __uint_as_float(ya[j]), v1 = __uint_as_float(yb[j]);
        sum += v0 + v1;
        sq = fmaf(v0, v0, sq), sq = fmaf(v1, v1, sq);
      }
      stat[hf * 128 + row] = make_float2(sum, sq);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // both epilogue groups
      const float2 other = stat[(hf ^ 1) * 128 + row];
      sum += other.x, sq += other.y;
      const float mean = sum * (1.f / 128.f);
      const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
      const float rstd = 1.0f / sqrtf(var + a.eps);
      if (a.pre[net] && row0t + row < a.M) {  // training tape: the pre-LayerNorm sum (row-per-lane stores; not the MH path)
        float4* prow = reinterpret_cast<float4*>(a.pre[net] + (row0t + row) * 128 + hf * 64);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          prow[j] = make_float4(__uint_as_float(ya[4 * j]), __uint_as_float(ya[4 * j + 1]), __uint_as_float(ya[4 * j + 2]), __uint_as_float(ya[4 * j + 3]));
          prow[8 + j] = make_float4(__uint_as_float(yb[4 * j]), __uint_as_float(yb[4 * j + 1]), __uint_as_float(yb[4 * j + 2]), __uint_as_float(yb[4 * j + 3]));
        }
      }
      auto park32 = [&](const uint32_t (&r)[32], int off) {  // 32 normalised columns -> staging buffer
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 gm = *reinterpret_cast<const float4*>(vecs + 128 + hf * 64 + off + 4 * j);
          const float4 bt = *reinterpret_cast<const float4*>(vecs + 256 + hf * 64 + off + 4 * j);
          float4 o;
          o.x = (__uint_as_float(r[4 * j]) - mean) * rstd * gm.x + bt.x;
          o.y = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * gm.y + bt.y;
          o.z = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * gm.z + bt.z;
          o.w = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * gm.w + bt.w;
          *reinterpret_cast<float4*>(park + (off + 4 * j) * 4) = o;
        }
      };
      park32(ya, 0);
      park32(yb, 32);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk stores
    };

    if (hf == 1 && my_tiles > 0) x_to_tmem(tile_row0(0) + row < a.M);
    int64_t g = 0;
    for (int64_t it = 0; it < my_tiles; it++) {
      const bool has_next = it + 1 < my_tiles;
      const bool row_valid = tile_row0(it) + row < a.M, next_valid = has_next && tile_row0(it + 1) + row < a.M;
      const int cb = c_begin(it), ce = c_end(it);
      for (int c = cb; c < ce; c++, g++) {
        const int buf = (int)(g & 1);
        if (hf == 1 && ce - cb > 1 && c == ce - 1 && has_next) {
          // the last G1 of this tile and the release of the X images retire together: publish the next tile's images
          // FIRST, so that its first GEMM is queued behind G2 of the previous chunk without a bubble
          mbar_wait(x_free, ph_xfree);
          ph_xfree ^= 1;
          tc_fence_after();
          x_to_tmem(next_valid);
        }
        // ---- DH[buf][:, hf*64 .. +64): + b1, ReLU, hi/lo split, written back in place as the A operand of G2
        if (q == 0) { FFN_TRACE(1 + hf, 0, g); }
        mbar_wait(&d1_full[buf], (ph_d1 >> buf) & 1u);
        ph_d1 ^= 1u << buf;
        tc_fence_after();
        if (q == 0) { FFN_TRACE(1 + hf, 1, g); }
        if (!(a.dbg & 2)) {
          const float* bc = b1s + c * kFfnChunk + hf * 64;
          const uint32_t base = tmem + lane_base + TM_DH + buf * 128 + hf * 64;
          uint32_t r0[32], r1[32];
          tmem_ld32(base, r0);
          tmem_ld32(base + 32, r1);
          tmem_ld_wait();
          if (q == 0) { FFN_TRACE(1 + hf, 2, g); }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 bb = *reinterpret_cast<const float2*>(bc + j);
            split2(fmaxf(__uint_as_float(r0[j]) + bb.x, 0.f), fmaxf(__uint_as_float(r0[j + 1]) + bb.y, 0.f), hi[j >> 1], lo[j >> 1]);
          }
          tmem_st16(base, hi);
          if (kSplit == 3) tmem_st16(base + 32, lo);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 bb = *reinterpret_cast<const float2*>(bc + 32 + j);
            split2(fmaxf(__uint_as_float(r1[j]) + bb.x, 0.f), fmaxf(__uint_as_float(r1[j + 1]) + bb.y, 0.f), hi[j >> 1], lo[j >> 1]);
          }
          tmem_st16(base + 16, hi);
          if (kSplit == 3) tmem_st16(base + 48, lo);
          tmem_st_wait();
        }
        tc_fence_before();
        if (q == 0) { FFN_TRACE(1 + hf, 3, g); }
        mbar_arrive_cluster(&h_full[buf * 2 + hf], 0);
        if (q == 0) { FFN_TRACE(1 + hf, 4, g); }

        if (c == cb) {  // tile boundary: hand the accumulator back first, LayerNorm of the previous tile afterwards
          uint32_t ya[32], yb[32];
          if (it > 0) drain_y(ya, yb);
          if (hf == 0) {  // group 1 has already waited for this tile's staged rows in x_to_tmem
            mbar_wait(xs_full, ph_xsfull);
            ph_xsfull ^= 1;
          }
          init_y(row_valid, cb == 0);
          if (it > 0) layer_norm_half(tile_row0(it - 1), ya, yb);
          mbar_arrive(ln_staged);
        }
      }
      if (ce - cb == 1 && hf == 1 && has_next) {  // single-chunk iteration: publish the next tile's images here
        mbar_wait(x_free, ph_xfree);
        ph_xfree ^= 1;
        tc_fence_after();
        x_to_tmem(next_valid);
      }
    }
    if (my_tiles > 0) {
      uint32_t ya[32], yb[32];
      drain_y(ya, yb);
      const int64_t row0t = tile_row0(my_tiles - 1);
      if (!has_partial) {
        mbar_wait(xs_full, ph_xsfull);  // the staging buffer is free (the previous tile's rows have been stored)
        layer_norm_half(row0t, ya, yb);
      } else {
        // split tile: every part writes its accumulator to its own slot of the scratch tile (plain coalesced-per-row stores; the
        // first version added into one tile with float4 atomics: 8-15 us under 16-way contention); when all S parts are in, EVERY
        // participating pair sums, normalises and stores 1/S of the rows -- one warp per row, straight to global memory (the first
        // version left the whole tile to the last pair to arrive: another 11 us with 147 SMs idle).
        float* tile_base = a.tail[net] + (size_t)part_l * 16 * 2 * 128 * 128;
        float4* slot = reinterpret_cast<float4*>(tile_base + (((size_t)part_p * 2 + rank) * 128 + row) * 128 + hf * 64);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          __stcg(slot + j, make_float4(__uint_as_float(ya[4 * j]), __uint_as_float(ya[4 * j + 1]), __uint_as_float(ya[4 * j + 2]), __uint_as_float(ya[4 * j + 3])));
          __stcg(slot + 8 + j, make_float4(__uint_as_float(yb[4 * j]), __uint_as_float(yb[4 * j + 1]), __uint_as_float(yb[4 * j + 2]), __uint_as_float(yb[4 * j + 3])));
        }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        int* cnt = a.tail_cnt[net] + part_l * 4 + rank;
        if (warp == 2 && lane == 0) {
          atomicAdd(cnt, 1);
          // every participant is resident (persistent grid, one CTA per SM): spinning on the others cannot deadlock
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(cnt) < S) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        __threadfence();  // (acquire: the other parts' sums)
        const int rows_per = 128 / S;
        const float4 gm = *reinterpret_cast<const float4*>(vecs + 128 + 4 * lane), bt = *reinterpret_cast<const float4*>(vecs + 256 + 4 * lane);
        for (int rr = warp - 2; rr < rows_per; rr += 8) {  // one warp per row, a float4 of columns per lane
          const int r = part_p * rows_per + rr;
          if (row0t + r >= a.M) continue;  // (warp-uniform)
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int pp = 0; pp < S; pp++) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(tile_base + (((size_t)pp * 2 + rank) * 128 + r) * 128) + lane);
            y.x += v.x, y.y += v.y, y.z += v.z, y.w += v.w;
          }
          float sum = (y.x + y.y) + (y.z + y.w);
          float sq = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, y.w * y.w)));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          const float mean = sum * (1.f / 128.f);
          const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
          const float rstd = 1.0f / sqrtf(var + a.eps);
          if (a.pre[net]) *(reinterpret_cast<float4*>(a.pre[net] + (row0t + r) * 128) + lane) = y;  // training tape: the pre-LayerNorm sum
          float4 o4;
          o4.x = (y.x - mean) * rstd * gm.x + bt.x;
          o4.y = (y.y - mean) * rstd * gm.y + bt.y;
          o4.z = (y.z - mean) * rstd * gm.z + bt.z;
          o4.w = (y.w - mean) * rstd * gm.w + bt.w;
          *(reinterpret_cast<float4*>(a.out[net] + (row0t + r) * 128) + lane) = o4;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {  // the last pair to have read its rows resets both counters
          if (atomicAdd(cnt + 2, 1) == S - 1) atomicExch(cnt, 0), atomicExch(cnt + 2, 0);
        }
      }
      mbar_arrive(ln_staged);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / arrive on its barriers until here
  if (tr_on && warp == 1) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * 2048 + 6] = clock64(), a.trace[3 * 2048 + 7] = gt;
  }
  if (a.trace != nullptr && tid == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[3 * 2048 + 8 + 2 * (blockIdx.y * gridDim.x + blockIdx.x) + 1] = gt;
  }
  if (warp == 1) tmem_dealloc_pair<512>(tmem);
#undef FFN_TRACE
}

static long long* g_ffn_trace = nullptr;
static long long* g_inmlp_trace = nullptr;
static long long* g_mix_trace = nullptr;
static long long* g_attn_trace = nullptr;
void tc_set_ffn_trace(long long* buf) { g_ffn_trace = buf; }
void tc_set_trace(int cls, long long* buf) {
  if (cls == 1) g_ffn_trace = buf;
  if (cls == 2) g_mix_trace = buf;
  if (cls == 3) g_attn_trace = buf;
  if (cls == 4) tc_set_fm_trace(buf);
  if (cls == 5) g_inmlp_trace = buf;
}

static int launch_ffn_tc(const tw_flow_config* c, const FfnArgs& a_in, cudaStream_t st) {
  FfnArgs a = a_in;
  a.trace = g_ffn_trace;
  static DeviceOnce attr_done;
  static int use_pair = 1;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_ffn_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL + 1024));
    TW_CUDA(cudaFuncSetAttribute(k_ffn_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL + 1024));
    TW_CUDA(cudaFuncSetAttribute(k_ffn_pair<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnPairSmem::TOTAL + 1024));
    TW_CUDA(cudaFuncSetAttribute(k_ffn_pair<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnPairSmem::TOTAL + 1024));
    const char* e = getenv("TW_FFN_PAIR");  // bring-up switch: 0 = single-CTA kernel
    use_pair = e ? atoi(e) : 1;
    attr_done.mark();
  }
  TW_CHECK_ARG(a.F <= 4096, "dim_feedforward > 4096 not supported by the tensor-core FFN");
  const int sms = 148;
  const int64_t n_tiles = (a.M + 127) / 128;
  if (n_tiles < 1) return TW_OK;
  ProfScope prof(PROF_FFN, st);
  if (use_pair && n_tiles >= 2 && a.F <= kPairMaxF) {  // CTA pairs: 256-token tiles, grid.x = 2 * pairs per network
    const int64_t n_ptiles = (a.M + 255) / 256;
    const int pairs = (int)((n_ptiles < sms / 4) ? n_ptiles : sms / 4);
    dim3 grid(2 * pairs, 2);
    if (c->precision == TW_PRECISION_BF16X3)
      k_ffn_pair<3><<<grid, kPairThreads, FfnPairSmem::TOTAL + 1024, st>>>(a);
    else
      k_ffn_pair<1><<<grid, kPairThreads, FfnPairSmem::TOTAL + 1024, st>>>(a);
  } else {
    const int gx = (int)((n_tiles < sms / 2) ? n_tiles : sms / 2);
    dim3 grid(gx, 2);
    if (c->precision == TW_PRECISION_BF16X3)
      k_ffn_tc<3><<<grid, kFfnThreads, FfnSmem::TOTAL + 1024, st>>>(a);
    else
      k_ffn_tc<1><<<grid, kFfnThreads, FfnSmem::TOTAL + 1024, st>>>(a);
  }
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// ============================================================================================
// Attention, step 1 (scores operand images).  The attention weights depend only on the conditioning
// coordinates, so they are converted ONCE per pass into the B-operand image the mixing kernel reads:
// per (state b, head h) a [VP x VP] K-major, un-swizzled (8x8 core matrices) bf16 matrix, hi then lo.
//   element (i, j) at  (i/8)*(VP/8)*128 + (j/8)*128 + (i%8)*16 + (j%8)*2      (VP = V rounded up to 16)
__global__ void __launch_bounds__(256) k_scores_img(const float* __restrict__ scores, int64_t n_cond, int V, int VP, int H,
                                                    uint8_t* __restrict__ img, int transpose) {
  // one warp per (b, h, i) row, i < VP
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_cond * H * VP) return;
  const int lane = threadIdx.x & 31;
  const int i = (int)(row % VP);
  const int64_t bh = row / VP;
  const float* src = scores + (bh * V + i) * (int64_t)V;
  const size_t mat = (size_t)VP * VP * 2;
  uint8_t* hi = img + (size_t)bh * 2 * mat;
  uint8_t* lo = hi + mat;
  for (int j = lane; j < VP; j += 32) {
    float v = (i < V && j < V) ? (transpose ? scores[(bh * V + j) * (int64_t)V + i] : src[j]) : 0.f;  // transpose: A_h^T (backward)
    __nv_bfloat16 h = __float2bfloat16(v);
    __nv_bfloat16 l = __float2bfloat16(v - __bfloat162float(h));
    uint32_t off = (i >> 3) * ((VP >> 3) * 128u) + (j >> 3) * 128u + (i & 7) * 16u + (j & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
  }
}

// Scores straight to operand images (inference): the fp32 score tensor [n_cond, H, V, V] (104 MB at 1024 x 65 atoms) is
// only an intermediate of the tensor-core path, so one CTA per conditioning state computes
//   w_hij = exp(-(|x_i - x_j| / l_h)^2), key-masked, L1-row-normalised (+1e-5)      kernel_attention.py:98-119
// for every head and writes the bf16 hi/lo images of k_scores_img.  The distances of the state go to shared memory
// once; then one thread per (head, image row): row sum in a first sweep, and in a second sweep eight columns at a time
// -- exactly one 16-byte chunk of the swizzle-free [8 x 8] core-matrix layout -- so eight neighbouring lanes (rows) fill one
// 128-byte line with plain 16-byte stores and the padding rows / columns are written as zeros on the way.  (The first
// version ran one warp per row and (state, head) CTA through a shared staging image: 4 column slots per lane for 65 columns,
// IEEE divisions per element, 2-byte scatter stores -- issue-bound at 179 us per launch for 157 MB of output.)
__global__ void __launch_bounds__(1024) k_scores_direct_img(const float* __restrict__ xc, const uint8_t* __restrict__ mask,
                                                            const float* __restrict__ ls, int V, int VP, int H,
                                                            uint8_t* __restrict__ img, const float* __restrict__ cheb, int order,
                                                            int force_zero) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* xs = reinterpret_cast<float*>(sm_raw);                  // [V][3]
  uint8_t* ms = sm_raw + (((size_t)V * 12 + 15) & ~(size_t)15);  // [V]
  float* dist = reinterpret_cast<float*>(ms + (((size_t)V + 15) & ~(size_t)15));  // [V][VS], odd row stride
  const int VS = V | 1;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < V * 3; i += blockDim.x) xs[i] = xc[b * V * 3 + i];
  for (int i = tid; i < V; i += blockDim.x) ms[i] = mask[b * V + i];
  __syncthreads();
  for (int e = tid; e < V * V; e += blockDim.x) {
    const int i = e / V, j = e - i * V;
    const float dx = xs[i * 3] - xs[j * 3], dy = xs[i * 3 + 1] - xs[j * 3 + 1], dz = xs[i * 3 + 2] - xs[j * 3 + 2];
    dist[i * VS + j] = sqrtf(dx * dx + dy * dy + dz * dz);
  }
  __syncthreads();
  const uint32_t mat = (uint32_t)VP * VP * 2;
  const int nb = VP >> 3;
  for (int t = tid; t < H * VP; t += blockDim.x) {
    const int h = t / VP, i = t - h * VP;
    uint8_t* dst = img + ((size_t)b * H + h) * 2 * mat + (size_t)(i >> 3) * ((size_t)nb * 128) + (size_t)(i & 7) * 16;
    if (i >= V) {
      for (int jb = 0; jb < nb; jb++) {
        *reinterpret_cast<uint4*>(dst + jb * 128) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dst + mat + jb * 128) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const float inv_l = 1.0f / ls[h];
    const float* coef = cheb ? cheb + (size_t)h * order : nullptr;
    const float cmean = cheb_mean(coef, order, force_zero);
    const float* drow = dist + i * VS;
    float sum = 0.f;
    for (int j = 0; j < V; j++) {
      const float w = ms[j] ? 0.f : attention_basis(drow[j] * inv_l, coef, order, cmean);
      sum += fabsf(w);
    }
    const float inv = 1.0f / (sum + 1e-5f);
    for (int jb = 0; jb < nb; jb++) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int j = jb * 8 + u;
        v[u] = (j < V && !ms[j]) ? attention_basis(drow[j] * inv_l, coef, order, cmean) * inv : 0.f;
      }
      uint4 hi, lo;
      split2(v[0], v[1], hi.x, lo.x), split2(v[2], v[3], hi.y, lo.y), split2(v[4], v[5], hi.z, lo.z), split2(v[6], v[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(dst + jb * 128) = hi;
      *reinterpret_cast<uint4*>(dst + mat + jb * 128) = lo;
    }
  }
}

// ============================================================================================
// Attention, step 2 (mixing): for every sample n and head h,   mixed_h = A_h x   (kernel_attention.py:139
// re-associated with the value projection, see k_pack_all).  Features on the TMEM lanes:
//   D^T[128 f, VP tokens] = X^T[128 f, VP atoms] (A operand, TMEM) * A_h^T (B operand [VP i, VP j] K-major)
// so a sample of ANY atom count uses a full-width MMA.  The result is written straight into the A-operand
// images ([128 tokens x 64] K-major SW128 tiles, hi/lo) that the projection kernel bulk-copies.
constexpr uint32_t MX_HS = 0;      // X^T hi/lo, double buffered: [buf][hi 64 | lo 64] columns
constexpr uint32_t MX_D = 256;     // accumulators, double buffered: 2 x 128 columns

struct MixArgs {
  const float* x[2];         // [n*V, 128] layer input
  uint8_t* img[2];           // mixed operand images: [tile][H*2 K blocks][hi 16K | lo 16K]
  const uint8_t* scores_img;  // [n_cond][H][hi | lo][VP*VP*2]
  int64_t n, n_cond;
  int V, VP, H;
  int n_stages;              // ring depth (3, or 2 when VP > 96)
  long long* trace;          // optional event trace (debug)
};

__device__ __forceinline__ void split1(float v, uint16_t& hi, uint16_t& lo) {
  __nv_bfloat16 h = __float2bfloat16(v);
  __nv_bfloat16 l = __float2bfloat16(v - __bfloat162float(h));
  hi = *reinterpret_cast<uint16_t*>(&h);
  lo = *reinterpret_cast<uint16_t*>(&l);
}
__device__ __forceinline__ void group_bar_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// Warp roles: warp 0 streams the score images, warp 1 issues the MMAs, then `n_groups` (1 or 2) epilogue groups of
// four warps.  Group g drains accumulator buffer g (every other head): TMEM -> bf16 hi/lo rows in its staging
// buffer -> 16-byte chunks of the swizzled A-operand images in global memory.
template <int kSplit>
__global__ void __launch_bounds__(320, 1) k_mix_tc(MixArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V, VP = a.VP, H = a.H, n_stages = a.n_stages;
  const int n_groups = (blockDim.x - 64) / 128;
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  const uint32_t stage_bytes = 2 * mat_bytes;
  const uint32_t stage_stride = (stage_bytes + 1023) & ~1023u;
  uint8_t* ring = smem;
  uint8_t* staging0 = smem + n_stages * stage_stride;  // per group: [hi: VP x 256 B][lo: VP x 256 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging0 + (size_t)n_groups * VP * 512);
  uint64_t* full = bars;
  uint64_t* empty = bars + 3;
  uint64_t* hs_full = empty + 3;   // [2]
  uint64_t* hs_free = hs_full + 2;  // [2]
  uint64_t* d_full = hs_free + 2;   // [2]
  uint64_t* d_free = d_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_free + 2);

  if (tid == 0) {
    for (int i = 0; i < 3; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&hs_full[i], 128 * n_groups);
      mbar_init(&hs_free[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_free[i], 128);
    }
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
#define MIX_TRACE(role, ev, item) TW_TRACE(a.trace, tr_on, tr_n, role, ev, item)

  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x) {
      const uint8_t* src = a.scores_img + (size_t)(n % a.n_cond) * H * stage_bytes;
      for (int h = 0; h < H; h++) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], kSplit == 3 ? stage_bytes : mat_bytes);
          bulk_g2s(ring + stage * stage_stride, src + (size_t)h * stage_bytes, kSplit == 3 ? stage_bytes : mat_bytes, &full[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)n_stages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    uint32_t ph_hs[2] = {0, 0}, ph_dfree[2] = {0, 0};
    const uint32_t idesc = make_idesc_bf16(128, VP, 0, 0);
    const uint32_t lbo = 128, sbo = (uint32_t)(VP >> 3) * 128;
    int64_t it = 0, hcount = 0;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x, it++) {
      const int sb = it & 1;
      MIX_TRACE(0, 0, it);
      mbar_wait(&hs_full[sb], ph_hs[sb]);
      ph_hs[sb] ^= 1;
      tc_fence_after();
      MIX_TRACE(0, 1, it);
      const uint32_t hs_hi = tmem + MX_HS + sb * 128, hs_lo = hs_hi + 64;
      for (int h = 0; h < H; h++, hcount++) {
        const int db = hcount & 1;
        if (hcount >= 2) {
          mbar_wait(&d_free[db], ph_dfree[db]);
          ph_dfree[db] ^= 1;
        }
        MIX_TRACE(0, 2, hcount);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        MIX_TRACE(0, 3, hcount);
        if (elect_one()) {
          const uint32_t s_hi = smem_u32(ring + stage * stage_stride), s_lo = s_hi + mat_bytes;
          const uint32_t d = tmem + MX_D + db * 128;
          for (int k = 0; k < ksteps; k++)
            mma_ts(d, hs_hi + k * 8, make_smem_desc(s_hi + k * 256, lbo, sbo, LAYOUT_NONE), idesc, k > 0);
          if (kSplit == 3) {
            for (int k = 0; k < ksteps; k++)
              mma_ts(d, hs_lo + k * 8, make_smem_desc(s_hi + k * 256, lbo, sbo, LAYOUT_NONE), idesc, 1);
            for (int k = 0; k < ksteps; k++)
              mma_ts(d, hs_hi + k * 8, make_smem_desc(s_lo + k * 256, lbo, sbo, LAYOUT_NONE), idesc, 1);
          }
          mma_commit(&empty[stage]);
          mma_commit(&d_full[db]);
          if (h == H - 1) mma_commit(&hs_free[sb]);
        }
        __syncwarp();
        MIX_TRACE(0, 4, hcount);
        if (++stage == (uint32_t)n_stages) stage = 0, phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;  // epilogue group
    const int f = q * 32 + lane;    // feature = TMEM lane
    const int et = (tid - 64) & 127;  // index inside the group
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float* x = a.x[net];
    uint8_t* img = a.img[net];
    const size_t tile_bytes = (size_t)H * 2 * 2 * 16384;  // H*2 K blocks x (hi + lo) x 16 KB
    uint8_t* staging = staging0 + (size_t)g * VP * 512;
    uint8_t* st_hi = staging + f * 2;
    uint8_t* st_lo = staging + (size_t)VP * 256 + f * 2;
    uint32_t ph_hsfree[2] = {0, 0}, ph_dfull = 0;

    auto load_hs = [&](int64_t n, int64_t it) {  // column blocks of X^T are split between the groups
      const int sb = it & 1;
      if (q == 0) { MIX_TRACE(1 + g, 0, it); }
      if (it >= 2) {
        mbar_wait(&hs_free[sb], ph_hsfree[sb]);
        ph_hsfree[sb] ^= 1;
      }
      const float* xs = x + n * V * 128 + f;
      int blk = 0;
      for (int c0 = 0; c0 < VP / 2; c0 += 16, blk++) {  // 16 columns = 32 atoms per store
        if (n_groups == 2 && (blk & 1) != g) continue;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          int a0 = 2 * (c0 + j);
          float v0 = (a0 < V) ? __ldg(xs + (size_t)a0 * 128) : 0.f;
          float v1 = (a0 + 1 < V) ? __ldg(xs + (size_t)(a0 + 1) * 128) : 0.f;
          split2(v0, v1, hi[j], lo[j]);
        }
        tmem_st16(tmem + lane_base + MX_HS + sb * 128 + c0, hi);
        if (kSplit == 3) tmem_st16(tmem + lane_base + MX_HS + sb * 128 + 64 + c0, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&hs_full[sb]);
      if (q == 0) { MIX_TRACE(1 + g, 1, it); }
    };

    int64_t it = 0, hcount = 0;
    if ((int64_t)blockIdx.x < a.n) load_hs(blockIdx.x, 0);
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x, it++) {
      if (n + gridDim.x < a.n) load_hs(n + gridDim.x, it + 1);  // next sample's X^T while this one's heads drain
      const int64_t t0 = n * V;
      for (int h = 0; h < H; h++, hcount++) {
        const int db = hcount & 1;
        if (n_groups == 2 && db != g) continue;
        if (q == 0) { MIX_TRACE(1 + g, 2, hcount); }
        if (n_groups == 2) {
          mbar_wait(&d_full[db], ph_dfull);
          ph_dfull ^= 1;
        } else {
          mbar_wait(&d_full[db], (uint32_t)((hcount >> 1) & 1));
        }
        tc_fence_after();
        if (q == 0) { MIX_TRACE(1 + g, 3, hcount); }
        // (1) accumulator -> bf16 hi/lo rows in the staging buffer [token i][feature f]; rows >= V are scratch
        for (int c0 = 0; c0 < VP; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem + lane_base + MX_D + db * 128 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            uint32_t hi, lo;
            split2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), hi, lo);
            const int i = c0 + j;
            *reinterpret_cast<uint16_t*>(st_hi + i * 256) = (uint16_t)hi;
            *reinterpret_cast<uint16_t*>(st_hi + i * 256 + 256) = (uint16_t)(hi >> 16);
            if (kSplit == 3) {
              *reinterpret_cast<uint16_t*>(st_lo + i * 256) = (uint16_t)lo;
              *reinterpret_cast<uint16_t*>(st_lo + i * 256 + 256) = (uint16_t)(lo >> 16);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&d_free[db]);
        if (q == 0) { MIX_TRACE(1 + g, 4, hcount); }
        group_bar_sync(g);
        if (q == 0) { MIX_TRACE(1 + g, 5, hcount); }
        // (2) 16-byte chunks -> swizzled operand images in global memory
        const int n_chunks = V * (kSplit == 3 ? 32 : 16);
        for (int cid = et; cid < n_chunks; cid += 128) {
          const int i = (kSplit == 3) ? (cid >> 5) : (cid >> 4);
          const int hl = (kSplit == 3) ? ((cid >> 4) & 1) : 0;
          const int c = cid & 15;
          const uint4 v = *reinterpret_cast<const uint4*>(staging + (size_t)hl * VP * 256 + i * 256 + c * 16);
          const int64_t t = t0 + i;
          const uint32_t row = (uint32_t)(t & 127);
          uint8_t* dst = img + (t >> 7) * tile_bytes + (size_t)(h * 2 + (c >> 3)) * 32768 + hl * 16384 + row * 128u +
                         ((((uint32_t)c & 7u) ^ (row & 7u)) << 4);
          *reinterpret_cast<uint4*>(dst) = v;
        }
        if (q == 0) { MIX_TRACE(1 + g, 6, hcount); }
        group_bar_sync(g);
        if (q == 0) { MIX_TRACE(1 + g, 7, hcount); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef MIX_TRACE
}

// ============================================================================================
// Attention, step 2, token-major form (atom counts up to 80):  mixed_h = A_h x  with the TOKENS of one sample on the
// TMEM lanes,
//   D[128 i, 128 f] = A_h[i, j] (A operand: the [VP x VP] K-major score image, rows >= VP read scratch and are never
//                                 stored) * X[j, f] (B operand: the sample's x as bf16 hi/lo, MN-major SW128 tiles)
// so the accumulator row of a thread IS a row of the A-operand image the projection kernel reads: the epilogue only
// converts to bf16 hi/lo and writes swizzled 16-byte chunks (no transposition through 2-byte shared-memory stores,
// which bound the feature-major kernel), stages whole image rows and hands them to bulk stores.
//   warp 0: score images (bulk loads)          warp 1: MMA issuer          warps 2-3: x -> bf16 hi/lo operand tiles
//   warps 4-7 / 8-11: two epilogue groups draining alternate heads (double-buffered accumulator)
constexpr int kMixTokThreads = 384;
constexpr int kMixTokMaxVP = 80;

constexpr int kMixTokStages = 3;  // score-image ring depth (2 when three stages do not fit next to the other buffers)
struct MixTokSmem {
  uint32_t stage_bytes, xb_bytes, out_bytes;
  int n_stages;
  __host__ __device__ MixTokSmem(int v, int vp, int stages) : n_stages(stages) {
    stage_bytes = (uint32_t)(2 * vp * vp * 2);  // score image of one head: hi | lo (un-swizzled: 16-byte alignment is enough)
    xb_bytes = (uint32_t)(4 * vp * 128);        // [hi: 2 N blocks x VP rows x 128 B][lo: same], 1024-byte aligned (128-byte swizzle)
    out_bytes = (uint32_t)(4 * v * 128);        // [kb0 hi][kb0 lo][kb1 hi][kb1 lo], V rows x 128 B each
    overread = (uint32_t)(2 * vp * (128 - vp));
  }
  // The MMA reads 128 rows of every score image (A operand, M = 128) although only VP exist: rows >= VP fall up to
  // 2*VP*(128-VP) bytes behind the ring, so the staging buffers (never read by an MMA) and, for tiny samples, an
  // explicit pad sit behind it.
  uint32_t overread;
  __host__ __device__ uint32_t xb(int b) const { return b * xb_bytes; }
  __host__ __device__ uint32_t ring() const { return 2 * xb_bytes; }
  __host__ __device__ uint32_t out(int g) const { return (ring() + n_stages * stage_bytes + 15 & ~15u) + g * out_bytes; }
  __host__ __device__ uint32_t bars() const {
    const uint32_t behind = 2 * out_bytes;
    return (out(0) + (behind > overread ? behind : overread) + 15) & ~15u;
  }
  __host__ __device__ uint32_t total() const { return bars() + 256; }
};

// kMinBlocks = 2: small samples (shared memory below half an SM's) run two CTAs per SM -- the kernel allocates 256 TMEM columns --
// so that a training batch of a few hundred samples needs half as many rounds (register budget 80 per thread).
template <int kSplit, int kMinBlocks = 1>
__global__ void __launch_bounds__(kMixTokThreads, kMinBlocks) k_mix_tok(MixArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V, VP = a.VP, H = a.H;
  const int n_stages = a.n_stages;
  const MixTokSmem L(V, VP, n_stages);
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  const uint32_t stage_bytes = 2 * mat_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars());
  uint64_t* full = bars;                      // [kMixTokStages] score image of a head landed
  uint64_t* empty = full + kMixTokStages;     // [kMixTokStages]
  uint64_t* xb_full = empty + kMixTokStages;  // [2] x operand tiles of a sample written (64 arrivals)
  uint64_t* xb_free = xb_full + 2; // [2] every MMA of that sample has retired
  uint64_t* d_full = xb_free + 2;  // [2]
  uint64_t* d_free = d_full + 2;   // [2] 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_free + 2);

  if (tid == 0) {
    for (int i = 0; i < kMixTokStages; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&xb_full[i], 64), mbar_init(&xb_free[i], 1);
      mbar_init(&d_full[i], 1), mbar_init(&d_free[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ksteps = VP / 16;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define MIX_TRACE(role, ev, item) TW_TRACE(a.trace, tr_on, tr_n, role, ev, item)

  if (warp == 0) {
    // ------------------------------------------------------------------ score images
    uint32_t stage = 0, phase = 0;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x) {
      const uint8_t* src = a.scores_img + (size_t)(n % a.n_cond) * H * stage_bytes;
      for (int h = 0; h < H; h++) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], kSplit == 3 ? stage_bytes : mat_bytes);
          bulk_g2s(smem + L.ring() + stage * L.stage_bytes, src + (size_t)h * stage_bytes, kSplit == 3 ? stage_bytes : mat_bytes, &full[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)n_stages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t stage = 0, phase = 0, ph_xb = 0 /* bit per buffer */, ph_dfree = 0;
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 1);  // A K-major, B MN-major
    const uint32_t a_lbo = 128, a_sbo = (uint32_t)(VP >> 3) * 128;
    const uint32_t b_lbo = (uint32_t)VP * 128;  // stride between the two 64-feature N blocks
    const uint64_t a_desc0 = make_smem_desc(0, a_lbo, a_sbo, LAYOUT_NONE), b_desc0 = make_smem_desc(0, b_lbo, 1024, LAYOUT_SW128);
    int64_t it = 0, hcount = 0;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x, it++) {
      const int sb = (int)(it & 1);
      MIX_TRACE(0, 0, it);
      mbar_wait(&xb_full[sb], (ph_xb >> sb) & 1u);
      ph_xb ^= 1u << sb;
      tc_fence_after();
      MIX_TRACE(0, 1, it);
      const uint32_t x_hi = smem_u32(smem + L.xb(sb)), x_lo = x_hi + L.xb_bytes / 2;
      const uint64_t b_hi = b_desc0 | (uint64_t)((x_hi >> 4) & 0x3FFF), b_lo = b_desc0 | (uint64_t)((x_lo >> 4) & 0x3FFF);
      for (int h = 0; h < H; h++, hcount++) {
        const int db = (int)(hcount & 1);
        if (hcount >= 2) {
          mbar_wait(&d_free[db], (ph_dfree >> db) & 1u);
          ph_dfree ^= 1u << db;
        }
        MIX_TRACE(0, 2, hcount);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        MIX_TRACE(0, 3, hcount);
        if (elect_one()) {
          // descriptors differ only in the 14-bit start-address field (16-byte units): one 64-bit add per k step
          const uint32_t s_hi = smem_u32(smem + L.ring() + stage * L.stage_bytes), s_lo = s_hi + mat_bytes;
          const uint64_t a_hi = a_desc0 | (uint64_t)((s_hi >> 4) & 0x3FFF), a_lo = a_desc0 | (uint64_t)((s_lo >> 4) & 0x3FFF);  // (mask: inside a cluster the shared-window address carries the CTA rank)
          const uint32_t d = tmem + db * 128;
#pragma unroll
          for (int k = 0; k < kMixTokMaxVP / 16; k++)
            if (k < ksteps) mma_ss(d, a_hi + (uint64_t)(k * 16), b_hi + (uint64_t)(k * 128), idesc, k > 0);
          if (kSplit == 3) {
#pragma unroll
            for (int k = 0; k < kMixTokMaxVP / 16; k++)
              if (k < ksteps) mma_ss(d, a_lo + (uint64_t)(k * 16), b_hi + (uint64_t)(k * 128), idesc, 1);
#pragma unroll
            for (int k = 0; k < kMixTokMaxVP / 16; k++)
              if (k < ksteps) mma_ss(d, a_hi + (uint64_t)(k * 16), b_lo + (uint64_t)(k * 128), idesc, 1);
          }
          mma_commit(&empty[stage]);
          mma_commit(&d_full[db]);
          if (h == H - 1) mma_commit(&xb_free[sb]);
        }
        __syncwarp();
        MIX_TRACE(0, 4, hcount);
        if (++stage == (uint32_t)n_stages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ x of a sample -> bf16 hi/lo B-operand tiles
    // [atom j][64 features] rows of 128 B, 128-byte swizzle: the bytes of a K-major tile, read MN-major (N = feature).
    const int cw = warp - 2;  // rows j = cw, cw + 2, ...
    const float* x = a.x[net];
    uint32_t ph_free = 0;
    int64_t it = 0;
    const int nb = lane >> 4;                 // N block of the 4 features this lane converts
    const uint32_t c16 = (uint32_t)(lane & 15) >> 1, sub = (uint32_t)(lane & 1) * 8;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x, it++) {
      const int sb = (int)(it & 1);
      if (cw == 0) { MIX_TRACE(2, 0, it); }
      if (it >= 2) {
        mbar_wait(&xb_free[sb], (ph_free >> sb) & 1u);
        ph_free ^= 1u << sb;
      }
      if (cw == 0) { MIX_TRACE(2, 2, it); }
      uint8_t* hi_base = smem + L.xb(sb) + nb * (VP * 128);
      uint8_t* lo_base = hi_base + L.xb_bytes / 2;
      const float4* src = reinterpret_cast<const float4*>(x + n * V * 128) + lane;
      for (int j0 = cw; j0 < VP; j0 += 16) {  // 8 rows per warp in flight
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int j = j0 + 2 * u;
          v[u] = (j < V) ? __ldg(src + (size_t)j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int j = j0 + 2 * u;
          if (j < VP) {
            uint32_t h0, l0, h1, l1;
            split2(v[u].x, v[u].y, h0, l0);
            split2(v[u].z, v[u].w, h1, l1);
            const uint32_t off = (uint32_t)j * 128u + ((c16 ^ ((uint32_t)j & 7u)) << 4) + sub;
            *reinterpret_cast<uint2*>(hi_base + off) = make_uint2(h0, h1);
            if (kSplit == 3) *reinterpret_cast<uint2*>(lo_base + off) = make_uint2(l0, l1);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&xb_full[sb]);
      if (cw == 0) { MIX_TRACE(2, 1, it); }
    }
  } else {
    // ------------------------------------------------------------------ epilogue groups
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;
    const int i = q * 32 + lane;  // token (atom) of the sample = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool warp_active = q * 32 < V;
    const bool store_lane = q == 0 && lane < 4 && (kSplit == 3 || (lane & 1) == 0);  // one bulk-store issuer per region
    uint8_t* img = a.img[net];
    const size_t tile_bytes = (size_t)H * 2 * 2 * 16384;
    uint8_t* stg = smem + L.out(g);
    const uint32_t region = (uint32_t)V * 128;  // one (K block, hi/lo) region of the staging buffer
    uint32_t ph_dfull = 0;
    int64_t hcount = 0;
    bool stores_pending = false;
    for (int64_t n = blockIdx.x; n < a.n; n += gridDim.x) {
      const int64_t t0 = n * V;
      const uint32_t sw = (uint32_t)((t0 + i) & 7);  // swizzle phase of this token's row in the global image
      for (int h = 0; h < H; h++, hcount++) {
        const int db = (int)(hcount & 1);
        if (db != g) continue;
        if (q == 0 && g == 0) { MIX_TRACE(1, 2, hcount); }
        mbar_wait(&d_full[db], ph_dfull);
        ph_dfull ^= 1;
        tc_fence_after();
        if (q == 0 && g == 0) { MIX_TRACE(1, 3, hcount); }
        if (store_lane && stores_pending) bulk_wait_group_read0();  // the previous head's rows have left the staging buffer
        group_bar_sync(g);
        if (q == 0 && g == 0) { MIX_TRACE(1, 5, hcount); }
        if (warp_active) {
#pragma unroll 1
          for (int g4 = 0; g4 < 4; g4++) {
            uint32_t r[32];
            tmem_ld32(tmem + lane_base + db * 128 + g4 * 32, r);
            tmem_ld_wait();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; j++) split2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), hi[j], lo[j]);
            if (i < V) {
              uint8_t* row_hi = stg + (g4 >> 1) * 2 * region + i * 128;
              uint8_t* row_lo = row_hi + region;
#pragma unroll
              for (int cc = 0; cc < 4; cc++) {
                const uint32_t pos = ((((uint32_t)(g4 & 1) * 4 + cc) ^ sw) << 4);
                *reinterpret_cast<uint4*>(row_hi + pos) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
                if (kSplit == 3) *reinterpret_cast<uint4*>(row_lo + pos) = make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&d_free[db]);
        fence_proxy_async_smem();
        if (q == 0 && g == 0) { MIX_TRACE(1, 4, hcount); }
        group_bar_sync(g);
        if (q == 0 && g == 0) { MIX_TRACE(1, 7, hcount); }
        if (store_lane) {  // whole image rows -> global: lane = (K block, hi/lo) region; two pieces where the sample straddles a tile
          const int64_t tile = t0 >> 7;
          const int r0 = (int)(t0 & 127);
          const int n1 = (V < 128 - r0) ? V : 128 - r0;
          const int kb = lane >> 1, hl = lane & 1;
          const uint8_t* src = stg + (kb * 2 + hl) * region;
          uint8_t* dst = img + tile * tile_bytes + (size_t)(h * 2 + kb) * 32768 + hl * 16384;
          bulk_s2g(dst + r0 * 128, src, (uint32_t)n1 * 128u);
          if (n1 < V) bulk_s2g(dst + tile_bytes, src + n1 * 128, (uint32_t)(V - n1) * 128u);
          bulk_commit_group();
          stores_pending = true;
        }
        if (q == 0 && g == 0) { MIX_TRACE(1, 6, hcount); }
      }
    }
    if (store_lane) bulk_wait_group0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
#undef MIX_TRACE
}

// ============================================================================================
// Attention, step 3 (projection):  out = LayerNorm1(x + sum_h W_c,h mixed_h)     custom_attention_encoder.py:102-110
// Token-major GEMM, K = H*128: both operands arrive by bulk copy (A = mixed images, B = W_c images).
constexpr int kProjStages = 3;
constexpr int kProjStageBytes = 4 * 16384;  // A hi | A lo | W hi | W lo, each [128 x 64]

struct ProjArgs {
  const uint8_t* a_img[2];
  const uint8_t* w[2];
  const float* resid[2];
  float* out[2];
  float* pre[2];  // optional [M,128]: pre-LayerNorm sum (training)
  const float* gamma[2];
  const float* beta[2];
  int64_t M;
  int KB;  // K blocks of 64
  float eps;
};

template <int kSplit>
__global__ void __launch_bounds__(192, 1) k_proj_tc(ProjArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.M + 127) / 128;
  uint8_t* ring = smem;
  float* vecs = reinterpret_cast<float*>(smem + kProjStages * kProjStageBytes);  // gamma, beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(vecs + 256);
  uint64_t* full = bars;
  uint64_t* empty = bars + kProjStages;
  uint64_t* y_full = empty + kProjStages;  // [2]
  uint64_t* y_free = y_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_free + 2);
  if (tid == 0) {
    for (int i = 0; i < kProjStages; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; i++) mbar_init(&y_full[i], 1), mbar_init(&y_free[i], 128);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  for (int i = tid; i < 128; i += blockDim.x) vecs[i] = a.gamma[net][i], vecs[128 + i] = a.beta[net][i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const size_t a_tile_bytes = (size_t)a.KB * 32768;

  if (warp == 0) {
    {
      uint32_t stage = 0, phase = 0;
      const uint32_t a_bytes = kSplit == 3 ? 32768 : 16384;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < a.KB; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], 2 * a_bytes);
            uint8_t* dst = ring + stage * kProjStageBytes;
            bulk_g2s(dst, a.a_img[net] + tile * a_tile_bytes + (size_t)kb * 32768, a_bytes, &full[stage]);
            bulk_g2s(dst + 32768, a.w[net] + (size_t)kb * 32768, a_bytes, &full[stage]);
          }
          __syncwarp();
          if (++stage == kProjStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      uint32_t stage = 0, phase = 0, ph_free[2] = {0, 0};
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const int tb = it & 1;
        if (it >= 2) {
          mbar_wait(&y_free[tb], ph_free[tb]);
          ph_free[tb] ^= 1;
        }
        tc_fence_after();
        const uint32_t d = tmem + tb * 128;
        for (int kb = 0; kb < a.KB; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t ahi = smem_u32(ring + stage * kProjStageBytes), alo = ahi + 16384, whi = ahi + 32768, wlo = whi + 16384;
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(ahi + k * 32), desc_kmajor_sw128(whi + k * 32), idesc, (kb > 0 || k > 0) ? 1 : 0);
            if (kSplit == 3) {
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(alo + k * 32), desc_kmajor_sw128(whi + k * 32), idesc, 1);
#pragma unroll
              for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(ahi + k * 32), desc_kmajor_sw128(wlo + k * 32), idesc, 1);
            }
            mma_commit(&empty[stage]);
            if (kb == a.KB - 1) mma_commit(&y_full[tb]);
          }
          __syncwarp();
          if (++stage == kProjStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_full[2] = {0, 0};
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int tb = it & 1;
      mbar_wait(&y_full[tb], ph_full[tb]);
      ph_full[tb] ^= 1;
      tc_fence_after();
      float v[128];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + tb * 128 + g * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) v[g * 32 + j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
      mbar_arrive(&y_free[tb]);
      const int64_t grow = tile * 128 + row;
      if (grow < a.M)
        ln_store_row(v, nullptr, a.resid[net] + grow * 128, vecs, vecs + 128, a.eps, a.out[net] + grow * 128,
                     a.pre[net] ? a.pre[net] + grow * 128 : nullptr);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ============================================================================================
// Fused attention layer (inference):  out = LayerNorm1(x + sum_h (A_h x) W_c,h^T)   for ONE sample per CTA iteration.
// The two-kernel form (k_mix_tok -> 3 KB/token of operand images in HBM -> k_proj_tc) is bound by that round trip
// (409 MB written + 409 MB read per layer).  Here the per-head neighbourhood averages never leave the SM:
//   MMA1(h): DM[h&1][128 i, 128 f] = A_h (A operand: score image, smem) * X (B operand: the sample's x, bf16 hi/lo, MN-major)
//   epilogue: DM -> bf16 hi/lo IN PLACE (TMEM) -- the A operand of
//   MMA2(h): DO[n&1][128 i, 128 out] += mixed_h * W_c,h^T (B operand: [128 out x 128 K] K-major, streamed per sample)
// with DO pre-initialised with x (the residual rides in the accumulator), exactly the chunk pipeline of the fused FFN with
// heads as chunks (issue order MMA1(g), MMA2(g-1) over a global head counter).  Only V of the 128 token rows are real
// (51 % of the MMA rows for 65 atoms): the price of sample-aligned tiles; the kernel is tensor-bound, not HBM-bound.
//   warp 0 score-image + W_c producer   warp 1 MMA issuer   warp 2 activation I/O (x in, LayerNorm rows out, bulk copies)
//   warps 3-6 / 7-10 epilogue groups (feature halves [0,64) / [64,128)): in-place conversion of the mixed operand, sample
//   hand-over (accumulator initialisation + x -> bf16 hi/lo operand tiles), LayerNorm
// TMEM: DM0 | DM1 | DO0 | DO1, 128 columns each.
constexpr int kAttnThreads = 352;
constexpr int kAttnWcStages = 6;   // ring units of 16 KB: the hi or the lo image of one [128 out x 64 K] block of W_c,h
constexpr int kAttnWcStage = 16384;
constexpr int kAttnScStages = 2;
constexpr uint32_t AT_DM = 0, AT_DO = 256;

struct AttnArgs {
  const float* x[2];
  float* out[2];
  const uint8_t* scores_img;
  const uint8_t* wc[2];
  const float* gamma[2];
  const float* beta[2];
  int64_t n, n_cond;
  int V, VP, H;
  float eps;
  long long* trace;
};

struct AttnSmem {
  uint32_t xb_bytes, sc_stage, xs_bytes, pad;
  __host__ __device__ AttnSmem(int v, int vp) {
    xb_bytes = (uint32_t)(4 * vp * 128);
    sc_stage = (uint32_t)(2 * vp * vp * 2);
    xs_bytes = (uint32_t)(v * kXsRow);
    const uint32_t overread = (uint32_t)(2 * vp * (128 - vp));  // MMA1 reads 128 rows of a VP-row score image
    pad = xs_bytes >= overread ? 0u : ((overread - xs_bytes + 15) & ~15u);
  }
  __host__ __device__ uint32_t xb() const { return 0; }
  __host__ __device__ uint32_t wc() const { return xb_bytes; }
  __host__ __device__ uint32_t sc() const { return wc() + kAttnWcStages * kAttnWcStage; }
  __host__ __device__ uint32_t xs() const { return sc() + kAttnScStages * sc_stage; }
  __host__ __device__ uint32_t stat() const { return xs() + xs_bytes + pad; }
  __host__ __device__ uint32_t vec() const { return stat() + 2 * 128 * 8; }
  __host__ __device__ uint32_t bars() const { return vec() + 2 * 128 * 4; }
  __host__ __device__ uint32_t total() const { return bars() + 256; }
};

template <int kSplit>
__global__ void __launch_bounds__(kAttnThreads, 1) k_attn_fused(AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.V, VP = a.VP, H = a.H;
  const AttnSmem L(V, VP);
  const uint32_t mat_bytes = (uint32_t)VP * VP * 2;
  const int64_t my_n = ((int64_t)blockIdx.x < a.n) ? (a.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t G = my_n * H;
  constexpr int kParts = kSplit == 3 ? 2 : 1;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars());
  uint64_t* sc_full = bars;                       // [2]
  uint64_t* sc_empty = sc_full + kAttnScStages;   // [2]
  uint64_t* wc_full = sc_empty + kAttnScStages;   // [3]
  uint64_t* wc_empty = wc_full + kAttnWcStages;   // [3]
  uint64_t* dm_full = wc_empty + kAttnWcStages;   // [2] MMA1 into DM[b] retired
  uint64_t* h_full = dm_full + 2;                 // [4] (b, K half) of the mixed operand written in place, 128 arrivals
  uint64_t* xs_full = h_full + 4;                 // the sample's x rows landed in the staging buffer (tx bytes)
  uint64_t* xs_used = xs_full + 1;                // 256 arrivals: rows converted to operand tiles and used to initialise DO
  uint64_t* xb_full = xs_used + 1;                // 256 arrivals: operand tiles written
  uint64_t* xb_free = xb_full + 1;                // (unused: dm_full of a sample's last head implies it)
  uint64_t* do_init = xb_free + 1;                // 256 arrivals: DO initialised with x
  uint64_t* do_full = do_init + 1;                // commit: last MMA2 of the sample retired
  uint64_t* ln_staged = do_full + 1;              // 256 arrivals: LayerNorm rows parked in the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_staged + 1);
  float* vecs = reinterpret_cast<float*>(smem + L.vec());

  if (tid == 0) {
    for (int i = 0; i < kAttnScStages; i++) mbar_init(&sc_full[i], 1), mbar_init(&sc_empty[i], 1);
    for (int i = 0; i < kAttnWcStages; i++) mbar_init(&wc_full[i], 1), mbar_init(&wc_empty[i], 1);
    for (int i = 0; i < 2; i++) mbar_init(&dm_full[i], 1);
    for (int i = 0; i < 4; i++) mbar_init(&h_full[i], 128);
    mbar_init(xs_full, 1);
    mbar_init(xs_used, 256);
    mbar_init(xb_full, 256);
    mbar_init(xb_free, 1);
    mbar_init(do_init, 256);
    mbar_init(do_full, 1);
    mbar_init(ln_staged, 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < 128; i += blockDim.x) vecs[i] = a.gamma[net][i], vecs[128 + i] = a.beta[net][i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int ksteps = VP / 16;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define AT_TRACE(role, ev, item) TW_TRACE(a.trace, tr_on, tr_n, role, ev, item)
  auto sample_of = [&](int64_t it) -> int64_t { return blockIdx.x + it * gridDim.x; };

  if (warp == 0) {
    // ------------------------------------------------------------------ score images (lane 1) and W_c tiles (lane 0)
    // Two INDEPENDENT producer loops on two lanes of one warp: a single in-order producer would hold back W_c loads
    // whose ring stage is already free while it waits for a score stage (and vice versa).
    if (lane == 1) {
      uint32_t ss = 0, sp = 0;
      for (int64_t it = 0; it < my_n; it++) {
        const int64_t n = sample_of(it);
        const uint8_t* src0 = a.scores_img + (size_t)(a.n_cond == a.n ? n : n % a.n_cond) * H * (2 * mat_bytes);
        for (int h = 0; h < H; h++) {
          mbar_wait(&sc_empty[ss], sp ^ 1);
          mbar_arrive_expect_tx(&sc_full[ss], kParts * mat_bytes);
          bulk_g2s(smem + L.sc() + ss * L.sc_stage, src0 + (size_t)h * (2 * mat_bytes), kParts * mat_bytes, &sc_full[ss]);
          if (++ss == kAttnScStages) ss = 0, sp ^= 1;
        }
      }
    } else if (lane == 0) {
      uint32_t ws = 0, wp = 0;
      for (int64_t g = 0, h = 0; g < G; g++, h = (h + 1 == H ? 0 : h + 1)) {  // W_c,h: per K block the hi image, then the lo image
        for (int u = 0; u < 2 * kParts; u++) {
          const int kb = u / kParts, part = u % kParts;
          mbar_wait(&wc_empty[ws], wp ^ 1);
          mbar_arrive_expect_tx(&wc_full[ws], (uint32_t)kAttnWcStage);
          bulk_g2s(smem + L.wc() + ws * kAttnWcStage, a.wc[net] + (size_t)(h * 2 + kb) * 32768 + part * 16384, kAttnWcStage, &wc_full[ws]);
          if (++ws == kAttnWcStages) ws = 0, wp ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t ss = 0, sp = 0, ws = 0, wp = 0, ph_xb = 0, ph_h = 0, ph_doinit = 0;
    const uint32_t idesc1 = make_idesc_bf16(128, 128, 0, 1);  // A K-major (scores), B MN-major (x)
    const uint32_t idesc2 = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t a_desc0 = make_smem_desc(0, 128, (uint32_t)(VP >> 3) * 128, LAYOUT_NONE);
    const uint64_t b_desc0 = make_smem_desc(0, (uint32_t)VP * 128, 1024, LAYOUT_SW128);
    const uint32_t x_hi = smem_u32(smem + L.xb()), x_lo = x_hi + L.xb_bytes / 2;
    const uint64_t b_hi = b_desc0 | (uint64_t)((x_hi >> 4) & 0x3FFF), b_lo = b_desc0 | (uint64_t)((x_lo >> 4) & 0x3FFF);

    auto issue_mma2 = [&](int64_t gp, int hp, int dob) {  // DO[dob] += mixed(gp) W_c,hp^T
      const int b = (int)(gp & 1);
      if (hp == 0) {
        mbar_wait(do_init, ph_doinit);
        ph_doinit ^= 1;
      }
      for (int kb = 0; kb < 2; kb++) {
        mbar_wait(&wc_full[ws], wp);  // hi image of this K block
        AT_TRACE(0, 3 + 2 * kb, gp);
        mbar_wait(&h_full[b * 2 + kb], (ph_h >> (b * 2 + kb)) & 1u);
        ph_h ^= 1u << (b * 2 + kb);
        tc_fence_after();
        AT_TRACE(0, 4 + 2 * kb, gp);
        const uint32_t m_hi = tmem + AT_DM + b * 128 + kb * 64, m_lo = m_hi + 32;
        const uint32_t d = tmem + AT_DO + dob * 128;
        if (elect_one()) {
          const uint32_t w_hi = smem_u32(smem + L.wc() + ws * kAttnWcStage);
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ts(d, m_hi + k * 8, desc_kmajor_sw128(w_hi + k * 32), idesc2, 1);
          if (kSplit == 3) {
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts(d, m_lo + k * 8, desc_kmajor_sw128(w_hi + k * 32), idesc2, 1);
          }
          mma_commit(&wc_empty[ws]);
          if (kSplit != 3 && kb == 1 && hp == H - 1) mma_commit(do_full);
        }
        __syncwarp();
        if (++ws == kAttnWcStages) ws = 0, wp ^= 1;
        if (kSplit == 3) {
          mbar_wait(&wc_full[ws], wp);  // lo image
          tc_fence_after();
          if (elect_one()) {
            const uint32_t w_lo = smem_u32(smem + L.wc() + ws * kAttnWcStage);
#pragma unroll
            for (int k = 0; k < 4; k++) mma_ts(d, m_hi + k * 8, desc_kmajor_sw128(w_lo + k * 32), idesc2, 1);
            mma_commit(&wc_empty[ws]);
            if (kb == 1 && hp == H - 1) mma_commit(do_full);
          }
          __syncwarp();
          if (++ws == kAttnWcStages) ws = 0, wp ^= 1;
        }
      }
    };

    int64_t g = 0;
    for (int64_t it = 0; it < my_n; it++) {
      AT_TRACE(0, 7, it);
      mbar_wait(xb_full, ph_xb);
      ph_xb ^= 1;
      tc_fence_after();
      AT_TRACE(0, 8, it);
      for (int h = 0; h < H; h++, g++) {
        AT_TRACE(0, 0, g);
        mbar_wait(&sc_full[ss], sp);
        tc_fence_after();
        AT_TRACE(0, 1, g);
        if (elect_one()) {
          const uint32_t s_hi = smem_u32(smem + L.sc() + ss * L.sc_stage), s_lo = s_hi + mat_bytes;
          const uint64_t a_hi = a_desc0 | (uint64_t)((s_hi >> 4) & 0x3FFF), a_lo = a_desc0 | (uint64_t)((s_lo >> 4) & 0x3FFF);  // (mask: inside a cluster the shared-window address carries the CTA rank)
          const uint32_t d = tmem + AT_DM + (uint32_t)(g & 1) * 128;
#pragma unroll
          for (int k = 0; k < kMixTokMaxVP / 16; k++)
            if (k < ksteps) mma_ss(d, a_hi + (uint64_t)(k * 16), b_hi + (uint64_t)(k * 128), idesc1, k > 0);
          if (kSplit == 3) {
#pragma unroll
            for (int k = 0; k < kMixTokMaxVP / 16; k++)
              if (k < ksteps) mma_ss(d, a_lo + (uint64_t)(k * 16), b_hi + (uint64_t)(k * 128), idesc1, 1);
#pragma unroll
            for (int k = 0; k < kMixTokMaxVP / 16; k++)
              if (k < ksteps) mma_ss(d, a_hi + (uint64_t)(k * 16), b_lo + (uint64_t)(k * 128), idesc1, 1);
          }
          mma_commit(&sc_empty[ss]);
          mma_commit(&dm_full[g & 1]);
        }
        __syncwarp();
        AT_TRACE(0, 2, g);
        if (++ss == kAttnScStages) ss = 0, sp ^= 1;
        if (g >= 1) issue_mma2(g - 1, h > 0 ? h - 1 : H - 1, (int)((h > 0 ? it : it - 1) & 1));
      }
    }
    if (G > 0) issue_mma2(G - 1, H - 1, (int)((my_n - 1) & 1));
  } else if (warp == 2) {
    // ------------------------------------------------------------------ activation I/O: x rows in, LayerNorm rows out
    uint32_t ph_used = 0, ph_staged = 0;
    auto load_x = [&](int64_t n) {
      if (lane == 0) {
        mbar_arrive_expect_tx(xs_full, (uint32_t)V * 512u);
        const float* src = a.x[net] + n * V * 128;
        for (int r = 0; r < V; r++) bulk_g2s(smem + L.xs() + r * kXsRow, src + (size_t)r * 128, 512, xs_full);
      }
      __syncwarp();
    };
    auto store_out = [&](int64_t n) {
      if (lane == 0) {
        float* dst = a.out[net] + n * V * 128;
        for (int r = 0; r < V; r++) bulk_s2g(dst + (size_t)r * 128, smem + L.xs() + r * kXsRow, 512);
        bulk_commit_group();
      }
      __syncwarp();
    };
    if (my_n > 0) load_x(sample_of(0));
    for (int64_t it = 0; it < my_n; it++) {
      mbar_wait(xs_used, ph_used);  // x(it) converted and used for the accumulator: the buffer may carry other rows
      ph_used ^= 1;
      if (it >= 1) {
        mbar_wait(ln_staged, ph_staged);
        ph_staged ^= 1;
        store_out(sample_of(it - 1));
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
      }
      if (it + 1 < my_n) {
        load_x(sample_of(it + 1));
      } else {
        if (lane == 0) mbar_arrive(xs_full);  // nothing more to load: the buffer is free for the last LayerNorm
        __syncwarp();
      }
    }
    if (my_n > 0) {
      mbar_wait(ln_staged, ph_staged);
      store_out(sample_of(my_n - 1));
    }
    if (lane == 0) bulk_wait_group0();
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue groups (kb = feature half)
    const int q = warp & 3;
    const int kb = (warp - 3) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool row_valid = row < V;
    uint32_t ph_dm = 0, ph_xs = 0, ph_used = 1 /* first wait is for the SECOND sample's phase */, ph_dofull = 0;
    uint8_t* my_xs = smem + L.xs() + row * kXsRow + kb * 256;  // this thread's 64 columns of its row in the staging buffer
    float2* stat = reinterpret_cast<float2*>(smem + L.stat());

    // Sample hand-over (every epilogue thread: its token row, its group's 64 features): DO[it & 1] <- x (the residual rides
    // in the accumulator) and the row of the MN-major bf16 hi/lo operand tile of MMA1 -- [atom][64 features] = one swizzled
    // 128-byte row per (thread, hi/lo).  Called after dm_full of the previous sample's LAST head, i.e. when every MMA1 that
    // read the old tiles has retired.
    uint8_t* xb_hi = smem + L.xb() + kb * (VP * 128) + row * 128;
    uint8_t* xb_lo = xb_hi + L.xb_bytes / 2;
    auto prepare_sample = [&](int64_t it) {
      mbar_wait(xs_full, ph_xs);
      ph_xs ^= 1;
      const uint32_t base = tmem + lane_base + AT_DO + (uint32_t)(it & 1) * 128 + kb * 64;
#pragma unroll 1
      for (int b = 0; b < 4; b++) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_valid) v = *reinterpret_cast<const float4*>(my_xs + b * 64 + j * 16);
          r[4 * j] = __float_as_uint(v.x), r[4 * j + 1] = __float_as_uint(v.y);
          r[4 * j + 2] = __float_as_uint(v.z), r[4 * j + 3] = __float_as_uint(v.w);
        }
        tmem_st16(base + b * 16, r);
        if (row < VP) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; j++) split2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), hi[j], lo[j]);
#pragma unroll
          for (int cc = 0; cc < 2; cc++) {
            const uint32_t pos = (((uint32_t)(b * 2 + cc)) ^ ((uint32_t)row & 7u)) << 4;
            *reinterpret_cast<uint4*>(xb_hi + pos) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
            if (kSplit == 3) *reinterpret_cast<uint4*>(xb_lo + pos) = make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(do_init);
      mbar_arrive(xb_full);
      mbar_arrive(xs_used);
    };
    auto layer_norm = [&](int64_t it, bool last) {  // sample `it` finished: drain, normalise, park for the bulk store
      mbar_wait(do_full, ph_dofull);
      ph_dofull ^= 1;
      tc_fence_after();
      uint32_t ya[32], yb[32];
      const uint32_t base = tmem + lane_base + AT_DO + (uint32_t)(it & 1) * 128 + kb * 64;
      tmem_ld32(base, ya);
      tmem_ld32(base + 32, yb);
      tmem_ld_wait();
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const float v0 = __uint_as_float(ya[j]), v1 = __uint_as_float(yb[j]);
        sum += v0 + v1;
        sq = fmaf(v0, v0, sq), sq = fmaf(v1, v1, sq);
      }
      stat[kb * 128 + row] = make_float2(sum, sq);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = stat[(kb ^ 1) * 128 + row];
      sum += other.x, sq += other.y;
      const float mean = sum * (1.f / 128.f);
      const float var = fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f);
      const float rstd = 1.0f / sqrtf(var + a.eps);
      // the staging buffer is free once the NEXT sample's rows have been consumed (or, after the last sample, once the
      // I/O warp has read the previous rows out)
      if (last) {
        mbar_wait(xs_full, ph_xs);
      } else {
        mbar_wait(xs_used, ph_used);
        ph_used ^= 1;
      }
      if (row_valid) {
        auto park32 = [&](const uint32_t (&r)[32], int off) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 gm = *reinterpret_cast<const float4*>(vecs + kb * 64 + off + 4 * j);
            const float4 bt = *reinterpret_cast<const float4*>(vecs + 128 + kb * 64 + off + 4 * j);
            float4 o;
            o.x = (__uint_as_float(r[4 * j]) - mean) * rstd * gm.x + bt.x;
            o.y = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * gm.y + bt.y;
            o.z = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * gm.z + bt.z;
            o.w = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * gm.w + bt.w;
            *reinterpret_cast<float4*>(my_xs + (off + 4 * j) * 4) = o;
          }
        };
        park32(ya, 0);
        park32(yb, 32);
      }
      fence_proxy_async_smem();
      mbar_arrive(ln_staged);
    };

    if (my_n > 0) prepare_sample(0);
    int64_t g = 0;
    for (int64_t it = 0; it < my_n; it++) {
      for (int h = 0; h < H; h++, g++) {
        const int b = (int)(g & 1);
        if (q == 0 && kb == 0) { AT_TRACE(1, 0, g); }
        mbar_wait(&dm_full[b], (ph_dm >> b) & 1u);
        ph_dm ^= 1u << b;
        tc_fence_after();
        if (q == 0 && kb == 0) { AT_TRACE(1, 1, g); }
        {
          const uint32_t base = tmem + lane_base + AT_DM + b * 128 + kb * 64;
          uint32_t r0[32], r1[32];
          tmem_ld32(base, r0);
          tmem_ld32(base + 32, r1);
          tmem_ld_wait();
          if (!row_valid) {  // rows >= V hold scratch: feed zeros to MMA2 (the GPU is power-limited: idle multipliers are cheaper)
#pragma unroll
            for (int j = 0; j < 32; j++) r0[j] = 0u, r1[j] = 0u;
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; j++) split2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]), hi[j], lo[j]);
          tmem_st16(base, hi);
          if (kSplit == 3) tmem_st16(base + 32, lo);
#pragma unroll
          for (int j = 0; j < 16; j++) split2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]), hi[j], lo[j]);
          tmem_st16(base + 16, hi);
          if (kSplit == 3) tmem_st16(base + 48, lo);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&h_full[b * 2 + kb]);
        if (q == 0 && kb == 0) { AT_TRACE(1, 2, g); }
        if (h == 0 && it > 0) layer_norm(it - 1, false);
        if (q == 0 && kb == 0 && h == 0 && it > 0) { AT_TRACE(1, 3, g); }
        if (h == H - 1 && it + 1 < my_n) prepare_sample(it + 1);
        if (q == 0 && kb == 0 && h == H - 1) { AT_TRACE(1, 4, g); }
      }
    }
    if (my_n > 0) layer_norm(my_n - 1, true);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef AT_TRACE
}

// ============================================================================================
// in_mlp: token features (embedding gather + concat, flow.py:172 / custom_transformer_nvp.py:64-71) ->
// Linear(E+9 -> 256) + SiLU -> Linear(256 -> 128)            (mlp.py:18-23, custom_transformer_block.py:66)
// Weights stay resident in shared memory for the whole (persistent) CTA; the 256-wide hidden activation
// goes TMEM -> registers -> TMEM (A operand of the second GEMM).  TMEM: D1 [0,256) (D2 reuses [0,128)), H hi
// [256,384), H lo [384,512).
// SiLU with the hardware exp2/rcp approximations (relative error ~1e-7, far inside the 1e-4 parity budget)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ void epi_bar_sync256() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

constexpr int kInMlpThreads = 832;  // in_mlp: warp 0 weight loader, 1 MMA issuer, 2-17 SiLU epilogue, 18-25 feature gather + output rows

struct InMlpArgs {
  const uint8_t* w1[2];   // [256 x 64] hi 32K | lo 32K
  const uint8_t* w2[2];   // 4 K blocks x (hi 16K | lo 16K)
  const float* b1[2];
  const float* b2[2];
  float* out[2];          // [M,128]
  const float* embed;     // [n_types, E]
  const int64_t* atom_types;
  const float* xc;
  const float* xv;
  const float* z_other;   // [M,3]
  int64_t M, n_cond;
  int V, E, n_types;
  long long* trace;  // debug: event trace of CTA (0,0) (tw_debug_set_trace class 5)
};

struct InMlpSmem {
  static constexpr int A_HI = 0, A_LO = 16384, W1 = 32768, W2 = W1 + 65536, B1 = W2 + 131072, B2 = B1 + 1024, BARS = B2 + 512;
  static constexpr int TOTAL = BARS + 128;
};

// Chunk pipeline (the fused-FFN scheme with two hidden chunks of 128): items g = (tile, chunk c).
//   G1(g):   DH[c] = A_tile W1[c]^T                 12 SS MMAs of N = 128 (A: gathered features, smem; K = 64)
//   SiLU(g): DH[c] fp32 -> silu(. + b1) -> bf16 hi | lo IN PLACE                       [16 epilogue warps: 32 columns per thread]
//   G2(g):   Y[tile & 1] += H(g) W2[:, c]^T          24 TS MMAs of N = 128 (A: TMEM)
//   out:     Y[tile & 1] + b2 -> global rows; gather of the next tile's features       [8 gather / output warps]
// Issue order G1(g), G2(g-1): the tensor pipe executes in issue order, so DH[c] is not overwritten by G1(tile+1, c) before
// G2(tile, c) has read it, and SiLU(g) overlaps G1(g+1) + G2(g-1).  TMEM: DH0 | DH1 | Y0 | Y1 (128 columns each).
// (The first version ran gather -> GEMM1 -> SiLU -> GEMM2 -> store tile by tile on the same 8 warps: 88 us per launch for 26 us
// of tensor work.  A warp retires about one dependent instruction per 4-5 cycles, so the cure is roles that overlap and enough
// warps per role: with 8 SiLU + 4 output warps the row-per-thread output stores alone took 6.5 k cycles per tile -- 32 cache
// lines per store instruction -- hence 256-bit stores and half a row per thread.)
template <int kSplit>
__global__ void __launch_bounds__(kInMlpThreads, 1) k_in_mlp_tc(InMlpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.M + 127) / 128;
  const int64_t my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + InMlpSmem::BARS);
  uint64_t* w_full = bars;        // weights resident
  uint64_t* a_full = bars + 1;    // 256 arrivals: feature operand of a tile written
  uint64_t* a_free = bars + 2;    // commit: G1(tile, 1) retired
  uint64_t* d1_full = bars + 3;   // [2] commit: G1(tile, c) retired
  uint64_t* h_full = bars + 5;    // [2] 512 arrivals: H(tile, c) in TMEM
  uint64_t* y_full = bars + 7;    // [2] commit: G2(tile, 1) retired
  uint64_t* y_free = bars + 9;    // [2] 256 arrivals: Y[tile & 1] read out
  uint64_t* w2_full = bars + 11;  // second-layer weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* b1s = reinterpret_cast<float*>(smem + InMlpSmem::B1);
  float* b2s = reinterpret_cast<float*>(smem + InMlpSmem::B2);
  if (tid == 0) {
    mbar_init(w_full, 1);
    mbar_init(w2_full, 1);
    mbar_init(a_full, 256);
    mbar_init(a_free, 1);
    for (int i = 0; i < 2; i++) mbar_init(&d1_full[i], 1), mbar_init(&h_full[i], 512), mbar_init(&y_full[i], 1), mbar_init(&y_free[i], 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < 256; i += blockDim.x) b1s[i] = a.b1[net][i];
  for (int i = tid; i < 128; i += blockDim.x) b2s[i] = a.b2[net][i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t T_DH = 0, T_Y = 256;
  const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  int tr_n = 0;
#define IM_TRACE(role, ev, item)                                                      \
  if (tr_on && tr_n < 1024) {                                                         \
    a.trace[((role) * 1024 + tr_n) * 2] = (long long)(ev) | ((long long)(item) << 8); \
    a.trace[((role) * 1024 + tr_n) * 2 + 1] = clock64();                              \
    tr_n++;                                                                           \
  }

  if (warp == 0) {
    if (elect_one()) {
      // W1 first, on its own barrier: the first GEMM starts after a third of the bytes (every CTA of a launch pulls the same
      // 192 KB through L2 at the same moment: ~9 us until the last byte)
      const uint32_t w1b = kSplit == 3 ? 65536 : 32768;
      mbar_arrive_expect_tx(w_full, w1b);
      bulk_g2s(smem + InMlpSmem::W1, a.w1[net], w1b, w_full);
      mbar_arrive_expect_tx(w2_full, kSplit == 3 ? 131072 : 4 * 16384);
      for (int i = 0; i < 4; i++) bulk_g2s(smem + InMlpSmem::W2 + i * 32768, a.w2[net] + i * 32768, kSplit == 3 ? 32768 : 16384, w2_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    mbar_wait(w_full, 0);
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t ahi = smem_u32(smem + InMlpSmem::A_HI), alo = smem_u32(smem + InMlpSmem::A_LO);
    const uint32_t w1hi = smem_u32(smem + InMlpSmem::W1), w1lo = w1hi + 32768, w2 = smem_u32(smem + InMlpSmem::W2);
    auto issue_g2 = [&](int64_t gp) {
      const int64_t tp = gp >> 1;
      const int cp = (int)(gp & 1), yb = (int)(tp & 1);
      IM_TRACE(0, 2, gp);
      if (gp == 0) mbar_wait(w2_full, 0);
      mbar_wait(&h_full[cp], (uint32_t)(tp & 1));
      IM_TRACE(0, 3, gp);
      if (cp == 0 && tp >= 2) mbar_wait(&y_free[yb], (uint32_t)(((tp - 2) >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem + T_Y + yb * 128, hh = tmem + T_DH + cp * 128, hl = hh + 64;
        const uint32_t wb = w2 + (uint32_t)cp * 65536;  // K blocks 2 cp, 2 cp + 1
#pragma unroll
        for (int k = 0; k < 8; k++) mma_ts(d, hh + k * 8, desc_kmajor_sw128(wb + (k >> 2) * 32768 + (k & 3) * 32), idesc, (cp | k) != 0);
        if (kSplit == 3) {
#pragma unroll
          for (int k = 0; k < 8; k++) mma_ts(d, hl + k * 8, desc_kmajor_sw128(wb + (k >> 2) * 32768 + (k & 3) * 32), idesc, 1);
#pragma unroll
          for (int k = 0; k < 8; k++) mma_ts(d, hh + k * 8, desc_kmajor_sw128(wb + (k >> 2) * 32768 + 16384 + (k & 3) * 32), idesc, 1);
        }
        if (cp == 1) mma_commit(&y_full[yb]);
      }
      __syncwarp();
      IM_TRACE(0, 4, gp);
    };
    for (int64_t g = 0; g < 2 * my_tiles; g++) {
      const int64_t it = g >> 1;
      const int c = (int)(g & 1);
      IM_TRACE(0, 0, g);
      if (c == 0) mbar_wait(a_full, (uint32_t)(it & 1));
      tc_fence_after();
      IM_TRACE(0, 1, g);
      if (elect_one()) {
        const uint32_t d = tmem + T_DH + c * 128;
        const uint32_t wh = w1hi + (uint32_t)c * 16384, wl = w1lo + (uint32_t)c * 16384;  // hidden units 128 c .. 128 c + 127
#pragma unroll
        for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(ahi + k * 32), desc_kmajor_sw128(wh + k * 32), idesc, k > 0);
        if (kSplit == 3) {
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(alo + k * 32), desc_kmajor_sw128(wh + k * 32), idesc, 1);
#pragma unroll
          for (int k = 0; k < 4; k++) mma_ss(d, desc_kmajor_sw128(ahi + k * 32), desc_kmajor_sw128(wl + k * 32), idesc, 1);
        }
        mma_commit(&d1_full[c]);
        if (c == 1) mma_commit(a_free);
      }
      __syncwarp();
      if (g >= 1) issue_g2(g - 1);
    }
    if (my_tiles > 0) issue_g2(2 * my_tiles - 1);
  } else if (warp < 18) {
    // ------------------------------------------------------------------ SiLU epilogue: D1 + b1 -> SiLU -> hi | lo in place
    const int q = warp & 3;
    const int hq = (warp - 2) >> 2;  // column quarter of the chunk (32 hidden units)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int64_t g = 0; g < 2 * my_tiles; g++) {
      const int c = (int)(g & 1);
      if (warp == 2) { IM_TRACE(1, 0, g); }
      mbar_wait(&d1_full[c], (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      if (warp == 2) { IM_TRACE(1, 1, g); }
      const uint32_t base = tmem + lane_base + T_DH + c * 128;
      uint32_t r[32];
      tmem_ld32(base + hq * 32, r);
      tmem_ld_wait();
      asm volatile("bar.sync 2, 512;" ::: "memory");  // every quarter holds its fp32 columns: the in-place writes below cross quarters
      const float4* bb = reinterpret_cast<const float4*>(b1s + c * 128 + hq * 32);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float4 b4 = bb[j];
        split2(silu_f(__uint_as_float(r[4 * j]) + b4.x), silu_f(__uint_as_float(r[4 * j + 1]) + b4.y), hi[2 * j], lo[2 * j]);
        split2(silu_f(__uint_as_float(r[4 * j + 2]) + b4.z), silu_f(__uint_as_float(r[4 * j + 3]) + b4.w), hi[2 * j + 1], lo[2 * j + 1]);
      }
      tmem_st16(base + hq * 16, hi);
      if (kSplit == 3) tmem_st16(base + 64 + hq * 16, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&h_full[c]);
      if (warp == 2) { IM_TRACE(1, 2, g); }
    }
  } else {
    // ------------------------------------------------------------------ feature gather (next tile) + output rows (this tile)
    const int q = warp & 3;
    const int hf = (warp - 18) >> 2;  // half of the features / of the output row
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int E = a.E;
    const bool own_cond = a.n_cond * a.V == a.M;
    auto gather = [&](int64_t tile) {  // features 32 hf .. 32 hf + 31 of token `row` (E + 9 real ones): bf16 hi / lo K-major SW128 rows
      const int64_t m = tile * 128 + row;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; j++) f[j] = 0.f;
      if (m < a.M) {
        const int64_t smp = m / a.V;
        const int v = (int)(m - smp * a.V);
        const int64_t mc = (own_cond ? smp : smp % a.n_cond) * a.V + v;
        if (E == 32) {
          if (hf == 0) {
            int64_t t = a.atom_types[mc];
            t = t < 0 ? 0 : (t >= a.n_types ? a.n_types - 1 : t);
            const float4* er = reinterpret_cast<const float4*>(a.embed + t * E);
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 e4 = __ldg(er + j);
              f[4 * j] = e4.x, f[4 * j + 1] = e4.y, f[4 * j + 2] = e4.z, f[4 * j + 3] = e4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 3; j++) {
              f[j] = __ldg(a.xc + mc * 3 + j);
              f[3 + j] = __ldg(a.xv + mc * 3 + j);
              f[6 + j] = __ldg(a.z_other + m * 3 + j);
            }
          }
        } else {
          int64_t t = a.atom_types[mc];
          t = t < 0 ? 0 : (t >= a.n_types ? a.n_types - 1 : t);
          const float* er = a.embed + t * E;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int e = hf * 32 + j;
            if (e < E) f[j] = __ldg(er + e);
            else if (e < E + 3) f[j] = __ldg(a.xc + mc * 3 + (e - E));
            else if (e < E + 6) f[j] = __ldg(a.xv + mc * 3 + (e - E - 3));
            else if (e < E + 9) f[j] = __ldg(a.z_other + m * 3 + (e - E - 6));
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; j++) split2(f[c * 8 + 2 * j], f[c * 8 + 2 * j + 1], h[j], l[j]);
        const uint32_t off = row * 128u + (((uint32_t)(hf * 4 + c) ^ (row & 7u)) << 4);
        *reinterpret_cast<uint4*>(smem + InMlpSmem::A_HI + off) = make_uint4(h[0], h[1], h[2], h[3]);
        if (kSplit == 3) *reinterpret_cast<uint4*>(smem + InMlpSmem::A_LO + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(a_full);
    };
    if (my_tiles > 0) gather(blockIdx.x);
    for (int64_t it = 0; it < my_tiles; it++) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      if (warp == 18) { IM_TRACE(2, 0, it); }
      if (it + 1 < my_tiles) {
        mbar_wait(a_free, (uint32_t)(it & 1));  // G1 of this tile has read the operand
        if (warp == 18) { IM_TRACE(2, 1, it); }
        gather(tile + gridDim.x);
        if (warp == 18) { IM_TRACE(2, 2, it); }
      }
      const int yb = (int)(it & 1);
      mbar_wait(&y_full[yb], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (warp == 18) { IM_TRACE(2, 3, it); }
      const int64_t m = tile * 128 + row;
      float* orow = a.out[net] + m * 128 + hf * 64;
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem + lane_base + T_Y + yb * 128 + hf * 64, r0);
      tmem_ld32(tmem + lane_base + T_Y + yb * 128 + hf * 64 + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&y_free[yb]);  // the accumulator half is in registers
      if (m < a.M) {
        const float4* bb = reinterpret_cast<const float4*>(b2s + hf * 64);
#pragma unroll
        for (int j = 0; j < 8; j++) {  // 256-bit stores: a row-per-thread store instruction touches 32 cache lines whatever its width
          const float4 ba = bb[2 * j], bc = bb[2 * j + 1];
          const uint32_t* rr = j < 4 ? &r0[8 * j] : &r1[8 * (j - 4)];
          asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + 8 * j), "f"(__uint_as_float(rr[0]) + ba.x),
                       "f"(__uint_as_float(rr[1]) + ba.y), "f"(__uint_as_float(rr[2]) + ba.z), "f"(__uint_as_float(rr[3]) + ba.w),
                       "f"(__uint_as_float(rr[4]) + bc.x), "f"(__uint_as_float(rr[5]) + bc.y), "f"(__uint_as_float(rr[6]) + bc.z),
                       "f"(__uint_as_float(rr[7]) + bc.w)
                       : "memory");
        }
      }
      if (warp == 18) { IM_TRACE(2, 4, it); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
#undef IM_TRACE
}

// ============================================================================================
// out_mlp: Linear(128 -> 256) + SiLU -> Linear(256 -> 3)       (mlp.py:18-23, custom_transformer_block.py:82)
// The 256 -> 3 layer is evaluated in fp32 on the CUDA cores inside the epilogue (768 FMAs per token).
struct OutMlpArgs {
  const float* x[2];     // [M,128]
  const uint8_t* w3[2];  // 2 K blocks x (hi 32K | lo 32K), [256 x 64] each
  const float* b3[2];
  const float* w4[2];    // [3,256] fp32
  const float* b4[2];
  float* out[2];         // [M,3]
  int64_t M;
};
struct OutMlpSmem {
  static constexpr int X_HI = 0, X_LO = 32768, W3 = 65536, CST = W3 + 131072, RED = CST + 256 * 16 + 16;
  static constexpr int BARS = RED + 2 * 128 * 4 * 4;
  static constexpr int TOTAL = BARS + 128;
};
constexpr int kOutMlpThreads = 576;  // warp 0 weight loader, 1 MMA issuer, 2-9 epilogue (two groups of four), 10-17 x-tile loaders

// Roles overlap: while the epilogue warps run SiLU + the 256 -> 3 layer on the accumulator of tile t (two accumulators), the
// loader warps convert the rows of tile t + 1 (first half prefetched into registers before the operand buffer is free) and the
// tensor pipe runs its GEMM.  (First version: the same 8 warps loaded, waited and reduced tile by tile: 57 us per launch.)
template <int kSplit>
__global__ void __launch_bounds__(kOutMlpThreads, 1) k_out_mlp_tc(OutMlpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  const int net = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.M + 127) / 128;
  const int64_t my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OutMlpSmem::BARS);
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 1;   // 256 arrivals: operand images of a tile written
  uint64_t* d_full = bars + 2;   // [2] commit: GEMM of a tile retired (accumulator ready, operand buffer free)
  uint64_t* d_free = bars + 4;   // [2] 256 arrivals: accumulator read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float4* cst = reinterpret_cast<float4*>(smem + OutMlpSmem::CST);  // per hidden unit: b3, w4[0], w4[1], w4[2]; then b4
  float* red = reinterpret_cast<float*>(smem + OutMlpSmem::RED);    // [2][128][4] partial sums of group 1
  if (tid == 0) {
    mbar_init(w_full, 1);
    mbar_init(x_full, 256);
    for (int i = 0; i < 2; i++) mbar_init(&d_full[i], 1), mbar_init(&d_free[i], 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < 256; i += blockDim.x) cst[i] = make_float4(a.b3[net][i], a.w4[net][i], a.w4[net][256 + i], a.w4[net][512 + i]);
  if (tid == 0) cst[256] = make_float4(a.b4[net][0], a.b4[net][1], a.b4[net][2], 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t part = kSplit == 3 ? 65536 : 32768;
      mbar_arrive_expect_tx(w_full, 2 * part);
      bulk_g2s(smem + OutMlpSmem::W3, a.w3[net], part, w_full);
      bulk_g2s(smem + OutMlpSmem::W3 + 65536, a.w3[net] + 65536, part, w_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    mbar_wait(w_full, 0);
    const uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
    const uint32_t xhi = smem_u32(smem + OutMlpSmem::X_HI), xlo = smem_u32(smem + OutMlpSmem::X_LO), w3 = smem_u32(smem + OutMlpSmem::W3);
    uint32_t ph_free[2] = {0, 0};
    for (int64_t it = 0; it < my_tiles; it++) {
      const int tb = (int)(it & 1);
      mbar_wait(x_full, (uint32_t)(it & 1));
      if (it >= 2) {
        mbar_wait(&d_free[tb], ph_free[tb]);
        ph_free[tb] ^= 1;
      }
      tc_fence_after();
      const uint32_t d = tmem + tb * 256;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          uint32_t ao = (k >> 2) * 16384 + (k & 3) * 32, wo = (k >> 2) * 65536 + (k & 3) * 32;
          mma_ss(d, desc_kmajor_sw128(xhi + ao), desc_kmajor_sw128(w3 + wo), idesc, k > 0);
        }
        if (kSplit == 3) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            uint32_t ao = (k >> 2) * 16384 + (k & 3) * 32, wo = (k >> 2) * 65536 + (k & 3) * 32;
            mma_ss(d, desc_kmajor_sw128(xlo + ao), desc_kmajor_sw128(w3 + wo), idesc, 1);
          }
#pragma unroll
          for (int k = 0; k < 8; k++) {
            uint32_t ao = (k >> 2) * 16384 + (k & 3) * 32, wo = (k >> 2) * 65536 + 32768 + (k & 3) * 32;
            mma_ss(d, desc_kmajor_sw128(xhi + ao), desc_kmajor_sw128(w3 + wo), idesc, 1);
          }
        }
        mma_commit(&d_full[tb]);
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue: SiLU(D + b3) . w4 + b4 (hidden half hf of row `row`)
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t ph_full[2] = {0, 0};
    for (int64_t it = 0; it < my_tiles; it++) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int tb = (int)(it & 1);
      mbar_wait(&d_full[tb], ph_full[tb]);
      ph_full[tb] ^= 1;
      tc_fence_after();
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int g = 0; g < 4; g++) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + tb * 256 + hf * 128 + g * 32, r);
        tmem_ld_wait();
        const float4* cc = cst + hf * 128 + g * 32;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 c4 = cc[j];
          const float v = silu_f(__uint_as_float(r[j]) + c4.x);
          s0 = fmaf(v, c4.y, s0);
          s1 = fmaf(v, c4.z, s1);
          s2 = fmaf(v, c4.w, s2);
        }
      }
      tc_fence_before();
      mbar_arrive(&d_free[tb]);
      float* rr = red + (tb * 128 + row) * 4;
      if (hf == 1) rr[0] = s0, rr[1] = s1, rr[2] = s2;
      epi_bar_sync256();
      if (hf == 0) {
        const int64_t m = tile * 128 + row;
        if (m < a.M) {
          const float4 b4 = cst[256];
          float* o = a.out[net] + m * 3;
          o[0] = (s0 + rr[0]) + b4.x;
          o[1] = (s1 + rr[1]) + b4.y;
          o[2] = (s2 + rr[2]) + b4.z;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ x-tile loaders: 16 rows per warp, a float4 column per lane
    const int lw = warp - 10;  // 0..7
    const float* x = a.x[net];
    uint32_t ph_full[2] = {0, 0};
    float4 v[8];
    auto load8 = [&](int64_t tile, int half) {
      const int64_t row0 = tile * 128 + lw * 16 + half * 8;
#pragma unroll
      for (int u = 0; u < 8; u++)
        v[u] = (row0 + u < a.M) ? __ldg(reinterpret_cast<const float4*>(x + (row0 + u) * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store8 = [&](int half) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int r = lw * 16 + half * 8 + u;
        uint32_t h0, l0, h1, l1;
        split2(v[u].x, v[u].y, h0, l0);
        split2(v[u].z, v[u].w, h1, l1);
        const uint32_t off = sw128_offset(r, lane * 4, 128);
        *reinterpret_cast<uint2*>(smem + OutMlpSmem::X_HI + off) = make_uint2(h0, h1);
        if (kSplit == 3) *reinterpret_cast<uint2*>(smem + OutMlpSmem::X_LO + off) = make_uint2(l0, l1);
      }
    };
    if (my_tiles > 0) load8(blockIdx.x, 0);
    for (int64_t it = 0; it < my_tiles; it++) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      if (it >= 1) {  // the GEMM of the previous tile has read the operand images
        const int pb = (int)((it - 1) & 1);
        mbar_wait(&d_full[pb], ph_full[pb]);
        ph_full[pb] ^= 1;
      }
      store8(0);
      load8(tile, 1);
      store8(1);
      fence_proxy_async_smem();
      mbar_arrive(x_full);
      if (it + 1 < my_tiles) load8(tile + gridDim.x, 0);  // first half of the next tile: in registers while this tile's GEMM runs
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ============================================================================================
// orchestration of one conditioner pair (scale net, shift net) of coupling layer k
static int pad16(int v) { return (v + 15) / 16 * 16; }

void tc_carve(const tw_flow_config* c, int64_t n, int64_t n_cond, int64_t V, Arena& ar, TcScratch* out) {
  const int64_t M = n * V;
  const int64_t tiles = (M + 127) / 128;
  const int VP = pad16((int)V);
  const size_t tile_bytes = (size_t)c->num_heads * 2 * 2 * 16384;
  for (int i = 0; i < 2; i++) {
    ar.off = align_up(ar.off, 1024);
    out->mixed_img[i] = ar.take<uint8_t>(tiles * tile_bytes);
  }
  ar.off = align_up(ar.off, 1024);
  out->scores_img = ar.take<uint8_t>(V > 128 ? 0 : (size_t)n_cond * c->num_heads * 2 * VP * VP * 2);  // (V > 128: CUDA-core attention)
  ar.off = align_up(ar.off, 1024);
  out->ffn_tail = reinterpret_cast<float*>(ar.take<uint8_t>(kFfnTailBytes));
}

// scores -> operand images, once per pass
int tc_scores_images(const tw_flow_config* c, const float* scores, int64_t n_cond, int V, uint8_t* img, int transpose, cudaStream_t st) {
  TW_CHECK_ARG(V <= 128, "tensor-core attention supports at most 128 atoms per sample");
  const int VP = pad16(V);
  int64_t rows = n_cond * c->num_heads * VP;
  if (rows == 0) return TW_OK;
  k_scores_img<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(scores, n_cond, V, VP, c->num_heads, img, transpose);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

int tc_begin_pass(const tw_flow_config* c, const ParamView&, TcScratch& tc, const float* scores, const uint8_t*, int64_t,
                  int64_t n_cond, int V, cudaStream_t st) {
  if (tc.ffn_tail) TW_CUDA(cudaMemsetAsync(tc.ffn_tail + kFfnTailFloats, 0, kFfnTailBytes - kFfnTailFloats * 4, st));  // (the counters; the partial sums are overwritten)
  if (!(tc_stage_mask() & TC_MIX) || scores == nullptr) return TW_OK;
  return tc_scores_images(c, scores, n_cond, V, tc.scores_img, 0, st);
}

bool tc_scores_direct_supported(int V) {  // the CUDA-core attention fallback (bring-up switch TW_TC_STAGES) still reads fp32 scores
  return (tc_stage_mask() & TC_MIX) && (tc_stage_mask() & TC_ATTN_PROJ) && V <= 128;
}

// Inference: score images straight from the centred conditioning coordinates (no fp32 score tensor).
int tc_begin_pass_direct(const tw_flow_config* c, TcScratch& tc, const float* xc, const uint8_t* mask, const float* lengthscales,
                         int64_t n_cond, int V, cudaStream_t st, const float* cheb, bool clear_tail) {
  if (tc.ffn_tail && clear_tail) TW_CUDA(cudaMemsetAsync(tc.ffn_tail + kFfnTailFloats, 0, kFfnTailBytes - kFfnTailFloats * 4, st));
  const int VP = pad16(V);
  const size_t smem = (((size_t)V * 12 + 15) & ~(size_t)15) + (((size_t)V + 15) & ~(size_t)15) + (size_t)V * (V | 1) * 4;
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_scores_direct_img, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 129 * 4 + 4096));
    attr_done.mark();
  }
  if (n_cond < 1) return TW_OK;
  int threads = (c->num_heads * VP + 31) & ~31;
  if (threads > 1024) threads = 1024;
  k_scores_direct_img<<<(unsigned)n_cond, threads, smem, st>>>(xc, mask, lengthscales, V, VP, c->num_heads, tc.scores_img, cheb,
                                                             c->cheb_order, c->force_asymptotic_zero);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

size_t tc_scores_img_bytes(const tw_flow_config* c, int64_t n_cond, int V) {
  const int VP = pad16(V);
  return (size_t)n_cond * c->num_heads * 2 * VP * VP * 2;
}
size_t tc_mixed_img_bytes(const tw_flow_config* c, int64_t M) { return (size_t)((M + 127) / 128) * c->num_heads * 2 * 2 * 16384; }

// mixed_h = A_h x for every head, written as A-operand images (both networks)
int tc_mix(const tw_flow_config* c, const float* const x[2], uint8_t* const img[2], const uint8_t* scores_img, int64_t n,
           int64_t n_cond, int V, cudaStream_t st, int nets) {
  static DeviceOnce attr_done;
  const int VP = pad16(V), H = c->num_heads;
  const uint32_t stage_stride = (uint32_t)((2 * VP * VP * 2 + 1023) & ~1023);
  const int mix_stages = VP > 96 ? 2 : 3;
  const int mix_groups = VP > 96 ? 1 : 2;  // epilogue groups (a second staging buffer must fit next to the ring)
  const int mix_smem = mix_stages * stage_stride + mix_groups * VP * 512 + 256 + 1024;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_mix_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 2048));
    TW_CUDA(cudaFuncSetAttribute(k_mix_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 2048));
    static_assert(2 * 65536 + 128 * 512 + 2048 <= 3 * 65536 + 2048, "mix smem budget");
    attr_done.mark();
  }
  MixArgs a{};
  for (int s = 0; s < 2; s++) a.x[s] = x[s], a.img[s] = img[s];
  a.scores_img = scores_img;
  a.n = n, a.n_cond = n_cond, a.V = V, a.VP = VP, a.H = H, a.n_stages = mix_stages;
  a.trace = g_mix_trace;
  const int per_net = nets == 1 ? 148 : 74;  // one network alone (chebyshev_kernel: per-network scores) gets every SM
  int gx = (int)(n < per_net ? n : per_net);
  if (gx < 1) return TW_OK;
  dim3 grid(gx, nets);
  static int use_tok = -1;
  if (use_tok < 0) {
    const char* e = getenv("TW_MIX_TOK");  // bring-up switch: 0 = feature-major kernel for every atom count
    use_tok = e ? atoi(e) : 1;
    const int max_smem = 232448;
    TW_CUDA(cudaFuncSetAttribute(k_mix_tok<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    TW_CUDA(cudaFuncSetAttribute(k_mix_tok<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    TW_CUDA(cudaFuncSetAttribute((k_mix_tok<3, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    TW_CUDA(cudaFuncSetAttribute((k_mix_tok<1, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  }
  if (use_tok && VP <= kMixTokMaxVP) {
    int tok_stages = kMixTokStages;
    if ((int)MixTokSmem(V, VP, tok_stages).total() + 1024 > 232448) tok_stages = 2;
    const int smem_tok = (int)MixTokSmem(V, VP, tok_stages).total() + 1024;
    a.n_stages = tok_stages;
    if (smem_tok <= 112 * 1024 && n > per_net) {  // two CTAs per SM
      const int gx2 = (int)(n < 2 * per_net ? n : 2 * per_net);
      if (c->precision == TW_PRECISION_BF16X3)
        k_mix_tok<3, 2><<<dim3(gx2, nets), kMixTokThreads, smem_tok, st>>>(a);
      else
        k_mix_tok<1, 2><<<dim3(gx2, nets), kMixTokThreads, smem_tok, st>>>(a);
      TW_LAUNCH_CHECK();
      return TW_OK;
    }
    if (c->precision == TW_PRECISION_BF16X3)
      k_mix_tok<3><<<grid, kMixTokThreads, smem_tok, st>>>(a);
    else
      k_mix_tok<1><<<grid, kMixTokThreads, smem_tok, st>>>(a);
    TW_LAUNCH_CHECK();
    return TW_OK;
  }
  if (c->precision == TW_PRECISION_BF16X3)
    k_mix_tc<3><<<grid, 64 + 128 * mix_groups, mix_smem, st>>>(a);
  else
    k_mix_tc<1><<<grid, 64 + 128 * mix_groups, mix_smem, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

// out = LN1(x + attention(x)) for both networks of encoder layer t; `pre` (optional) receives the pre-LayerNorm sum
int tc_attention_layer(const tw_flow_config* c, const ParamView& pv, int k, int t, const TcScratch& tc, float* const x_in[2],
                       float* const out_in[2], int64_t n, int64_t n_cond, int V, cudaStream_t st, float* const* pre, int only_net) {
  static DeviceOnce attr_done;
  const int H = c->num_heads;
  // only_net >= 0: run ONE network (its pointers in both slots, grid.y = 1) -- chebyshev_kernel attention has a different
  // score image per network and layer, written to tc.scores_img right before this call
  const int nets = only_net >= 0 ? 1 : 2;
  const int net_of[2] = {only_net >= 0 ? only_net : 0, only_net >= 0 ? only_net : 1};
  float* const x[2] = {x_in[net_of[0]], x_in[net_of[1]]};
  float* const out[2] = {out_in[net_of[0]], out_in[net_of[1]]};
  const int per_net = nets == 1 ? 148 : 74;
  const int proj_smem = kProjStages * kProjStageBytes + 1024 + 256 + 1024;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_proj_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, proj_smem));
    TW_CUDA(cudaFuncSetAttribute(k_proj_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, proj_smem));
    attr_done.mark();
  }
  TcLayout L = TcLayout::make(c);
  const int64_t M = n * V;
  ProfScope prof(PROF_ATTN, st);
  {
    static int use_fused = -1;
    if (use_fused < 0) {
      const char* e = getenv("TW_ATTN_FUSED");  // bring-up switch: 0 = mixing + projection kernels
      use_fused = e ? atoi(e) : 1;
    }
    static DeviceOnce fused_attr;
    if (!fused_attr.done()) {
      TW_CUDA(cudaFuncSetAttribute(k_attn_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      TW_CUDA(cudaFuncSetAttribute(k_attn_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      fused_attr.mark();
    }
    const int VP = pad16(V);
    // inference: the feature-major kernel (values projected first, mixed per sample afterwards); the taped training forward
    // keeps the two-kernel form because the backward reads the per-head mixed images
    if (use_fused && !pre && tc_attn_fm_supported(V, n)) {
      const uint8_t* wcs[2];
      const float *gs[2], *bs[2];
      for (int s = 0; s < 2; s++) {
        wcs[s] = tc.packed + L.net_offset(k, net_of[s]) + L.enc0 + (size_t)t * L.enc_stride + L.enc_wc;
        gs[s] = pv.enc(k, net_of[s], t, 7), bs[s] = pv.enc(k, net_of[s], t, 8);
      }
      if (tc_attn_fm3_supported(V, n)) return tc_attn_fm3(c, x, out, wcs, gs, bs, tc.scores_img, n, n_cond, V, nets, st);  // <= 80 tokens per group
      return tc_attn_fm(c, x, out, wcs, gs, bs, tc.scores_img, n, n_cond, V, nets, st);
    }
    const int fused_smem = (int)AttnSmem(V, VP).total() + 1024;
    // the training tape needs the per-head mixed images, so the taped forward keeps the two-kernel form
    if (use_fused && !pre && VP <= kMixTokMaxVP && fused_smem <= 232448 && n >= 1) {
      AttnArgs a{};
      for (int s = 0; s < 2; s++) {
        a.x[s] = x[s], a.out[s] = out[s];
        a.wc[s] = tc.packed + L.net_offset(k, net_of[s]) + L.enc0 + (size_t)t * L.enc_stride + L.enc_wc;
        a.gamma[s] = pv.enc(k, net_of[s], t, 7), a.beta[s] = pv.enc(k, net_of[s], t, 8);
      }
      a.scores_img = tc.scores_img;
      a.n = n, a.n_cond = n_cond, a.V = V, a.VP = VP, a.H = H, a.eps = c->layer_norm_eps;
      a.trace = g_attn_trace;
      dim3 grid((unsigned)(n < per_net ? n : per_net), nets);
      if (c->precision == TW_PRECISION_BF16X3)
        k_attn_fused<3><<<grid, kAttnThreads, fused_smem, st>>>(a);
      else
        k_attn_fused<1><<<grid, kAttnThreads, fused_smem, st>>>(a);
      TW_LAUNCH_CHECK();
      return TW_OK;
    }
  }
  uint8_t* const mixed[2] = {tc.mixed_img[net_of[0]], tc.mixed_img[net_of[1]]};
  TW_TRY(tc_mix(c, x, mixed, tc.scores_img, n, n_cond, V, st, nets));
  {
    ProjArgs a{};
    for (int s = 0; s < 2; s++) {
      a.a_img[s] = mixed[s];
      a.w[s] = tc.packed + L.net_offset(k, net_of[s]) + L.enc0 + (size_t)t * L.enc_stride + L.enc_wc;
      a.resid[s] = x[s];
      a.out[s] = out[s];
      a.pre[s] = pre ? pre[net_of[s]] : nullptr;
      a.gamma[s] = pv.enc(k, net_of[s], t, 7);
      a.beta[s] = pv.enc(k, net_of[s], t, 8);
    }
    a.M = M, a.KB = H * 2, a.eps = c->layer_norm_eps;
    int64_t n_tiles = (M + 127) / 128;
    int gx = (int)(n_tiles < per_net ? n_tiles : per_net);
    dim3 grid(gx, nets);
    if (c->precision == TW_PRECISION_BF16X3)
      k_proj_tc<3><<<grid, 192, proj_smem, st>>>(a);
    else
      k_proj_tc<1><<<grid, 192, proj_smem, st>>>(a);
    TW_LAUNCH_CHECK();
  }
  return TW_OK;
}

int tc_ffn_layer(const tw_flow_config* c, const ParamView& pv, int k, int t, const TcScratch& tc, float* const x[2],
                 float* const out[2], int64_t M, cudaStream_t st, float* const* pre) {
  TcLayout L = TcLayout::make(c);
  FfnArgs a{};
  for (int s = 0; s < 2; s++) {
    a.x[s] = x[s];
    a.out[s] = out[s];
    a.pre[s] = pre ? pre[s] : nullptr;
    a.w[s] = tc.packed + L.net_offset(k, s) + L.enc0 + (size_t)t * L.enc_stride + L.enc_ffn;
    a.b1[s] = pv.enc(k, s, t, 4);
    a.b2[s] = pv.enc(k, s, t, 6);
    a.gamma[s] = pv.enc(k, s, t, 9);
    a.beta[s] = pv.enc(k, s, t, 10);
  }
  a.M = M;
  a.F = c->dim_feedforward;
  a.eps = c->layer_norm_eps;
  static int use_split = -1;
  if (use_split < 0) {
    const char* e = getenv("TW_FFN_SPLIT");  // bring-up switch: 0 = leftover tiles run whole on a few pairs
    use_split = e ? atoi(e) : 1;
  }
  if (tc.ffn_tail && use_split) {
    int* cnt = reinterpret_cast<int*>(tc.ffn_tail + kFfnTailFloats);
    for (int s = 0; s < 2; s++) {
      a.tail[s] = tc.ffn_tail + (size_t)s * (kFfnTailFloats / 2);
      a.tail_cnt[s] = cnt + s * kFfnTailTiles * 4;
    }
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("TW_FFN_DBG");
      dbg = e ? atoi(e) : 0;
    }
    a.dbg = dbg;
  }
  return launch_ffn_tc(c, a, st);
}

int tc_in_mlp(const tw_flow_config* c, const ParamView& pv, int k, const TcScratch& tc, const int64_t* atom_types, const float* xc,
              const float* xv, const float* z_other, float* const out[2], int64_t n, int64_t n_cond, int V, cudaStream_t st) {
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_in_mlp_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, InMlpSmem::TOTAL + 1024));
    TW_CUDA(cudaFuncSetAttribute(k_in_mlp_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, InMlpSmem::TOTAL + 1024));
    attr_done.mark();
  }
  TcLayout L = TcLayout::make(c);
  InMlpArgs a{};
  for (int s = 0; s < 2; s++) {
    a.w1[s] = tc.packed + L.net_offset(k, s) + L.in_w1;
    a.w2[s] = tc.packed + L.net_offset(k, s) + L.in_w2;
    a.b1[s] = pv.in_b(k, s, 0);
    a.b2[s] = pv.in_b(k, s, 1);
    a.out[s] = out[s];
  }
  a.embed = pv.embed(), a.atom_types = atom_types, a.xc = xc, a.xv = xv, a.z_other = z_other;
  a.M = n * V, a.n_cond = n_cond, a.V = V, a.E = c->atom_embedding_dim, a.n_types = c->num_atom_types;
  a.trace = g_inmlp_trace;
  int64_t n_tiles = (a.M + 127) / 128;
  dim3 grid((unsigned)(n_tiles < 74 ? n_tiles : 74), 2);
  ProfScope prof(PROF_MLP, st);
  if (c->precision == TW_PRECISION_BF16X3)
    k_in_mlp_tc<3><<<grid, kInMlpThreads, InMlpSmem::TOTAL + 1024, st>>>(a);
  else
    k_in_mlp_tc<1><<<grid, kInMlpThreads, InMlpSmem::TOTAL + 1024, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

int tc_out_mlp(const tw_flow_config* c, const ParamView& pv, int k, const TcScratch& tc, float* const x[2], float* const out[2],
               int64_t M, cudaStream_t st) {
  static DeviceOnce attr_done;
  if (!attr_done.done()) {
    TW_CUDA(cudaFuncSetAttribute(k_out_mlp_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutMlpSmem::TOTAL + 1024));
    TW_CUDA(cudaFuncSetAttribute(k_out_mlp_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutMlpSmem::TOTAL + 1024));
    attr_done.mark();
  }
  TcLayout L = TcLayout::make(c);
  OutMlpArgs a{};
  for (int s = 0; s < 2; s++) {
    a.x[s] = x[s];
    a.w3[s] = tc.packed + L.net_offset(k, s) + L.out_w1;
    a.b3[s] = pv.out_b(k, s, 0);
    a.w4[s] = pv.out_w(k, s, 1);
    a.b4[s] = pv.out_b(k, s, 1);
    a.out[s] = out[s];
  }
  a.M = M;
  int64_t n_tiles = (M + 127) / 128;
  dim3 grid((unsigned)(n_tiles < 74 ? n_tiles : 74), 2);
  ProfScope prof(PROF_MLP, st);
  if (c->precision == TW_PRECISION_BF16X3)
    k_out_mlp_tc<3><<<grid, kOutMlpThreads, OutMlpSmem::TOTAL + 1024, st>>>(a);
  else
    k_out_mlp_tc<1><<<grid, kOutMlpThreads, OutMlpSmem::TOTAL + 1024, st>>>(a);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // namespace tw

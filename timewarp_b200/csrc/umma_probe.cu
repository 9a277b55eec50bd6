// Hardware probe for the tcgen05 building blocks (debug entry point, used by tools/umma_probe.py
// and tests/test_gpu_umma.py): one CTA computes out[128,N] = A[128,K] * B[N,K]^T with bf16 inputs
// through a selectable operand placement / layout, so that every descriptor convention the
// production kernels rely on is checked against a CPU result on the real chip.
#include "common.cuh"
#include "umma.cuh"

namespace tw {
using namespace umma;

struct ProbeArgs {
  const float* A;  // [128,K]
  const float* B;  // [N,K]
  float* out;      // [128,N]
  int N, K;
  int a_mode;  // 0 smem K-major SW128, 1 smem K-major no-swizzle, 2 smem MN-major SW128, 3 TMEM (TS)
  int b_mode;  // 0 K-major SW128, 1 K-major no-swizzle, 2 MN-major SW128
  int d_col;   // column offset of the accumulator inside the TMEM allocation
  int a_col;   // column offset of A in TMEM (a_mode 3)
  int* status;  // 0 ok, 1 timeout
};

__device__ __forceinline__ uint32_t off_kmajor_none(uint32_t row, uint32_t k, uint32_t K) {
  return (row >> 3) * ((K >> 3) * 128u) + (k >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}
__device__ __forceinline__ uint32_t off_mnmajor_sw128(uint32_t mn, uint32_t k, uint32_t K) {
  return (mn >> 6) * ((K >> 3) * 1024u) + (k >> 3) * 1024u + (k & 7u) * 128u + ((((mn & 63u) >> 3) ^ (k & 7u)) << 4) + (mn & 7u) * 2u;
}

__global__ void __launch_bounds__(128, 1) k_umma_probe(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  uint8_t* sA = smem;                      // up to 128*256*2 = 64 KB
  uint8_t* sB = smem + 64 * 1024;          // up to 256*256*2 = 128 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, K = p.K;

  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // ---- operands -> smem (bf16), zero-filled first
  for (int i = tid; i < (64 + 128) * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (p.a_mode != 3) {
    for (int e = tid; e < 128 * K; e += 128) {
      int r = e / K, k = e % K;
      uint32_t off = p.a_mode == 0 ? sw128_offset(r, k, 128) : (p.a_mode == 1 ? off_kmajor_none(r, k, K) : off_mnmajor_sw128(r, k, K));
      *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(p.A[r * K + k]);
    }
  } else {
    // A -> TMEM: lane = row, column c holds elements (2c, 2c+1) as packed bf16x2
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        int k = 2 * (c0 + j);
        r[j] = (k < K) ? pack_bf16x2(p.A[tid * K + k], p.A[tid * K + k + 1]) : 0u;
      }
      tmem_st16(tmem + lane_base + p.a_col + c0, r);
    }
    tmem_st_wait();
  }
  for (int e = tid; e < N * K; e += 128) {
    int r = e / K, k = e % K;
    uint32_t off = p.b_mode == 0 ? sw128_offset(r, k, N) : (p.b_mode == 1 ? off_kmajor_none(r, k, K) : off_mnmajor_sw128(r, k, K));
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(p.B[r * K + k]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N, p.a_mode == 2 ? 1 : 0, p.b_mode == 2 ? 1 : 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint64_t bd;
      if (p.b_mode == 0) bd = desc_kmajor_sw128(b0 + (k0 >> 6) * (N * 128) + (k0 & 63) * 2);
      else if (p.b_mode == 1) bd = make_smem_desc(b0 + (k0 >> 3) * 128, 128, (K >> 3) * 128, LAYOUT_NONE);
      else bd = make_smem_desc(b0 + (k0 >> 3) * 1024, (K >> 3) * 1024, 1024, LAYOUT_SW128);
      if (p.a_mode == 3) {
        mma_ts(tmem + p.d_col, tmem + p.a_col + (k0 >> 1), bd, idesc, k0 > 0);
      } else {
        uint64_t ad;
        if (p.a_mode == 0) ad = desc_kmajor_sw128(a0 + (k0 >> 6) * (128 * 128) + (k0 & 63) * 2);
        else if (p.a_mode == 1) ad = make_smem_desc(a0 + (k0 >> 3) * 128, 128, (K >> 3) * 128, LAYOUT_NONE);
        else ad = make_smem_desc(a0 + (k0 >> 3) * 1024, (K >> 3) * 1024, 1024, LAYOUT_SW128);
        mma_ss(tmem + p.d_col, ad, bd, idesc, k0 > 0);
      }
    }
    mma_commit(&bar);
  }
  // wait with a timeout so that a wrong assumption cannot hang the box
  {
    long long t0 = clock64();
    bool ok = false;
    while (!(ok = mbar_try_wait(&bar, 0))) {
      if (clock64() - t0 > 2000000000LL) break;
    }
    if (!ok && tid == 0) *p.status = 1;
  }
  tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + lane_base + p.d_col + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (c0 + j < N) p.out[tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Timing micro-benchmark: an elected lane issues `n_mma` back-to-back M128 x N x K16 MMAs (SS or TS, B operand
// swizzled or not) into one accumulator, ONE commit at the end.  out[0] = cycles to issue, out[1] = cycles until done.
__global__ void __launch_bounds__(128, 1) k_umma_timing(int n_mma, int N, int ts, int b_noswizzle, long long* out, int fill, int a_units,
                                                        int a_off, int b_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 192 * 1024 / 16; i += 128) {  // fill: 0 zeros, 1 pseudo-random bf16 values of magnitude ~1
    uint32_t h = (uint32_t)i * 2654435761u;
    auto v = [&](uint32_t x) { x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15; return ((x & 0x80008000u) | 0x3f003f00u | (x & 0x007f007fu)); };
    reinterpret_cast<uint4*>(smem)[i] = fill ? make_uint4(v(h), v(h + 1), v(h + 2), v(h + 3)) : make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem + a_off), b0 = smem_u32(smem + b_off);  // byte offsets of the A units (16 KB each) and the B tile
    long long t0 = clock64(), t1 = 0;
    if (elect_one()) {
      // descriptors precomputed, issue loop unrolled by 8: the loop measures the tensor pipe, not the issuing thread
      uint64_t ad[8], bd[8];
      uint32_t at[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        ad[j] = desc_kmajor_sw128(a0 + ((j >> 2) % a_units) * 16384 + (j & 3) * 32);
        bd[j] = b_noswizzle ? make_smem_desc(b0 + (j & 3) * 256, 128, 1024, LAYOUT_NONE) : desc_kmajor_sw128(b0 + (j & 3) * 32);
        at[j] = tmem + 256 + j * 8;
      }
      if (ts) {
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; j++) mma_ts(tmem, at[j], bd[j], idesc, 1);
        }
      } else {
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; j++) mma_ss(tmem, ad[j], bd[j], idesc, 1);
        }
      }
      mma_commit(&bar);
      t1 = clock64();
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (t1 != 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// Cost of the bookkeeping around MMA blocks: `n_blocks` blocks of `per_block` TS MMAs (N columns), with optionally, per block,
// flags bit 0: tcgen05.commit to a (never waited) mbarrier, bit 1: tcgen05.fence::after_thread_sync, bit 2: elect.sync +
// __syncwarp around the block (otherwise ONE elected thread runs the whole loop), bit 3: a try_wait on a completed mbarrier,
// bit 4: two commits per block.  out[1] = total cycles.
__global__ void __launch_bounds__(128, 1) k_umma_overhead(int n_blocks, int per_block, int N, int flags, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // (pointer arithmetic keeps the shared address space: LDS / STS)
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (tid == 0) {
    mbar_init(&bar, 1), mbar_init(&bar2, 1), mbar_init(&bar3, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) mbar_arrive(&bar3);  // completes phase 0 of bar3: try_wait(bar3, 0) succeeds immediately from now on
  __syncthreads();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t b0 = smem_u32(smem);
    uint64_t bd[4];
    for (int j = 0; j < 4; j++) bd[j] = desc_kmajor_sw128(b0 + j * 32);
    const long long t0 = clock64();
    auto block = [&]() {
      for (int j = 0; j < per_block; j++) mma_ts(tmem, tmem + 256 + (j & 7) * 8, bd[j & 3], idesc, 1);
      if (flags & 1) mma_commit(&bar2);
      if (flags & 16) mma_commit(&bar2);
    };
    if (flags & 4) {
      for (int i = 0; i < n_blocks; i++) {
        if (flags & 8) mbar_wait(&bar3, 0);
        if (flags & 2) tc_fence_after();
        if (elect_one()) block();
        __syncwarp();
      }
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
    } else {
      if (elect_one()) {
        for (int i = 0; i < n_blocks; i++) {
          if (flags & 8) mbar_wait(&bar3, 0);
          if (flags & 2) tc_fence_after();
          block();
        }
        mma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if ((tid & 31) == 0) out[0] = n_blocks, out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace tw

using namespace tw;

extern "C" int tw_debug_umma_probe(const float* A, const float* B, float* out, int N, int K, int a_mode, int b_mode, int d_col,
                                   int a_col, int* status, void* stream) {
  TW_CHECK_ARG(A && B && out && status, "NULL pointer");
  TW_CHECK_ARG(N >= 8 && N <= 256 && N % 8 == 0 && K >= 16 && K <= 256 && K % 16 == 0, "bad N/K");
  TW_CHECK_ARG(d_col >= 0 && d_col + N <= 512 && a_col >= 0 && a_col + K / 2 <= 512, "bad TMEM columns");
  const int smem = (64 + 128) * 1024 + 1024;
  TW_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ProbeArgs p{A, B, out, N, K, a_mode, b_mode, d_col, a_col, status};
  k_umma_probe<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

extern "C" int tw_debug_umma_timing2(int n_mma, int N, int ts, int b_noswizzle, long long* out, int fill, int a_units, void* stream) {
  TW_CHECK_ARG(out && n_mma > 0 && N >= 16 && N <= 256 && N % 16 == 0, "bad args");
  TW_CHECK_ARG(a_units >= 1 && a_units <= 8, "bad a_units");
  const int smem = 193 * 1024 + 1024;
  TW_CUDA(cudaFuncSetAttribute(k_umma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_timing<<<1, 128, smem, (cudaStream_t)stream>>>(n_mma, N, ts, b_noswizzle, out, fill, a_units, 0, 128 * 1024);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

extern "C" int tw_debug_umma_timing3(int n_mma, int N, int ts, long long* out, int fill, int a_units, int a_off, int b_off, void* stream) {
  TW_CHECK_ARG(out && n_mma > 0 && N >= 16 && N <= 256 && N % 16 == 0 && a_units >= 1 && a_units <= 8, "bad args");
  TW_CHECK_ARG(a_off >= 0 && b_off >= 0 && a_off % 1024 == 0 && b_off % 1024 == 0 && a_off + a_units * 16384 <= 192 * 1024 &&
                   b_off + 32768 <= 192 * 1024, "bad offsets");
  const int smem = 193 * 1024 + 1024;
  TW_CUDA(cudaFuncSetAttribute(k_umma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_timing<<<1, 128, smem, (cudaStream_t)stream>>>(n_mma, N, ts, 0, out, fill, a_units, a_off, b_off);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

extern "C" int tw_debug_umma_timing(int n_mma, int N, int ts, int b_noswizzle, long long* out, void* stream) {
  TW_CHECK_ARG(out && n_mma > 0 && N >= 16 && N <= 256 && N % 16 == 0, "bad args");
  const int smem = 193 * 1024 + 1024;
  TW_CUDA(cudaFuncSetAttribute(k_umma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_timing<<<1, 128, smem, (cudaStream_t)stream>>>(n_mma, N, ts, b_noswizzle, out, 0, 1, 0, 32 * 1024);  // the round-1 placement
  TW_LAUNCH_CHECK();
  return TW_OK;
}

extern "C" int tw_debug_umma_overhead(int n_blocks, int per_block, int N, int flags, long long* out, void* stream) {
  TW_CHECK_ARG(out && n_blocks > 0 && per_block >= 0 && N >= 16 && N <= 256 && N % 16 == 0, "bad args");
  const int smem = 65 * 1024 + 1024;
  TW_CUDA(cudaFuncSetAttribute(k_umma_overhead, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_overhead<<<1, 128, smem, (cudaStream_t)stream>>>(n_blocks, per_block, N, flags, out);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

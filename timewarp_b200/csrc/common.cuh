// Shared helpers for the timewarp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

#include "../../include/timewarp_b200.h"

namespace tw {

// thread-local last-error string behind tw_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);
void count_launch();

// Optional per-kernel-class device timing (tw_prof_* in the ABI): CUDA events recorded on the
// launching stream around every launch of one kernel class.
enum ProfClass { PROF_NONE = 0, PROF_FFN = 1, PROF_ATTN = 2, PROF_MLP = 3, PROF_ENERGY = 4 };
struct ProfScope {
  cudaStream_t st;
  int slot;
  ProfScope(int cls, cudaStream_t s);
  ~ProfScope();
};

#define TW_CHECK_ARG(cond, ...)                               \
  do {                                                        \
    if (!(cond)) return ::tw::fail(TW_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define TW_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::tw::fail(TW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TW_LAUNCH_CHECK()                                                                      \
  do {                                                                                         \
    ::tw::count_launch();                                                                      \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::tw::fail(TW_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TW_TRY(call)          \
  do {                        \
    int _s = (call);          \
    if (_s != TW_OK) return _s; \
  } while (0)

// Function attributes (the opt-in to > 48 KB of dynamic shared memory) are per device: one flag per device ordinal, so
// that a process driving several GPUs sets them on each.  The attribute calls are idempotent, so two host threads
// racing through the first call are harmless; the flag is published only after the attributes are set.
struct DeviceOnce {
  std::atomic<uint64_t> bits[4] = {};
  static int dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 255;
  }
  bool done() const {
    const int d = dev();
    return (bits[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1u;
  }
  void mark() {
    const int d = dev();
    bits[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release);
  }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = (T*)(base ? base + off : nullptr);
    off += count * sizeof(T);
    return p;
  }
  bool ok() const { return off <= cap; }
};

// Basis function of the kernel attention weights for a scaled distance a = d / l  (kernel_attention.py:9-66):
// Gaussian exp(-a^2), or (coef != nullptr) the Chebyshev-rational expansion sum_c coef[c] R_c(a^2),
// R_n(y) = T_n((y - 1) / (y + 1)) by the three-term recursion; `mean` = per-head coefficient mean when the
// expansion is forced to vanish at infinity.
__device__ __forceinline__ float attention_basis(float a, const float* __restrict__ coef, int order, float mean) {
  if (coef == nullptr) return expf(-(a * a));
  const float y = a * a;
  const float rf = (y - 1.0f) / (y + 1.0f);
  float rprev = 1.0f, rcur = rf;
  float acc = (coef[0] - mean);
  if (order >= 2) acc = fmaf(coef[1] - mean, rcur, acc);
  for (int c = 2; c < order; c++) {
    const float rnext = 2.0f * rf * rcur - rprev;
    acc = fmaf(coef[c] - mean, rnext, acc);
    rprev = rcur, rcur = rnext;
  }
  return acc;
}
__device__ __forceinline__ float cheb_mean(const float* __restrict__ coef, int order, int force_zero) {
  if (coef == nullptr || !force_zero) return 0.f;
  float m = 0.f;
  for (int c = 0; c < order; c++) m += coef[c];
  return m / (float)order;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a block; result valid in every thread.  `red` must hold >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace tw

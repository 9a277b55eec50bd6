// Adam over ONE flat parameter buffer: the optimizer step of the reference's trainer (utilities/training_utils.py:356-368:
// torch.optim.Adam(lr, weight_decay), L2 decay added to the gradient) as a single grid-stride launch.  The backward
// already hands out every gradient as a slice of one flat buffer (flow.py); with the parameters and both moment buffers
// laid out the same way the step is a pure 28-bytes-per-element stream instead of ~50 multi-tensor launches over 659 tensors.
#include "common.cuh"

namespace tw {

// hyper (device, so that a captured CUDA graph sees later changes): {lr, beta1, beta2, eps, weight_decay, step}
__global__ void __launch_bounds__(256) k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n4, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5];
  // torch/optim/adam.py (_single_tensor_adam): step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2);
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = __ldg(g4 + i), mm = m4[i], vv = v4[i];
#define TW_ADAM1(c)                                             \
  {                                                             \
    const float gr = fmaf(wd, pp.c, gg.c);                      \
    mm.c = fmaf(1.f - b1, gr - mm.c, mm.c);                     \
    vv.c = fmaf(1.f - b2, gr * gr, b2 * vv.c);                  \
    pp.c -= step_size * (mm.c / (sqrtf(vv.c) * inv_bc2_sqrt + eps)); \
  }
    TW_ADAM1(x) TW_ADAM1(y) TW_ADAM1(z) TW_ADAM1(w)
#undef TW_ADAM1
    p4[i] = pp, m4[i] = mm, v4[i] = vv;
  }
}

}  // namespace tw

using namespace tw;

extern "C" {

int tw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper, void* stream) {
  TW_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && hyper, "NULL pointer");
  TW_CHECK_ARG(n >= 0 && (n & 3) == 0, "the flat buffers hold a multiple of 4 floats");
  TW_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "the flat buffers must be 16-byte aligned");
  if (n == 0) return TW_OK;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_adam_flat<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n4, hyper);
  TW_LAUNCH_CHECK();
  return TW_OK;
}

}  // extern "C"

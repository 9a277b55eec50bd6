// Tensor-core (tcgen05 / TMEM / bulk-TMA) path of the conditioner networks for the flagship
// layer sizes (d_model 128, one MLP hidden layer of 256, dim_feedforward % 128 == 0).
#pragma once
#include "flow_simt.cuh"

namespace tw {

struct TcScratch {
  void* packed;     // packed bf16 weight images of the layer in flight
  void* scores_op;  // attention-score operand images
  size_t packed_bytes;
};

bool tc_supported(const tw_flow_config* c);
void tc_carve(const tw_flow_config* c, int64_t n, int64_t n_cond, int64_t V, Arena& ar, TcScratch* out);
int tc_begin_pass(const tw_flow_config* c, const ParamView& pv, TcScratch& tc, const float* scores, const uint8_t* mask,
                  int64_t n, int64_t n_cond, int V, cudaStream_t st);
int tc_conditioner(const tw_flow_config* c, const ParamView& pv, int k, TcScratch& tc, const int64_t* atom_types,
                   const float* xc, const float* xv, const float* z_other, const float* scores, float* const actA[2],
                   float* const actB[2], float* const st_out[2], int64_t n, int64_t n_cond, int V, cudaStream_t st);

}  // namespace tw

/*
 * timewarp_b200.h -- C ABI of the B200-native Timewarp conditional-sampling hot path.
 *
 * The reference (microsoft/timewarp) is pure Python: its "FFI" for this path is the pair of
 * Python classes ConditionalFlowDensityModel (modules/model_wrappers/flow.py:106-336) and
 * OpenmmPotentialEnergyTorch (utils/openmm/openmm_bridge.py:252-307).  The functions below are
 * what a binding of those classes needs underneath: raw device pointers in, raw device
 * pointers out, one CUDA stream per call, an int status, no allocation, no host sync, no
 * exceptions.  `timewarp_b200/` (Python) binds them with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - all tensors are contiguous row-major; float = IEEE fp32; masks are uint8 (1 = padding,
 *     the reference's `masked_elements`, dataloader.py:403-417); atom types are int64.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - `workspace` is caller-owned device scratch of at least the size returned by the
 *     matching *_workspace_bytes() query; contents are undefined on return.
 *   - return value: TW_OK or a TW_ERR_* code; tw_last_error() gives a message for the
 *     calling thread.  Launch errors are reported; asynchronous faults surface on the stream.
 *   - thread-compatible: concurrent calls are fine if they use different streams/workspaces.
 */
#ifndef TIMEWARP_B200_H
#define TIMEWARP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TW_ABI_VERSION 1

#define TW_OK 0
#define TW_ERR_INVALID 1      /* bad argument / unsupported size */
#define TW_ERR_CUDA 2         /* CUDA runtime error (see tw_last_error) */
#define TW_ERR_WORKSPACE 3    /* workspace too small */
#define TW_ERR_UNSUPPORTED 4  /* configuration not supported by the requested precision path */

#define TW_MAX_MLP_HIDDEN 4
#define TW_MAX_HEADS 16

/* arithmetic of the token-wise GEMMs */
#define TW_PRECISION_FP32 0    /* CUDA-core fp32 FMA, any layer sizes (generic path)              */
#define TW_PRECISION_BF16X3 1  /* tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate   */
#define TW_PRECISION_BF16 2    /* tcgen05 plain bf16 inputs, fp32 accumulate (training configs)   */

/* Sizes of the flow.  Mirrors CustomAttentionTransformerNVPConfig (model_configs.py:61-69) and
 * CustomAttentionEncoderLayerConfig (modules/layers/custom_attention_encoder.py:126-137). */
typedef struct tw_flow_config {
  int32_t atom_embedding_dim;                 /* E                                              */
  int32_t num_mlp_hidden;                     /* len(latent_mlp_hidden_dims), <= TW_MAX_MLP_HIDDEN */
  int32_t mlp_hidden_dims[TW_MAX_MLP_HIDDEN];
  int32_t num_coupling_layers;                /* L, even (model_constructor.py:156-158)         */
  int32_t num_transformer_layers;             /* T encoder layers per scale/shift network       */
  int32_t d_model;                            /* D (= value dim per head)                       */
  int32_t dim_feedforward;                    /* F                                              */
  int32_t num_heads;                          /* H = len(lengthscales)                          */
  int32_t position_layer_index_mod_2;         /* model_constructor.py:169                       */
  int32_t num_atom_types;                     /* rows of the atom embedding (5)                 */
  float layer_norm_eps;                       /* 1e-5                                           */
  int32_t precision;                          /* TW_PRECISION_*                                 */
  int32_t attention_type;                     /* TW_ATTENTION_*: 0 kernel (also learnable_kernel: the caller passes the effective
                                                 lengthscales), 1 local, 2 chebyshev_kernel (kernel_attention.py:255-339) */
  int32_t cheb_order;                         /* chebyshev_kernel: 1..32 coefficients per head                     */
  int32_t force_asymptotic_zero;              /* chebyshev_kernel: subtract the per-head mean of the coefficients  */
  float max_radius;                           /* local: neighbourhood radius in nm (local_self_attention.py:26)    */
} tw_flow_config;
#define TW_ATTENTION_KERNEL 0
#define TW_ATTENTION_LOCAL 1 /* dot-product attention over the atoms within max_radius; per encoder layer the table slots
                                [wv, lengthscales, wo] hold [qkv_proj.weight (H*3D x D), (ignored), output_proj.weight (D x H*D)] */
#define TW_ATTENTION_CHEBYSHEV 2
#define TW_MAX_CHEB_ORDER 32

/* Parameter table: an array of device pointers to fp32 tensors in the reference's own shapes
 * ([out,in] Linear weights), in this order (state_dict keys of SURVEY.md section 2.1):
 *   [0] flow.atom_embedder.weight   [1] coords_prior_log_scale   [2] velocs_prior_log_scale
 *   (chebyshev_kernel only: AFTER everything below, 2*L*T more pointers -- for k, for net, for t:
 *    self_attn.attention.cheb_coeffs [H, cheb_order])
 *   then for k in 0..L-1, for net in (scale_transformer, shift_transformer):
 *     in_mlp._layers.{0,2,..}.{weight,bias}                      2*(num_mlp_hidden+1) pointers
 *     for t in 0..T-1: self_attn.values_proj.weight, self_attn.attention.lengthscales,
 *                      self_attn.attention._out_projection.weight, linear1.weight, linear1.bias,
 *                      linear2.weight, linear2.bias, norm1.weight, norm1.bias, norm2.weight,
 *                      norm2.bias                                 11 pointers
 *     out_mlp._layers.{0,2,..}.{weight,bias}                     2*(num_mlp_hidden+1) pointers
 * tw_flow_num_params() returns the expected length. */
int tw_abi_version(void);
const char* tw_last_error(void);

/* Measurement hooks (bench.py): number of kernels this library has launched so far in the process;
 * CUDA-event timing of one kernel class on the launching stream (1 = fused FFN, 2 = attention
 * block, 3 = in/out MLPs, 4 = energy; 0 disables).  tw_prof_collect synchronises on the recorded
 * events, returns their summed duration and the number of timed scopes, and resets the list. */
long long tw_debug_launch_count(void);
int tw_prof_enable(int kernel_class);
int tw_prof_collect(double* total_ms, long long* scopes);
int tw_flow_num_params(const tw_flow_config* cfg);

/* Scratch needed for one flow pass over n_samples samples of n_atoms (padded) atoms conditioned
 * on n_cond states (n_samples must be a multiple of n_cond). */
int tw_flow_workspace_bytes(const tw_flow_config* cfg, int64_t n_samples, int64_t n_cond, int64_t n_atoms,
                            size_t* bytes);

/* Tensor-core precisions (TW_PRECISION_BF16X3 / BF16) read the weights from bf16 operand images that
 * are re-packed from the fp32 parameters once per weight update: query the size, pack into a caller-
 * owned, 1024-byte-aligned device buffer, and pass that buffer as `packed_weights` to the flow calls
 * (NULL for TW_PRECISION_FP32).  tw_flow_packed_bytes returns 0 for TW_PRECISION_FP32. */
int tw_flow_packed_bytes(const tw_flow_config* cfg, size_t* bytes);
int tw_flow_pack_weights(const tw_flow_config* cfg, const void* const* params, void* packed, size_t packed_bytes,
                         void* stream);

/* compute_kernel_attention_scores (modules/layers/kernel_attention.py:69-121):
 * out[b,h,i,j] = w / (sum_j |w| + 1e-5),  w = mask_j ? 0 : exp(-(|x_i-x_j| / l_h)^2).
 * coords [B,V,3], mask [B,V], lengthscales [H] (device), out [B,H,V,V]. */
int tw_attn_scores(const float* coords, const uint8_t* mask, const float* lengthscales, int64_t B, int64_t V,
                   int32_t H, float* out, void* stream);

/* One coupling layer's conditioner, NVPCouplingLayer._get_scale_and_shift
 * (modules/custom_transformer_nvp.py:44-93): returns scale = exp(s) and shift, both [B,V,3].
 * x_coords are the (already centred) conditioning coordinates. */
int tw_flow_scale_shift(const tw_flow_config* cfg, const void* const* params, int32_t layer_idx,
                        const int64_t* atom_types, const float* x_coords_centred, const float* x_velocs,
                        const float* z_coords, const float* z_velocs, const uint8_t* mask, int64_t B, int64_t V,
                        float* out_scale, float* out_shift, const void* packed_weights, void* workspace,
                        size_t workspace_bytes, void* stream);

/* flags for the two calls below (ConditionalFlowDensityConfig, flow.py:339-347) */
#define TW_FLOW_DISPLACEMENT_TARGET 1   /* use_displacement_as_target=True: the flow models y - x */

/* ConditionalFlowDensityModel.log_likelihood (modules/model_wrappers/flow.py:131-215).
 * Inputs [B,V] / [B,V,3]; out_log_prob [B]; out_z_* (optional, may be NULL) the latents [B,V,3]. */
int tw_flow_log_likelihood(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types,
                           const float* x_coords, const float* x_velocs, const float* y_coords,
                           const float* y_velocs, const uint8_t* mask, int64_t B, int64_t V, int32_t flags,
                           float* out_log_prob, float* out_z_coords, float* out_z_velocs, const void* packed_weights,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ConditionalFlowDensityModel.conditional_sample_with_logp (flow.py:242-336) given the latent
 * draws.  Conditioning tensors have n_cond rows; the S*n_cond flow samples are laid out as the
 * reference's `.repeat(S,1,1)` does (sample n uses conditioning row n % n_cond).
 * z_coords/z_velocs [S*n_cond,V,3] are the prior draws ALREADY scaled by exp(log_scale)
 * (flow.py:274-275).  Outputs y_coords/y_velocs [S*n_cond,V,3], log_prob [S*n_cond] (may be NULL). */
int tw_flow_sample(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types,
                   const float* x_coords, const float* x_velocs, const uint8_t* mask, int64_t n_cond, int64_t V,
                   int64_t S, int32_t flags, const float* z_coords, const float* z_velocs, float* out_y_coords,
                   float* out_y_velocs, float* out_log_prob, const void* packed_weights, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training (NLL loss, losses.py:321-356 -> density_model_base.py:27-42): the `*_backward` twin of
 * tw_flow_log_likelihood.  The reference differentiates log_likelihood with torch autograd; here
 * the forward records a tape (every layer boundary, caller-owned memory) and the backward returns
 * d(sum_b grad_log_prob[b] * log p_b)/d(parameter) for every trainable parameter.  Tensor-core
 * precisions only (TW_ERR_UNSUPPORTED for TW_PRECISION_FP32 / non-flagship layer sizes).
 *
 * tw_flow_train_bytes: sizes of the tape and of the backward workspace (both 1024-byte aligned).
 * tw_flow_log_likelihood_train: same result as tw_flow_log_likelihood, fills `tape`.
 * tw_flow_log_likelihood_backward: `grads` is a table parallel to `params` (same order, same shapes,
 *   fp32); gradients are ACCUMULATED into it (zero it first for a fresh gradient).  Entries of
 *   lengthscales (buffers, kernel_attention.py:169-171) and of frozen prior log-scales may be NULL.
 *   A NON-NULL entry for the lengthscales of coupling layer 0 / scale network / encoder layer 0 (the
 *   ones the pass reads) requests dL/d(lengthscale) [H] there -- learnable_kernel attention
 *   (kernel_attention.py:217-253; the caller applies d/d(log l) = l * d/dl).
 *   Gradients w.r.t. the coordinate inputs are not produced (the NLL loss does not need them). */
int tw_flow_train_bytes(const tw_flow_config* cfg, int64_t B, int64_t V, size_t* tape_bytes, size_t* workspace_bytes);
int tw_flow_log_likelihood_train(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types,
                                 const float* x_coords, const float* x_velocs, const float* y_coords,
                                 const float* y_velocs, const uint8_t* mask, int64_t B, int64_t V, int32_t flags,
                                 float* out_log_prob, const void* packed_weights, void* tape, size_t tape_bytes,
                                 void* stream);
int tw_flow_log_likelihood_backward(const tw_flow_config* cfg, const void* const* params, void* const* grads,
                                    const int64_t* atom_types, const float* x_velocs, const uint8_t* mask, int64_t B,
                                    int64_t V, const float* grad_log_prob, const void* packed_weights, void* tape,
                                    size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream);

/* tw_flow_log_likelihood_backward that also differentiates w.r.t. the INPUTS (AcceptanceLoss, losses.py:274-555: the
 * reverse-move density is conditioned on the proposal).  Optional outputs [B,V,3] (NULL = not wanted; the first two come as
 * a pair): out_grad_xc = d/d(CENTRED conditioning coordinates) through the conditioner inputs and the attention scores (the
 * caller applies the centring  x - mean_unmasked(x)  chain rule and, with TW_FLOW_DISPLACEMENT_TARGET, subtracts
 * out_grad_z0_coords); out_grad_x_velocs; out_grad_z0_coords / out_grad_z0_velocs = d/d(flow input), i.e. d/dy_coords
 * (displacement or absolute) and d/dy_velocs. */
int tw_flow_log_likelihood_backward_inputs(const tw_flow_config* cfg, const void* const* params, void* const* grads,
                                           const int64_t* atom_types, const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V,
                                           const float* grad_log_prob, const void* packed_weights, void* tape, size_t tape_bytes,
                                           void* workspace, size_t workspace_bytes, float* out_grad_xc, float* out_grad_x_velocs,
                                           float* out_grad_z0_coords, float* out_grad_z0_velocs, void* stream);

/* Sampling direction under autograd (conditional_sample_with_logp inside the energy-based losses, losses.py:396-664; the
 * reference differentiates flow.py:242-336 with torch autograd).  tw_flow_sample_train = tw_flow_sample for S = 1 with a tape
 * (same tape / workspace sizes as tw_flow_train_bytes): z_coords / z_velocs [B,V,3] are the scaled latent draws;
 * out_delta [B] = sum of the log-scales, i.e. log p(y|x) = prior(z) + out_delta (the caller adds the prior term, which
 * also carries the gradient of the prior log-scales).  tw_flow_sample_backward: given d/dy_coords, d/dy_velocs and
 * d/d(delta) it ACCUMULATES the parameter gradients into `grads` (layout as above) and returns d/dz_coords, d/dz_velocs. */
int tw_flow_sample_train(const tw_flow_config* cfg, const void* const* params, const int64_t* atom_types, const float* x_coords,
                         const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V, int32_t flags, const float* z_coords,
                         const float* z_velocs, float* out_y_coords, float* out_y_velocs, float* out_delta,
                         const void* packed_weights, void* tape, size_t tape_bytes, void* stream);
int tw_flow_sample_backward(const tw_flow_config* cfg, const void* const* params, void* const* grads, const int64_t* atom_types,
                            const float* x_velocs, const uint8_t* mask, int64_t B, int64_t V, const float* grad_y_coords,
                            const float* grad_y_velocs, const float* grad_delta, const void* packed_weights, void* tape,
                            size_t tape_bytes, void* workspace, size_t workspace_bytes, float* out_grad_z_coords,
                            float* out_grad_z_velocs, void* stream);

/* Debug: one generic operand-image GEMM of the training path on fp32 row-major inputs.
 * mode 0: C[ar,br] = A B^T; 1: C[ar,bc] = A B; 2: C[ar,128] = sum_h A[:,h*128:(h+1)*128] B[:,h*128:(h+1)*128]
 * (B is [128, H*128]); 3: C[ac,bc] += A^T B (atomic accumulation, `splits` CTAs per output tile). */
int tw_debug_gemm(int mode, int precision, const float* A, int a_rows, int a_cols, const float* B, int b_rows,
                  int b_cols, float* C, int bn, int splits, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Potential energy.  Replaces OpenmmPotentialEnergyTorch.forward (utils/openmm/openmm_bridge.py:
 * 281-294 -> OpenMMBridge.evaluate :170-249 -> OpenMM 7.7 Context.getState) for the implicit-
 * solvent systems built by simulation/md.py:128-173: HarmonicBond + HarmonicAngle +
 * PeriodicTorsion + NonbondedForce(CutoffNonPeriodic) + GBSAOBCForce.  An openmm.System cannot
 * exist here, so the system is passed as arrays (all device pointers, fp32 unless noted). */
typedef struct tw_energy_system {
  int32_t n_atoms;
  int32_t n_bonds;      const int32_t* bond_idx;    /* [n_bonds,2]  */ const float* bond_param;    /* [n_bonds,2]  r0 (nm), k (kJ/mol/nm^2) */
  int32_t n_angles;     const int32_t* angle_idx;   /* [n_angles,3] */ const float* angle_param;   /* [n_angles,2] theta0 (rad), k (kJ/mol/rad^2) */
  int32_t n_torsions;   const int32_t* torsion_idx; /* [n_torsions,4] */ const float* torsion_param; /* [n_torsions,3] periodicity, phase (rad), k (kJ/mol) */
  const float* charge;        /* [n_atoms] e                                   */
  const float* sigma;         /* [n_atoms] nm                                  */
  const float* epsilon;       /* [n_atoms] kJ/mol                              */
  const uint8_t* excluded;    /* [n_atoms,n_atoms] 1 = pair excluded from the direct nonbonded sum (1-2, 1-3, 1-4) */
  int32_t n_exceptions; const int32_t* exception_idx; /* [n_exceptions,2] */ const float* exception_param; /* [n_exceptions,3] chargeProd, sigma, epsilon */
  double cutoff;              /* nm (2.0); <= 0 disables                        */
  double reaction_field_eps;  /* 1.0 when a GB force is present (OpenMM app), else 78.3 */
  double one_4pi_eps0;        /* kJ nm / (mol e^2): 138.935456 (OpenMM 7.7 SimTKOpenMMRealType.h) */
  int32_t use_gb;             /* 0: no implicit solvent                         */
  const float* gb_radius;     /* [n_atoms] nm                                   */
  const float* gb_scale;      /* [n_atoms]                                      */
  double gb_alpha, gb_beta, gb_gamma;  /* OBC1 0.8,0,2.909125 ; OBC2 1,0.8,4.85  */
  double gb_offset;           /* dielectric offset 0.009 nm                     */
  double solute_dielectric, solvent_dielectric; /* 1.0, 78.5                    */
  double surface_area_energy; /* kJ/mol/nm^2 (28.3919551); 0 disables the ACE term; probe radius 0.14 nm */
} tw_energy_system;

/* coords [B,n_atoms,3] nm -> out_energy [B] kJ/mol.  Optional (NULL ok): out_forces [B,n_atoms,3]
 * kJ/mol/nm (= -dU/dx), out_terms [B,5] = bond, angle, torsion, nonbonded(+exceptions), GB/SA
 * (the decomposition of simulation/md.py:288-356).  Non-finite energies are returned as-is. */
int tw_peptide_energy(const tw_energy_system* sys, const float* coords, int64_t B, float* out_energy,
                      float* out_forces, float* out_terms, void* stream);

/* `sim.step(n_steps)` of openmm_step (utils/evaluation_utils.py:439-464) for the integrators of
 * simulation/md.py:116-123 (constraints=None), B conformations at once, in place on coords / velocs
 * [B,n_atoms,3] (nm, nm/ps); masses [n_atoms] dalton; timestep ps; friction 1/ps; kT kJ/mol.
 *   TW_INTEGRATOR_LANGEVIN         v' = a v + (1-a)/g F/m + sqrt(kT(1-a^2)/m) xi;  x' = x + dt v'   (a = exp(-g dt))
 *   TW_INTEGRATOR_LANGEVIN_MIDDLE  v += dt F/m; x += dt/2 v; v = a v + sqrt(kT(1-a^2)/m) xi; x += dt/2 v
 * xi = noise[n_steps,B,n_atoms,3] (standard normals supplied by the caller) or, noise == NULL,
 * Philox4x32-10(seed, conformation*128 + thread) starting at `offset`. */
#define TW_INTEGRATOR_LANGEVIN 0
#define TW_INTEGRATOR_LANGEVIN_MIDDLE 1
int tw_langevin_steps(const tw_energy_system* sys, float* coords, float* velocs, const float* masses, int64_t B,
                      int32_t n_steps, int32_t integrator, double timestep, double friction, double kT,
                      const float* noise, uint64_t seed, uint64_t offset, void* stream);

/* compute_chirality_sign + check_symmetry_change (utils/chirality.py:41-80): centers [C,4] int64
 * (centre, 3 neighbours).  out_signs [B,C] fp32 (optional) = sign of the triple product;
 * out_changed [B] uint8 (optional, needs ref_signs [C] fp32) = any(sign != ref_sign). */
int tw_chirality(const float* coords, const int64_t* centers, const float* ref_signs, int64_t B, int64_t V, int32_t C,
                 uint8_t* out_changed, float* out_signs, void* stream);

/* compute_kinetic_energy (utils/evaluation_utils.py:416-436): velocs [B,V,3]; masses [V] or NULL
 * with inv_kbT ignored (random_velocs mode: 0.5*sum v^2). */
int tw_kinetic_energy(const float* velocs, const float* masses, float inv_kbT, int64_t B, int64_t V, float* out,
                      void* stream);

/* Metropolis-Hastings decision of sample_with_model (utils/evaluation_utils.py:663-689), one
 * decision per row, no host sync:
 *   exponent = (e_pot_y - e_pot_x) + (e_kin_y - e_kin_x) + p_xy - p_yx      (e_* already / kbT)
 *   p_acc    = min(1, exp(-exponent));  accepted = u < p_acc   (NaN => reject)
 * If x_coords/x_velocs are non-NULL the accepted rows are overwritten in place with y (independent
 * chains mode); out_first_accept (optional, int32[1]) receives the first accepted row or -1
 * (the reference's S-proposals-from-one-state mode).  Outputs exponent/p_acc/accepted: [n]. */
int tw_mh_accept(const float* e_pot_x, const float* e_pot_y, const float* e_kin_x, const float* e_kin_y,
                 const float* p_xy, const float* p_yx, const float* u, int64_t n, int64_t V, float* x_coords,
                 float* x_velocs, const float* y_coords, const float* y_velocs, float* out_exponent,
                 float* out_p_acc, uint8_t* out_accepted, int32_t* out_first_accept, void* stream);

/* Energy-threshold acceptance of exploration.py:243-246: keep the old state where
 * e_new - e_old > threshold, else take the new one.  Updates x_coords [n,V,3] and e_old [n] in place. */
int tw_threshold_accept(float* x_coords, float* e_old, const float* y_coords, const float* e_new, float threshold,
                        int64_t n, int64_t V, uint8_t* out_accepted, void* stream);

/* One Adam step over flat buffers (the reference's optimizer, utilities/training_utils.py:356-368:
 * torch.optim.Adam(lr, weight_decay) -- L2 decay added to the gradient, bias-corrected moments):
 * params / grads / exp_avg / exp_avg_sq [n] fp32, 16-byte aligned, n a multiple of 4; hyper (DEVICE memory, so a
 * captured CUDA graph follows a learning-rate schedule) = {lr, beta1, beta2, eps, weight_decay, step (1-based)}. */
int tw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper,
                 void* stream);

/* Debug: single-CTA tcgen05 probe, out[128,N] = bf16(A[128,K]) * bf16(B[N,K])^T with selectable operand
 * placement (a_mode 0 smem K-major SW128 / 1 K-major no swizzle / 2 MN-major SW128 / 3 TMEM; b_mode 0..2),
 * accumulator at TMEM column d_col.  status[0] = 1 if the MMA never completed.  Used by the GPU tests to
 * pin the descriptor conventions of the production kernels on the real chip. */
int tw_debug_umma_probe(const float* A, const float* B, float* out, int N, int K, int a_mode, int b_mode, int d_col,
                        int a_col, int* status, void* stream);

/* Debug: issue/complete cycle counts of n_mma back-to-back M128xNxK16 MMAs (out[0] issue, out[1] done). */
int tw_debug_umma_timing(int n_mma, int N, int ts, int b_noswizzle, long long* out, void* stream);
/* Debug: device buffer of 3*1024*2 + 8 int64 that CTA (0,0) of every following fused-FFN launch fills with
 * {event | item << 8, clock64} records per warp role (0 MMA issuer, 1/2 epilogue groups); NULL disables. */
int tw_debug_set_ffn_trace(long long* device_buf);
/* Same for kernel_class 1 = fused FFN, 2 = attention mixing kernel, 3 = fused attention layer. */
int tw_debug_set_trace(int kernel_class, long long* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* TIMEWARP_B200_H */

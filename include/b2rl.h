/* b2rl.h -- C ABI of libb2rl.so: the sm_100a hot path behind the DeepRL (ShangtongZhang/DeepRL) API.
 *
 * The reference has NO FFI / plugin layer: its seams are duck-typed Python factories on Config
 * (SURVEY.md 8b).  This header is the boundary a maintainer binds with ctypes
 * from the reference-side classes named beside each entry point.  Citations are file:line under
 * /root/reference/deep_rl.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch caching
 *     allocator in our host mirror) owns every buffer; the library allocates nothing on the hot path
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns at once;
 *     all calls are CUDA-graph capturable (no host sync, no allocation)
 *   - return value: 0 on success, negative b2rl_status on error; b2rl_last_error() gives the message
 *     (thread-local).  Nothing throws across the boundary.
 *   - "ring state" is a device int64[8]: [0]=pos [1]=size [2]=capacity [3]=tree write cursor
 *     [4]=philox counter [5..7] reserved; device-resident so that feed/sample/update can be captured
 *     in one CUDA graph
 */
#ifndef B2RL_H
#define B2RL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2RL_OK = 0,
  B2RL_ERR_ARG = -1,      /* bad argument (null pointer, size out of range, unsupported shape) */
  B2RL_ERR_CUDA = -2,     /* a CUDA runtime call / kernel launch failed */
  B2RL_ERR_UNSUPPORTED = -3
} b2rl_status;

typedef enum { B2RL_U8 = 0, B2RL_F16 = 1, B2RL_BF16 = 2, B2RL_F32 = 3 } b2rl_dtype;

int b2rl_version(void);
const char* b2rl_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench.py "gpu_launches") */
int64_t b2rl_launch_count(void);
void b2rl_reset_launch_count(void);
/* Programmatic dependent launch of the per-update kernels (each kernel's prologue overlaps the tail of the one before it;
 * csrc/common.cuh).  On by default; 0 launches every kernel with plain stream order (also: environment B2RL_PDL=0). */
void b2rl_set_pdl(int32_t on);

/* ---------------------------------------------------------------------------------------------
 * Replay ring -- UniformReplay.feed / valid_index / construct_transition / sample
 * (component/replay.py:75-90, 105-110, 112-140, 92-103)
 * frames: uint8 [capacity][row_bytes]  (row_bytes = 84*84 for Atari frames, 4*state_dim for f32 features)
 * action: int32 [capacity]; reward: float64 [capacity] (the reference keeps python floats); mask: int32
 * ------------------------------------------------------------------------------------------- */

/* feed n items (staged on the device in new_*) at the ring cursor; updates ring_state[0..1].
 * reference_quirk != 0 reproduces replay.py:87 (a multi-item feed into a FULL ring writes every item to
 * the slot the call started at); 0 writes each item to its own slot. */
int b2rl_replay_feed(uint8_t* frames, int32_t* action, double* reward, int32_t* mask, int64_t* ring_state,
                     int64_t row_bytes, const uint8_t* new_frames, const int32_t* new_action,
                     const double* new_reward, const int32_t* new_mask, int32_t n, int32_t reference_quirk,
                     void* stream);

/* choose B valid ring indices in candidate-stream order (replay.py:96-100).  candidates: int64 [n_cand] drawn
 * like np.random.randint(0,size) (parity mode) or NULL -> Philox4x32-10 (seed, ring_state[4]) draws n_cand
 * candidates on the device and advances the counter.  idx_out: int64 [B].  status_out: int32 [2] =
 * {accepted (== B on success, < B if the stream ran dry), candidates consumed}. */
int b2rl_replay_select_uniform(int64_t* ring_state, const int64_t* candidates, int32_t n_cand, uint64_t seed,
                               int32_t history, int32_t n_step, int32_t B, int64_t* idx_out, int32_t* status_out,
                               void* stream);

/* construct_transition for B indices (replay.py:112-140): frame-stack gather + n-step return.
 * out_dtype B2RL_U8: raw stacks [B][history][row_bytes] (lut must be NULL, layout 0).
 * converted dtypes (F16/BF16/F32): value = lut[v] (float32 [256] table, e.g. float32(float64(v)/255) = the reference's
 * ImageNormalizer + tensor() rounding) or, with lut == NULL, the integer v itself (exact; the consumer folds the scale
 * into its weights).  layout 0 = [B][history][row_bytes] (NCHW), 1 = [B][row_bytes][history] (NHWC),
 * 2 = space-to-depth by 4 over frames of width frame_w: [B][H/4][W/4][history*16], channel = f*16 + dy*4 + dx.
 * action_out int64 [B]; reward_out float32 [B] (float64 n-step sum rounded once, as tensor() does,
 * utils/torch_utils.py:23); mask_out float32 [B].  Any *_out may be NULL to skip it. */
int b2rl_replay_gather(const uint8_t* frames, const int32_t* action, const double* reward, const int32_t* mask,
                       int64_t capacity, int64_t row_bytes, const int64_t* idx, int32_t B, int32_t history,
                       int32_t n_step, double discount, const float* lut, int32_t out_dtype, int32_t layout,
                       int32_t frame_w, void* state_out, void* next_out, int64_t* action_out, float* reward_out,
                       float* mask_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sum tree in HBM -- SumTree (utils/sum_tree.py:6-67) + PrioritizedReplay (component/replay.py:152-196)
 * tree: float64 [2*capacity-1] array heap, leaves at [capacity-1, 2*capacity-2]; pending: uint8 [capacity]
 * (the reference's pending_idx set, indexed by data index); max_priority: float64 [1] on the device.
 * All arithmetic is sequential-order float64, bit-identical to the reference.
 * ------------------------------------------------------------------------------------------- */

/* n x SumTree.add(max_priority) at the tree write cursor ring_state[3] (replay.py:160-162, sum_tree.py:39-51). */
int b2rl_sumtree_add(double* tree, uint8_t* pending, int64_t capacity, int64_t* ring_state, const double* max_priority,
                     int32_t n, double* scratch /* 16*n bytes */, void* stream);

/* PrioritizedReplay.sample index part (replay.py:167-186): stratified descents with s_i = seg*i +
 * (seg*(i+1) - seg*i) * u_i (CPython random.uniform), sum_tree.py:23-33 descent rule, validity filter in batch
 * order, back-fill.  uniforms: float64 [B] in [0,1) or NULL (Philox, 53-bit).  fills: int64 [B] positions for
 * the back-fill draws (used modulo the current list length) or NULL (Philox).
 * Outputs: tree_idx int64 [B], data_idx int64 [B], sampling_prob float64 [B] (= p / total), status int32[2] =
 * {valid before back-fill, 0}. */
int b2rl_sumtree_sample(const double* tree, uint8_t* pending, int64_t capacity, int64_t* ring_state,
                        const double* uniforms, const int64_t* fills, uint64_t seed, int32_t history, int32_t n_step,
                        int32_t B, int64_t* tree_idx_out, int64_t* data_idx_out, double* sampling_prob_out,
                        int32_t* status_out, void* stream);

/* SumTree.get(s) for B explicit prefix values (sum_tree.py:63-67): tree_idx_out int64 [B], priority_out float64 [B];
 * marks the leaves pending. */
int b2rl_sumtree_get(const double* tree, uint8_t* pending, int64_t capacity, const double* prefix, int32_t B,
                     int64_t* tree_idx_out, double* priority_out, void* stream);

/* PrioritizedReplay.update_priorities (replay.py:193-196): for i in batch order: max_priority = max(.., p_i);
 * SumTree.update(tree_idx_i, p_i) honouring the pending guard (sum_tree.py:54-60).  priority: float32 [B]
 * (to_np of an fp32 tensor, DQN_agent.py:121-123).  scratch: 16*B bytes. */
int b2rl_sumtree_update(double* tree, uint8_t* pending, int64_t capacity, const int64_t* tree_idx,
                        const float* priority, int32_t B, double* max_priority, void* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused target/loss kernels.  Each computes the reference's per-sample loss tensor, the reduced scalar and
 * dLoss/d(network output) in ONE launch.  is_prob (float32 [B], tensor(sampling_prob)) != NULL switches on
 * the PER block of DQN_agent.py:120-127: priority = (|loss|+eps)^alpha ; w = (B*P+1e-6)^-beta / max ;
 * loss *= w before reduce_loss.  priority_out float32 [B] may be NULL when is_prob is NULL.
 * ------------------------------------------------------------------------------------------- */

/* DQNAgent.compute_loss + reduce_loss (DQN_agent.py:78-99): delta = r + gamma_n*q_next*mask - q[a];
 * loss = mean(0.5*(delta*w)^2).  q_next_online != NULL selects double-Q.  dq_out [B][A] receives dLoss/dq. */
int b2rl_dqn_loss(const float* q, const float* q_next_target, const float* q_next_online, const int64_t* action,
                  const float* reward, const float* mask, float gamma_n, int32_t B, int32_t A,
                  const float* is_prob, float beta, float eps, float alpha,
                  float* delta_out, float* priority_out, float* loss_out, float* dq_out,
                  const float* beta_dev /* optional device scalar overriding beta (CUDA-graph replays) */, void* stream);

/* CategoricalDQNAgent.compute_loss + reduce_loss (CategoricalDQN_agent.py:60-89).  log_prob [B][A][N] (online, s),
 * prob_next_target / prob_next_online [B][A][N] (online NULL -> not double).  kl_out [B], loss_out [1] = mean,
 * dlogp_out [B][A][N] = dLoss/dlog_prob.  target_prob_out [B][N] optional (NULL to skip). */
int b2rl_c51_loss(const float* log_prob, const float* prob_next_target, const float* prob_next_online,
                  const int64_t* action, const float* reward, const float* mask, float gamma_n, float v_min,
                  float v_max, int32_t B, int32_t A, int32_t N, const float* is_prob, float beta, float eps,
                  float alpha, float* kl_out, float* priority_out, float* loss_out, float* dlogp_out,
                  float* target_prob_out, int32_t* counter /* int32 scratch [1], zero on first use */,
                  const float* beta_dev /* optional device scalar overriding beta */, void* stream);

/* QuantileRegressionDQNAgent.compute_loss + reduce_loss (QuantileRegressionDQN_agent.py:55-77).
 * quantile / quantile_next [B][A][N].  vec_out [N] = the reference's per-TARGET-quantile vector, loss_out = its
 * mean, dquant_out [B][A][N].  partial: float32 scratch [B][N]; counter: int32 scratch [1], zero on first use. */
int b2rl_qr_loss(const float* quantile, const float* quantile_next, const int64_t* action, const float* reward,
                 const float* mask, float gamma_n, float kappa, int32_t B, int32_t A, int32_t N, float* vec_out,
                 float* loss_out, float* dquant_out, float* partial, int32_t* counter,
                 const float* grad_weight /* float32 [N] = dLoss/dvec / B, or NULL for the mean (1/(B*N)); with
                 partial == NULL only the gradient is computed */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * On-policy: GAE backward recurrence (A2C_agent.py:43-53 == PPO_agent.py:51-61) and the losses.
 * reward, mask: [T][N]; value: [T+1][N]; adv_out, ret_out: [T][N].
 * mode 0 = sequential per env (bit-identical to the reference loop); mode 1 = warp segmented scan
 * (same recurrence re-associated; <= 1e-5 relative).  use_gae == 0 gives adv = ret - v (A2C_agent.py:46-47).
 * ------------------------------------------------------------------------------------------- */
int b2rl_gae(const float* reward, const float* mask, const float* value, float discount, float tau, int32_t T,
             int32_t N, int32_t use_gae, int32_t mode, float* adv_out, float* ret_out, void* stream);

/* advantage normalisation of PPO_agent.py:66: (adv - mean) / std, unbiased std, no epsilon; in place, M elements. */
int b2rl_normalize_advantage(float* adv, int32_t M, void* stream);

/* PPO clipped surrogate for one minibatch of M rows (PPO_agent.py:77-86).  Inputs [M]: log_pi_a (new), entropy,
 * v, old_log_pi_a, advantage, ret.  out[0]=policy_loss out[1]=value_loss out[2]=approx_kl.
 * Gradients: dlogp_out, dent_out = d policy_loss / d(log_pi_a, entropy); dv_out = d value_loss / dv. */
int b2rl_ppo_loss(const float* log_pi_a, const float* entropy, const float* v, const float* old_log_pi_a,
                  const float* advantage, const float* ret, float clip, float entropy_weight, int32_t M, float* out,
                  float* dlogp_out, float* dent_out, float* dv_out, void* stream);

/* A2C objective (A2C_agent.py:55-62): -mean(logp*adv) - ew*mean(ent) + vw*0.5*mean((ret-v)^2) over M rows.
 * out (float32 [4]) = {objective, policy_loss, value_loss, entropy_loss}. */
int b2rl_a2c_loss(const float* log_pi_a, const float* entropy, const float* v, const float* advantage,
                  const float* ret, float entropy_weight, float value_loss_weight, int32_t M, float* out,
                  float* dlogp_out, float* dent_out, float* dv_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-tensor optimizer step with global-norm clip (DQN_agent.py:132-134; examples.py:67-68,139,204):
 * flat float32 views of all parameters / gradients (one contiguous arena, n elements).
 * clip_grad_norm_(max_norm) then RMSprop(centered) or Adam, torch semantics.  norm_scratch: 2048 bytes, zero on first
 * use; after the call float32 norm_scratch[0] = total gradient norm (what clip_grad_norm_ returns).
 * bf16_shadow (optional, uint16 [n]): refreshed copy of the updated parameters in bf16.
 * ------------------------------------------------------------------------------------------- */
int b2rl_clip_rmsprop(float* param, const float* grad, float* square_avg, float* grad_avg, int64_t n,
                      float max_norm, float lr, float alpha, float eps, int32_t centered, float grad_scale,
                      void* norm_scratch, uint16_t* bf16_shadow, void* stream);
int b2rl_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm,
                   float lr, float beta1, float beta2, float eps, int64_t* step_dev, float grad_scale,
                   void* norm_scratch, uint16_t* bf16_shadow, void* stream);
/* b2rl_clip_adam taken only if the DEVICE scalar *gate <= gate_max (nothing is touched otherwise): the KL gate of the PPO
 * actor step, `if approx_kl <= 1.5 * target_kl:` PPO_agent.py:94, decided on the device so the update can be graph-captured */
int b2rl_clip_adam_gated(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float max_norm,
                         float lr, float beta1, float beta2, float eps, int64_t* step_dev, float grad_scale,
                         void* norm_scratch, uint16_t* bf16_shadow, const float* gate, float gate_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense-layer epilogues (network_bodies.py:27-33,70-73: y = relu(layer(x))), bf16 activations [rows][C] (NHWC
 * flattened), fp32 bias.  Forward: y = act(y + bias) in place.  Backward: gx = gy * (y > 0) (gx may alias gy or be
 * NULL) and dbias[c] = sum over rows of gx (block partials + fp32 atomics; dbias is zeroed by the call).  partial /
 * counter: unused (kept for ABI stability), may be NULL.
 * ------------------------------------------------------------------------------------------- */
int b2rl_bias_act_bf16(uint16_t* y, const float* bias, int64_t rows, int32_t C, int32_t relu, void* stream);
/* same epilogue for a split-K GEMM result: y(bf16) = act(x(fp32) + bias) */
int b2rl_bias_act_f32_to_bf16(const float* x, const float* bias, uint16_t* y, int64_t rows, int32_t C, int32_t relu,
                              void* stream);
/* row_map re-lays the gradient out for the next GEMM of the grid-convolution stack (csrc/gemm.cu): 0 identity;
 * 1: compact V x V rows per image -> G x G grid rows; 2: space-to-depth(2) rows (4 groups of C channels per row) ->
 * G x G grid rows.  With a map, `rows` counts the DESTINATION rows (batch*G*G) and gx is [rows][C]; padding rows of the
 * grid are written as zeros. */
int b2rl_act_bwd_bias_grad_bf16(const uint16_t* gy, const uint16_t* y, int64_t rows, int32_t C, int32_t relu,
                                uint16_t* gx, float* dbias, float* partial, int32_t* counter, int32_t row_map, int32_t G,
                                int32_t V, void* stream);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM for the dense contractions (network_bodies.py:27-33, network_heads.py:18-21):
 *   D[M,N] (+)= A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulation in tensor memory (TMEM), TMA-fed.
 * a_mn / b_mn = 0: operand stored row-major [rows][K] (K-major); 1: stored row-major [K][rows] (MN-major) -- weight
 * gradients dW = g^T x read both operands as stored, nothing is transposed in memory.  lda/ldb/ldd: row strides in
 * elements (operands: multiples of 8).  out_mode 0: bf16 store, 1: fp32 store, 2: fp32 atomicAdd into D (required for
 * splits > 1; D must be pre-zeroed or hold the value to accumulate into).  bias (fp32 [N]) and relu are fused into the
 * epilogue.  block_n in {32, 64, 128} = output tile width (tile height is 128).
 * ------------------------------------------------------------------------------------------- */
int b2rl_gemm_bf16(const uint16_t* A, int32_t a_mn, int64_t lda, const uint16_t* B, int32_t b_mn, int64_t ldb, void* D,
                   int64_t ldd, int32_t M, int32_t N, int32_t K, const float* bias, int32_t relu, int32_t out_mode,
                   int32_t splits, int32_t block_n, void* stream);

/* Convolution over a G x G position grid as a shifted-row GEMM, no im2col (csrc/gemm.cu header).
 * mode 0 (forward / dgrad): D[r,:] = sum_taps X[r + shift(tap), :] * W[:, tap*C .. tap*C + C]^T, X bf16 [rows][C] (C % 64
 *   == 0), W bf16 [n_out][taps*C]; shift(tap) = shift_sign * ((tap / taps_x) * grid_w + tap % taps_x); rows off the grid
 *   read zeros.  out_map 1 writes the valid V x V positions of the G grid in space-to-depth(2) layout
 *   ([batch*(V/2)^2][4*n_out], ldd = 4*n_out), out_map 2 compacts them to [batch*V*V][n_out].
 * mode 1 (wgrad): D[n, tap*C + c] += sum_r Gr[r, n] * X[r + shift(tap), c], Gr bf16 [rows][n_out] passed as W_or_G,
 *   D fp32 [n_out][taps*C] (out_mode 2, atomic accumulation over `splits` K slices). */
/* forward / dgrad calls use the slab kernel (one activation slab per tile, resident weights) unless switched off */
void b2rl_set_conv_slab(int32_t on);
int b2rl_conv_gemm_bf16(int32_t mode, const uint16_t* X, int64_t rows, int32_t C, const uint16_t* W_or_G, int32_t n_out,
                        int32_t taps, int32_t taps_x, int32_t grid_w, int32_t shift_sign, void* D, int64_t ldd,
                        const float* bias, int32_t relu, int32_t out_mode, int32_t out_map, int32_t G, int32_t V,
                        int32_t splits, int32_t block_n, void* stream);

/* Dual launches: the same GEMM / convolution for TWO independent operand sets of identical shape (online network on the
 * sampled states, target network on the next states -- DQN_agent.py:84-99 evaluates both every update) in ONE grid;
 * half of the CTAs work on each set, so the per-kernel fixed cost (launch, prologue, weight load, drain) is paid once. */
int b2rl_conv_gemm_dual_bf16(const uint16_t* X, const uint16_t* X2, int64_t rows, int32_t C, const uint16_t* W,
                             const uint16_t* W2, int32_t n_out, int32_t taps, int32_t taps_x, int32_t grid_w,
                             int32_t shift_sign, void* D, void* D2, int64_t ldd, const float* bias, const float* bias2,
                             int32_t relu, int32_t out_mode, int32_t out_map, int32_t G, int32_t V, int32_t block_n,
                             void* stream);
int b2rl_gemm_dual_bf16(const uint16_t* A, const uint16_t* A2, int64_t lda, const uint16_t* B, const uint16_t* B2,
                        int64_t ldb, void* D, void* D2, int64_t ldd, int32_t M, int32_t N, int32_t K, const float* bias,
                        const float* bias2, int32_t relu, int32_t out_mode, int32_t block_n, void* stream);

/* Backward GEMMs with the element-wise backward pass fused into the epilogue (replaces b2rl_act_bwd_bias_grad_bf16 between the
 * dgrad GEMMs of NatureConvBody).  mask: the saved forward activation (bf16) in the GEMM's own output coordinates,
 * [M][mask_ld] -- the ReLU gradient D = mask > 0 ? D : 0; dbias: fp32 [dbias_mod], receives (atomically, zero it first) the
 * column sums of the masked output, index = column % dbias_mod (dbias_mod 0: index = column, dbias is [N]) = the bias gradient
 * of the layer below; sub_c: channels per
 * position for the scatter maps.  out_map 3: rows are space-to-depth(2) positions of a (V/2)^2 grid with 4 x sub_c columns ->
 * rows of the G x G grid; out_map 4: rows are images with V*V x sub_c columns -> rows of the G x G grid.  Grid rows that no
 * tile covers are left untouched (keep the destination zeroed). */
typedef struct {
  const uint16_t* mask;
  int64_t mask_ld;
  float* dbias;
  int32_t dbias_mod;
  int32_t sub_c;
} b2rl_bwd_epilogue;
int b2rl_conv_gemm_bwd_bf16(const uint16_t* G_rows, int64_t rows, int32_t C, const uint16_t* W, int32_t n_out, int32_t taps,
                            int32_t taps_x, int32_t grid_w, void* D, int64_t ldd, int32_t out_map, int32_t G, int32_t V,
                            const b2rl_bwd_epilogue* ext, int32_t block_n, void* stream);
int b2rl_gemm_bwd_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int32_t b_mn, int64_t ldb, void* D, int64_t ldd,
                       int32_t M, int32_t N, int32_t K, int32_t out_map, int32_t G, int32_t V,
                       const b2rl_bwd_epilogue* ext, int32_t block_n, void* stream);

/* NatureConvBody weights (network_bodies.py:13-20) between the reference's parameter layouts and the tap-major bf16
 * operands of the grid-GEMM stack: w1 [32,c1,8,8] -> w1f [32][4 taps][16*c1] (times `scale` = ImageNormalizer's 1/255);
 * w2 [64,32,4,4] -> w2f [64][4][128], w2d [128][4][64]; w3 [64,64,3,3] -> w3f [64][9][64], w3d [64][9][64];
 * w4 [n4, 64*7*7 in (c,h,w) order] -> w4p [n4][(h,w,c)].  unpack maps fp32 gradients in the GEMM layouts back and ADDS
 * them (and the four bias gradients) into the reference-layout .grad buffers. */
int b2rl_nature_pack_weights(const float* w1, const float* w2, const float* w3, const float* w4, int32_t c1, int32_t n4,
                             float scale, uint16_t* w1f, uint16_t* w2f, uint16_t* w2d, uint16_t* w3f, uint16_t* w3d,
                             uint16_t* w4p, void* stream);
int b2rl_nature_unpack_grads(const float* g1f, const float* g2f, const float* g3f, const float* g4p, const float* db1,
                             const float* db2, const float* db3, const float* db4, int32_t c1, int32_t n4, float scale,
                             float* gw1, float* gw2, float* gw3, float* gw4, float* gb1, float* gb2, float* gb3,
                             float* gb4, int32_t p1, int32_t p2, int32_t p3 /* split-K partial counts of g1f/g2f/g3f, stored
                             n_out*K floats apart (b2rl_conv_wgrad_partials); 1 = plain */, void* stream);

/* Narrow value heads (VanillaNet / DuelingNet, network_heads.py:11-37) on the bf16 features phi [B][K] of the fused body:
 * forward q [B][A] (fp32) = phi Wa^T + ba, or with Wv/bv != NULL the dueling combine q = v + adv - mean(adv);
 * backward from gq [B][A]: gphi [B][K] (bf16), and gWa [A][K], gba [A], gWv [K], gbv [1] ACCUMULATED (atomics) into the
 * given fp32 buffers.  0 < A < 32. */
int b2rl_head_fwd(const uint16_t* phi, const float* Wa, const float* ba, const float* Wv, const float* bv, int32_t B,
                  int32_t K, int32_t A, float* q, void* stream);
int b2rl_head_bwd(const float* gq, const uint16_t* phi, const float* Wa, const float* Wv, int32_t B, int32_t K, int32_t A,
                  uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, void* stream);
/* b2rl_head_bwd when phi = relu(layer(.)) (NatureConvBody's fc4 output): gphi is masked (0 where phi <= 0) and
 * relu_colsum[K] (fp32, zero it first) receives the column sums of the masked gphi = that layer's bias gradient. */
int b2rl_head_bwd_relu(const float* gq, const uint16_t* phi, const float* Wa, const float* Wv, int32_t B, int32_t K, int32_t A,
                       uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, float* relu_colsum, void* stream);

/* Convolution weight gradient as split-K partials (no atomics): partial i of *n_partials_host (<= 148, written on the
 * HOST, deterministic for given shapes) is stored at partials + i * n_out*taps*C floats. */
int b2rl_conv_wgrad_partials(const uint16_t* X, int64_t rows, int32_t C, const uint16_t* G, int32_t n_out, int32_t taps,
                             int32_t taps_x, int32_t grid_w, float* partials, int32_t* n_partials_host, void* stream);

/* K1 -- conv1 of NatureConvBody straight from the uint8 replay ring: the fused gather -> normalize -> conv1 of the
 * reference chain replay.py:124-134 (frame-stack gather) -> normalizer.py:58-61 -> network_bodies.py:27, with no
 * materialised batch.  frames: ring [capacity][row_bytes] uint8 (16-byte aligned rows); idx: int64 [batch] sampled ring indices; first: ring row of
 * the oldest stacked frame relative to idx[b] (-(history-1) for the state, n_step-(history-1) for the next state);
 * history must be 4 (4 frames x 16 pixels = the 64 channels of one space-to-depth(4) position).  The pixels enter as
 * exact integers 0..255; ImageNormalizer's 1/255 is folded into W [n_out][4 taps * 64] (b2rl_nature_pack_weights).
 * fwd: D = act(conv1 + bias) over rows (b, gy, gx) of the (frame_w/4)^2 grid, bf16, output row maps of
 * b2rl_conv_gemm_bf16.  wgrad_partials: split-K partials of dW[n][tap*64+c] = sum_r G[r][n] * x[r + shift(tap)][c],
 * contract of b2rl_conv_wgrad_partials. */
int b2rl_conv1_u8_fwd(const uint8_t* frames, int64_t capacity, const int64_t* idx, int32_t first, int64_t row_bytes,
                      int32_t frame_w, int32_t batch, int32_t history, const uint16_t* W, int32_t n_out, void* D, int64_t ldd,
                      const float* bias, int32_t relu, int32_t out_map, int32_t V, void* stream);
int b2rl_conv1_u8_wgrad_partials(const uint8_t* frames, int64_t capacity, const int64_t* idx, int32_t first, int64_t row_bytes,
                                 int32_t frame_w, int32_t batch, int32_t history, const uint16_t* G_rows, int32_t n_out, float* partials,
                                 int32_t* n_partials_host, void* stream);

/* D = act(A B^T + bias) (bf16 out) with split-K and an in-kernel fix-up: one launch instead of zero-fill + atomic split-K +
 * bias/activation pass (fc4 of NatureConvBody at small batch, network_bodies.py:33).  A [M][K], B [N][K] bf16 K-major.
 * ws: fp32 [splits][ceil(M/128)*128][ceil(N/block_n)*block_n]; counters: int32 [tiles], zeroed once (self re-arming).
 * Concurrent launches (different streams) need their own ws / counters. */
int b2rl_gemm_splitk_bf16(const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb, void* D, int64_t ldd, int32_t M,
                          int32_t N, int32_t K, const float* bias, int32_t relu, int32_t splits, int32_t block_n, float* ws,
                          int32_t* counters, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tail of one gradient update for a NatureConvBody network on the tcgen05 path (csrc/tail.cu):
 * loss.backward()'s last step + clip_grad_norm_ + optimizer.step (DQN_agent.py:131-134) in two launches.
 * Work is described by unit tables (int32 x 4 per unit: arena offset, length, kind, row | segment << 16;
 * kinds 0 plain, 1-4 one output row (or 256-element segment of it) of conv1 / conv2 / conv3 / fc4 weights,
 * 5-8 their biases), built once by the host (deeprl_b200/network/tail.py).
 *
 * grad_reduce: sums the split-K partials of the conv weight gradients (g1p/g2p/g3p: [p][n_out][taps*C] fp32),
 *   maps all four weight gradients from the GEMM layouts to the reference layouts and WRITES them into the flat
 *   gradient arena, moves the bias gradients there (db1..db4 are re-zeroed), and leaves the sum of squares of the
 *   gradient elements of unit i in unit_sumsq[i] (plain units: whatever the arena already holds, e.g. the head's
 *   gradients).  step_dev (Adam) is incremented by one if not NULL.  norm_scratch != NULL: the CTA that finishes last adds the
 *   unit partials (fixed order) and leaves clip_grad_norm_'s total norm and coefficient there, as b2rl_grad_norm does.
 * fused_opt: clip coefficient from the unit partials (NULL: from norm_scratch, see b2rl_grad_norm), RMSprop
 *   (opt 0 plain, 1 centered; a_ = alpha) or Adam (opt 2; a_, b_ = betas) exactly as b2rl_clip_rmsprop / b2rl_clip_adam,
 *   gradient re-zeroed when zero_grad != 0, and the updated conv / fc4 weights written to the bf16 tap-major GEMM
 *   operands of b2rl_nature_pack_weights (all six pointers, or all NULL); bf16_shadow != NULL: a bf16 copy of every updated
 *   parameter at the same arena offset (the GEMM operands of the distributional heads).
 * ------------------------------------------------------------------------------------------- */
int b2rl_nature_grad_reduce(const int32_t* units, int32_t n_units, const float* g1p, int32_t p1, const float* g2p, int32_t p2,
                            const float* g3p, int32_t p3, const float* g4p, float* db1, float* db2, float* db3, float* db4,
                            int32_t c1, int32_t n4, float scale, float* grad, float* unit_sumsq, int64_t* step_dev,
                            void* norm_scratch, float max_norm, float grad_scale, void* stream);
int b2rl_nature_fused_opt(const int32_t* units, int32_t n_units, float* param, float* grad, float* s1, float* s2, int32_t opt,
                          float lr, float a_, float b_, float eps, float max_norm, float grad_scale, const float* unit_sumsq,
                          int32_t n_sumsq, void* norm_scratch, const int64_t* step_dev, int32_t c1, int32_t n4, float scale,
                          uint16_t* w1f, uint16_t* w2f, uint16_t* w2d, uint16_t* w3f, uint16_t* w3d, uint16_t* w4p,
                          int32_t zero_grad, uint16_t* bf16_shadow, void* stream);
/* clip_grad_norm_'s coefficient alone (torch.nn.utils.clip_grad_norm_): norm_scratch[0] = ||grad * grad_scale||,
 * norm_scratch[1] = min(max_norm / (norm + 1e-6), 1) * grad_scale. */
int b2rl_grad_norm(const float* grad, int64_t n, float grad_scale, float max_norm, void* norm_scratch, void* stream);

/* b2rl_dqn_head_fused in two launches (row kernel: one warp per batch row; then the head backward): same arguments plus
 * geff = float scratch [B][33]; scratch: int32 counter + 12 bytes padding + float [ceil(B/4)], zero-initialised once. */
int b2rl_dqn_head_two(const uint16_t* phi, const uint16_t* phi_t, const uint16_t* phi_o, const float* Wa, const float* ba,
                      const float* Wv, const float* bv, const float* Wa_t, const float* ba_t, const float* Wv_t,
                      const float* bv_t, const int64_t* action, const float* reward, const float* mask, float gamma_n, int32_t B,
                      int32_t K, int32_t A, const float* is_prob, float beta, const float* beta_dev, float eps, float alpha,
                      uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, float* relu_colsum, float* q_out,
                      float* delta_out, float* prio_out, float* loss_out, float* scratch, float* geff, void* stream);

/* Device-side actor step of the on-policy agents (PPO_agent.py:45-50 between two task.step() calls) in one
 * launch: MeanStdNormalizer (normalizer.py:36-51 over baselines' RunningMeanStd: float64 Chan merge of the batch moments into
 * rm_mean / rm_var / rm_count -- all three NULL: no normalisation; update_stats 0: read-only -- then clip((x - mean) /
 * sqrt(var + eps), +-clip) rounded once to float32), GaussianActorCriticNet.forward (network_heads.py:173-214) with two-layer
 * tanh FCBody actor / critic bodies, and the Normal sample / log_prob / entropy.  obs [N][D]; weights in the reference's
 * nn.Linear layouts ([out][in]); z: supplied standard normals [N][A] (parity mode) or NULL -> Philox4x32-10 (seed, *counter)
 * with Box-Muller; given_action != NULL: log_prob of those actions instead of sampling.  Outputs: state_out [N][D] (normalised,
 * may be NULL), action / mean [N][A] (mean may be NULL), log_pi_a / entropy / v [N].  Limits: N <= 64, D, hidden <= 128, A <= 32. */
int b2rl_gaussian_actor_step(const float* obs, double* rm_mean, double* rm_var, double* rm_count, int32_t update_stats, double clip,
                             double eps, const float* aw1, const float* ab1, const float* aw2, const float* ab2, const float* faw,
                             const float* fab, const float* cw1, const float* cb1, const float* cw2, const float* cb2,
                             const float* fcw, const float* fcb, const float* std_param, int32_t N, int32_t D, int32_t H1,
                             int32_t H2, int32_t A, const float* z, uint64_t seed, int64_t* counter, const float* given_action,
                             float* state_out, float* action, float* log_pi_a, float* entropy, float* mean, float* v,
                             void* stream);

/* Every minibatch update of one PPO iteration (PPO_agent.py:68-99, non-shared representation; optimization_epochs x
 * rows / mini_batch_size updates) in ONE launch of one persistent thread block: per minibatch b the rows perm[b][0..mb) of the
 * rollout (state [R][D], action [R][A], old_log_pi_a / ret / advantage [R], advantage already normalised, PPO_agent.py:66) go
 * through GaussianActorCriticNet (network_heads.py:198-214: DummyBody phi, two-layer tanh FCBody actor / critic bodies,
 * mean = tanh(fc_action), std = softplus(std)), the clipped-surrogate / entropy / value losses (PPO_agent.py:79-88), the
 * backward pass, the actor Adam step iff approx_kl <= kl_gate (= 1.5 * target_kl, PPO_agent.py:94) and the critic Adam step
 * (torch.optim.Adam arithmetic).  a_* / c_*: flat arenas holding the actor's / critic's parameters, exp_avg, exp_avg_sq
 * (float32) and step count (int64 [1], device); a_off int32 [7] = element offsets of actor_body.layers.0.weight, .bias,
 * layers.1.weight, .bias, fc_action.weight, fc_action.bias, std inside the actor arenas (host array); c_off int32 [6] likewise
 * for critic_body.* and fc_critic.*.  stats float32 [4] = policy loss, value loss, approx_kl of the LAST minibatch and the
 * number of actor steps taken.  Limits: D <= 256, A <= 32, hidden <= 128, mini batch <= 128 and a multiple of 4. */
/* Profiling hook of b2rl_ppo_minibatch_updates: install (NULL: remove) a device buffer int64 [2 + 9 * n_batches] that the next
 * launches fill with clock64() of thread 0 after every phase barrier (scripts/ppo_phase_clocks.py). */
int b2rl_ppo_set_phase_clocks(int64_t* clocks);

/* Dynamic shared memory (bytes) b2rl_ppo_minibatch_updates needs for these sizes; it must fit the 227 KB of one SM. */
int64_t b2rl_ppo_minibatch_smem_bytes(int32_t D, int32_t A, int32_t H1, int32_t H2, int32_t mb);

int b2rl_ppo_minibatch_updates(const float* state, const float* action, const float* old_log_pi_a, const float* ret,
                               const float* advantage, int32_t D, int32_t A, int32_t H1, int32_t H2, int32_t mb,
                               const int64_t* perm, int32_t n_batches, float* a_flat, float* a_exp_avg, float* a_exp_avg_sq,
                               int64_t* a_step, const int32_t* a_off, float* c_flat, float* c_exp_avg, float* c_exp_avg_sq,
                               int64_t* c_step, const int32_t* c_off, float a_lr, float a_beta1, float a_beta2, float a_eps,
                               float c_lr, float c_beta1, float c_beta2, float c_eps, float ratio_clip, float entropy_weight,
                               float kl_gate, float* stats, void* stream);

/* Element-wise halves of the distributional heads (CategoricalNet / QuantileNet, network_heads.py:40-55, 89-102) around the
 * tcgen05 GEMMs: softmax + log_softmax over the N atoms of every (b, a) row (either output may be NULL), and the backward
 * preparation dlogits = dout - prob * sum_n dout (prob == NULL: dlogits = dout, QR-DQN) written as the bf16 GEMM operand
 * g [B][ld] (ld >= A*N, multiple of 8; padding zeroed) with its column sums ADDED to dbias [A*N] (the bias gradient). */
int b2rl_dist_softmax(const float* logits, int32_t rows, int32_t N, float* prob, float* log_prob, void* stream);
int b2rl_dist_head_bwd_prep(const float* dout, const float* prob, int32_t B, int32_t A, int32_t N, uint16_t* g, int32_t ld,
                            float* dbias, void* stream);

/* DQN update, head part, in one launch (DQN_agent.py:78-99, 120-127 and the head's backward): q = head(phi) on s,
 * q_next = target_head(phi_t) on s' [argmax from head(phi_o) for double-Q], delta / priorities / IS weights / loss as
 * b2rl_dqn_loss, then the gradients of the head (accumulated into gWa / gba / gWv / gbv), dphi masked by phi > 0 and
 * its column sums (fc4's bias gradient) accumulated into relu_colsum.  Heads: VanillaNet (Wv == NULL) or DuelingNet
 * (network_heads.py:11-37).  scratch: int32 counter (zero-initialised once, self re-arming) + 12 bytes
 * padding + float [ceil(B/16)]. */
int b2rl_dqn_head_fused(const uint16_t* phi, const uint16_t* phi_t, const uint16_t* phi_o, const float* Wa, const float* ba,
                        const float* Wv, const float* bv, const float* Wa_t, const float* ba_t, const float* Wv_t,
                        const float* bv_t, const int64_t* action, const float* reward, const float* mask, float gamma_n,
                        int32_t B, int32_t K, int32_t A, const float* is_prob, float beta, const float* beta_dev, float eps,
                        float alpha, uint16_t* gphi, float* gWa, float* gba, float* gWv, float* gbv, float* relu_colsum,
                        float* q_out, float* delta_out, float* prio_out, float* loss_out, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2RL_H */

/*
 * b200rl.h -- C ABI of libb200rl.so: the PPO data path of rl-algo-impls as hand-written
 * sm_100a CUDA kernels.
 *
 * The reference (rl-algo-impls) is pure Python and has no FFI of its own; each entry point
 * below replaces one eager numpy / torch op sequence of the reference, cited as file:line
 * relative to the reference root.  The reference-side binding (a ctypes stub) is shown in
 * INTEGRATION.md; this repo's own binding is rl_algo_impls_b200/_lib.py.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current device unless the name ends in _host;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises the device and never allocates or frees memory: scratch space is a caller
 *    buffer sized by the matching *_workspace_bytes();
 *  - returns 0 on success, B200RL_EINVAL (-1) for a bad argument, B200RL_EUNSUPPORTED (-2)
 *    for a shape/dtype this build has no kernel for, B200RL_ECUDA (-3) for a CUDA launch
 *    error; b200rl_last_error() returns the thread-local message;
 *  - there is no CPU path: without a CUDA device every compute call fails with B200RL_ECUDA.
 *  - layouts are the reference's flattened rollout layouts: time-major [T, N, ...] for the
 *    rollout buffer (flat sample = t*N + n, rollout/rollout.py:120-121) and row-major
 *    [B, ...] for minibatches.
 */
#ifndef B200RL_H
#define B200RL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RL_VERSION 102 /* major*100 + minor */

#define B200RL_OK 0
#define B200RL_EINVAL (-1)
#define B200RL_EUNSUPPORTED (-2)
#define B200RL_ECUDA (-3)

#define B200RL_MAX_HEADS 16       /* action planes per cell (A) */
#define B200RL_MAX_VALUE_HEADS 32 /* value / reward heads (V) */
#define B200RL_MAX_GATHER 16      /* tensors per gather call */

/* element types of the tensors whose dtype the reference lets vary */
#define B200RL_F32 0
#define B200RL_BF16 1
#define B200RL_U8 2
#define B200RL_I32 3
#define B200RL_I64 4

typedef void* b200rl_stream_t; /* cudaStream_t */

int b200rl_version(void);
const char* b200rl_last_error(void);

/* Page-lock (cudaHostRegister) / release a HOST buffer the caller owns -- a host vec env's observation or mask
 * array (rollout/sync_step_rollout.py:202-212 hands them over every step) -- so that the per-step upload is a
 * DMA straight out of it instead of a copy through a staging buffer.  A refusal returns B200RL_ECUDA and leaves
 * no pending CUDA error behind. */
int b200rl_host_register(void* ptr, size_t bytes);
int b200rl_host_unregister(void* ptr);
/* One env step's uploads in one call: dst[i] (device) <- src[i] (page-locked host), bytes[i] each, asynchronous on
 * `stream`, in order.  Replaces the per-field obs / mask transfers of rollout/sync_step_rollout.py:181-216 through
 * policy.step (shared/policy/actor_critic.py:306-318). */
int b200rl_h2d_batch(int n, void* const* dst_host, const void* const* src_host, const int64_t* bytes_host,
                     b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K1  GAE(lambda) reverse-time scan + returns.
 * Replaces shared/gae.py:97-124 (compute_advantages) and rollout/vec_rollout.py:88 (returns).
 *   rewards, values, advantages, returns : [T, N, V] f32     (V == 1: [T, N])
 *   episode_starts                       : [T, N] u8 (bool)
 *   next_episode_starts                  : [N] u8,  next_values : [N, V] f32
 *   gamma_host, gae_lambda_host          : HOST arrays of V doubles
 *   gamma_is_scalar: non-zero when the reference was given a Python-float gamma (a weak
 *     scalar: gamma*next_value is rounded to f32, gae.py:121); zero for a per-head ndarray
 *     gamma (product in f64).  Either way the carry is f64 and the store f32, like numpy.
 * Bit-exact with the reference.  `returns` may be NULL.
 */
int b200rl_gae_scan_f32(const float* rewards, const float* values, const uint8_t* episode_starts,
                        const uint8_t* next_episode_starts, const float* next_values,
                        const double* gamma_host, const double* gae_lambda_host, int gamma_is_scalar,
                        float* advantages, float* returns, int64_t T, int64_t N, int64_t V,
                        b200rl_stream_t stream);

/* K1b  GAE over ragged trajectories (segments concatenated along the first axis, seg_offsets [n+1] on
 * the device).  Replaces compute_advantages per Trajectory (rollout/trajectory.py:56-95) when
 * `episode_starts` [total] is given (entry t = dones[t-1]; next_episode_starts[s] = dones[-1] of
 * segment s), and DiscreteSkipsTrajectoryBuilder.trajectory
 * (rollout/discrete_skips_trajectory_builder.py:84-100: gamma ** steps_elapsed[t] discounting, value 0
 * after a finished trajectory) when `steps_elapsed` [total] i32 is given instead.  Exactly one of the two.
 */
int b200rl_gae_segments_f32(const float* rewards, const float* values, const uint8_t* episode_starts,
                            const int32_t* steps_elapsed, const int64_t* seg_offsets,
                            const uint8_t* next_episode_starts, const float* next_values,
                            const double* gamma_host, const double* gae_lambda_host, int gamma_is_scalar,
                            float* advantages, float* returns, int64_t n_segments, int64_t V,
                            b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K2  per-minibatch advantage moments and normalisation.  Replaces ppo/ppo.py:307-318.
 * adv_mode:
 *   0 none;  1 (A - mean(0)) / (std(0) + 1e-8) per head (unbiased std);  2 A / (std(0) + 1e-8);
 *   3 contract with weights first, then scalar (A - mean) / (std + 1e-8)
 *     (normalize_advantages_after_scaling).
 * In modes 0-2 the [V] result is contracted with `weights_host` when it is non-NULL
 * (multi_reward_weights, ppo.py:317-318); V > 1 without weights is only legal when the
 * consumer wants [B, V] back (b200rl_adv_normalize_f32 with out_v == V).
 *
 * b200rl_adv_moments_f64: moments[0..Vm) = sum, moments[Vm..2Vm) = sum of squares, moments[2Vm] = B,
 *   over rows idx[0..B) of adv[., V] (idx NULL: rows 0..B); Vm = 1 in mode 3 else V.  f64
 *   accumulation, deterministic reduction order.  The caller may all-reduce `moments`
 *   across ranks before normalising (exact global-minibatch statistics).
 */
size_t b200rl_adv_moments_workspace_bytes(int64_t B, int64_t V);
int b200rl_adv_moments_f64(const float* adv, const int64_t* idx, int64_t B, int64_t V, int adv_mode,
                           const float* weights_host, double* moments, void* workspace,
                           size_t workspace_bytes, b200rl_stream_t stream);
int b200rl_adv_normalize_f32(const float* adv, const int64_t* idx, int64_t B, int64_t V, int adv_mode,
                             const float* weights_host, const double* moments, float* out, int64_t out_v,
                             b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K3  minibatch row gather.  Replaces rollout/rollout.py:56-69 (Batch.__getitem__) and
 * shared/tensor_utils.py:66-72: dst[t][b, :] = src[t][idx[b], :] for n_tensors tensors in one
 * launch.  src_host / dst_host / row_bytes_host are HOST arrays of device pointers / sizes.
 * Rows wider than 256 bytes are cut into 16 KB chunks, one CTA each, staged through REGISTERS:
 * 128-bit streaming loads (L1 no-allocate), every load of a chunk issued before its first store
 * (16-byte aligned rows; 4-byte and byte paths otherwise).  Narrow rows go one per thread in a
 * second grid that overlaps the first and completes after it, so whatever follows on the stream
 * sees every gathered row.  idx is int64 [B] on the device; rows with idx outside [0, n_src_rows)
 * are left untouched.
 */
int b200rl_gather_rows(const void* const* src_host, void* const* dst_host, const int64_t* row_bytes_host,
                       int n_tensors, const int64_t* idx, int64_t B, int64_t n_src_rows,
                       b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * PPO per-sample loss terms, shared by every K4 variant.  Replaces ppo/ppo.py:326-361,373-374
 * and the stats of :379-409.
 *   old_logp [B]; adv [B, adv_v] raw advantages; moments from K2 (NULL when adv_mode == 0);
 *   adv_weights_host [adv_v] or NULL; old_values, returns, new_values, dvalues: [B, V] f32.
 *   clip_range_vf < 0 disables value clipping.  vf_coef_host [V].  loss_scale multiplies the
 *   total loss and every gradient (1 / num_minibatches under gradient accumulation).
 *   kl_cutoff < 0 disables the cut-off; pi_coef_state is a device float (1 or 0) that the
 *   scalar kernel reads, zeroes when approx_kl > kl_cutoff and leaves sticky (ppo.py:279,354).
 * stats_out (device, f32): [0] loss [1] pi_loss [2] entropy_loss [3] approx_kl [4] clipped_frac
 *   [5 .. 5+V) v_loss per head (after halving) [5+V .. 5+2V) val_clipped_frac [5+2V] teacher_kl_loss.
 * Teacher-KL term (loss/teacher_kl_loss.py:35-50, ppo/ppo.py:363-371): when teacher_logp is non-NULL,
 *   loss += teacher_kl_coef * mean(w * f(teacher_logp - new_logp)), f(d) = (e^d - 1) - d (unbiased) or
 *   d^2 / 2, w = the PPO ratio (NOT detached, as the reference) when importance sampling is on.
 */
typedef struct b200rl_ppo_args {
  const float* old_logp;
  const float* adv;
  const double* moments;
  const float* adv_weights_host;
  int64_t adv_v;
  int adv_mode;
  const float* old_values;
  const float* returns;
  const float* new_values;
  float* dvalues;
  int64_t V;
  double clip_range;    /* Python float: 1 - clip_range is formed in double, then rounded to f32 */
  double clip_range_vf; /* < 0: value clipping off */
  const float* vf_coef_host;
  float ent_coef;
  float pi_coef;
  int vf_halving;
  float loss_scale;
  float* stats_out;
  const float* teacher_logp; /* [B] or NULL */
  float teacher_kl_coef;
  int teacher_unbiased;   /* loss/teacher_kl_loss.py:14 `unbiased` */
  int teacher_importance; /* ppo.py:139 teacher_loss_importance_sampling */
  int vf_loss;            /* ppo.py:186 vf_loss_fn = getattr(F, name): B200RL_VF_* (0 = mse_loss, the default) */
} b200rl_ppo_args;

/* element-wise value losses of torch.nn.functional with their default delta / beta = 1 */
#define B200RL_VF_MSE 0       /* (x - y)^2 */
#define B200RL_VF_HUBER 1     /* |d| < 1 ? d^2 / 2 : |d| - 1/2 */
#define B200RL_VF_SMOOTH_L1 2 /* the same function at beta = 1 */
#define B200RL_VF_L1 3        /* |d| */

#define B200RL_PPO_NSTATS(V) (6 + 2 * (V))

/* Per-sample stage alone (distribution-level path): given new_logp [B] and entropy
 * [B, ent_d] produced by any differentiable head, writes dlogp [B], dentropy [B, ent_d],
 * dvalues and stats.  Honours kl_cutoff without a host sync (two-phase inside one launch). */
size_t b200rl_ppo_workspace_bytes(int64_t B, int64_t V);
int b200rl_ppo_scalar_loss_f32(const float* new_logp, const float* entropy, int64_t ent_d, int64_t B,
                               const b200rl_ppo_args* args, float kl_cutoff, float* pi_coef_state,
                               float* dlogp, float* dentropy, void* workspace, size_t workspace_bytes,
                               b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K4a  (Masked)Categorical over the last dim.  Replaces shared/actor/categorical.py:12-54 and
 * torch.distributions.Categorical log_prob / entropy and their autograd backward.
 *   logits [R, n] f32; mask [R, n] u8 or NULL; actions [R] (act_dtype U8/I32/I64).
 * fwd: logp [R], entropy [R].   bwd: dlogits [R, n] from dlogp [R], dentropy [R].
 * fused: PPO loss forward + backward in one launch (R == B): dlogits, dvalues, stats.
 */
int b200rl_categorical_fwd_f32(const float* logits, const uint8_t* mask, const void* actions, int act_dtype,
                               int64_t R, int64_t n, float* logp, float* entropy, b200rl_stream_t stream);
int b200rl_categorical_bwd_f32(const float* logits, const uint8_t* mask, const void* actions, int act_dtype,
                               int64_t R, int64_t n, const float* dlogp, const float* dentropy,
                               float* dlogits, b200rl_stream_t stream);
int b200rl_ppo_categorical_loss_f32(const float* logits, const uint8_t* mask, const void* actions,
                                    int act_dtype, int64_t B, int64_t n, const b200rl_ppo_args* args,
                                    float* dlogits, void* workspace, size_t workspace_bytes,
                                    b200rl_stream_t stream);

/* K4b  diagonal Gaussian.  Replaces shared/actor/gaussian.py:11-16,42-45 (+ torch Normal).
 *   mu [B, D], log_std [D], actions [B, D] f32.  logp summed over D; entropy [B, D] (not
 *   summed, as the reference), so entropy_loss averages over B*D.
 * fused: dmu [B, D], dlog_std [D], dvalues, stats. */
int b200rl_gaussian_fwd_f32(const float* mu, const float* log_std, const float* actions, int64_t B, int64_t D,
                            float* logp, float* entropy, b200rl_stream_t stream);
int b200rl_ppo_gaussian_loss_f32(const float* mu, const float* log_std, const float* actions, int64_t B,
                                 int64_t D, const b200rl_ppo_args* args, float* dmu, float* dlog_std,
                                 void* workspace, size_t workspace_bytes, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K4c  GridNet per-cell MultiDiscrete heads (+ pick_position categoricals over the cells).
 * Replaces shared/actor/gridnet.py:38-193 over shared/actor/categorical.py:12-54.
 *   logits  [B, HW, S + n_pick]  f32 or bf16 (logits_dtype); S = sum(nvec)
 *   mask    [B, HW, S] u8;  pick_mask [B, n_pick, HW] u8 (NULL when n_pick == 0)
 *   actions [B, HW, A] (act_dtype U8/I32/I64); pick_actions [B, n_pick] (pick_dtype I32/I64)
 *   nvec_host [A]; gate_ref_host / gate_val_host [A]: head h only counts where
 *     actions[.., gate_ref[h]] == gate_val[h] (gate_ref[h] < 0: ungated)  (gridnet.py:119-127)
 *   logits_ld: elements between the logit rows of consecutive cells, in logits AND dlogits (0 = S + n_pick, dense).
 *     A trunk whose head emits channel-padded NHWC ([B, H, W, 80] for S = 78: the layout the convolution library
 *     wants) passes its padded width here; columns S + n_pick .. logits_ld of dlogits are written as zeros.
 * Mask semantics are the reference's, bit for bit: masked logits become finfo.min; a row
 * with no valid entry has log-prob 0, entropy -0 and zero gradient.
 */
typedef struct b200rl_gridnet_desc {
  int64_t B;
  int64_t HW;
  int A;
  int n_pick;
  int logits_dtype;
  int act_dtype;
  int pick_dtype;
  const int32_t* nvec_host;
  const int32_t* gate_ref_host;
  const int32_t* gate_val_host;
  int64_t logits_ld;
} b200rl_gridnet_desc;

/* Scratch for the two-launch scheme (streaming pre-pass writes per-sample lists of the cells that
 * have a valid action; the fused compute launch reads them).  The PPO entry point needs
 * b200rl_ppo_gridnet_workspace_bytes(): PPO partials first, GridNet lists after. */
size_t b200rl_gridnet_workspace_bytes(int64_t B, int64_t HW, int n_pick);
size_t b200rl_ppo_gridnet_workspace_bytes(int64_t B, int64_t HW, int n_pick, int64_t V);

int b200rl_gridnet_fwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                       const uint8_t* pick_mask, const void* actions, const void* pick_actions, float* logp,
                       float* entropy, void* workspace, size_t workspace_bytes, b200rl_stream_t stream);
int b200rl_gridnet_bwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                       const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                       const float* dlogp, const float* dentropy, void* dlogits, void* workspace,
                       size_t workspace_bytes, b200rl_stream_t stream);
/* Streaming pre-pass (zero fill of dlogits + mask compaction) followed by ONE fused launch: masked
 * logsumexp / log-prob / entropy forward, PPO ratio / clip / value-clip / entropy loss, and the
 * backward into dlogits and dvalues; logits are read only for cells with a valid action, dlogits
 * written once. */
int b200rl_ppo_gridnet_loss(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                            const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                            const b200rl_ppo_args* args, void* dlogits, float* logp_out /*nullable*/,
                            float* entropy_out /*nullable*/, void* workspace, size_t workspace_bytes,
                            b200rl_stream_t stream);
/* The same launch for a dlogits buffer that PERSISTS across calls (one minibatch after another of the same shape).
 * d loss / d logits is zero everywhere except the rows of cells with a valid action (2-6 % of a MicroRTS / Lux map), so
 * a buffer that was correct after the previous call only needs those rows cleared before this call's rows are
 * written: `rows` (device, b200rl_gridnet_rows_bytes(), owned by the caller together with dlogits) remembers which
 * rows a call wrote.  rows_valid == 0: `rows` holds nothing yet -- the whole of dlogits is zero-filled as in
 * b200rl_ppo_gridnet_loss and the written rows are recorded.  rows_valid != 0: the caller asserts that dlogits is
 * exactly what the previous call with this `rows` left (same shape, not written since); the previous rows are
 * cleared, this call's recorded.  dlogits traffic drops from HW * logits_ld elements per sample to the unit rows.
 * The result is bit-identical to b200rl_ppo_gridnet_loss. */
size_t b200rl_gridnet_rows_bytes(int64_t B, int64_t HW, int n_pick);
int b200rl_ppo_gridnet_loss_inplace(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                    const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                    const b200rl_ppo_args* args, void* dlogits, float* logp_out /*nullable*/,
                                    float* entropy_out /*nullable*/, void* workspace, size_t workspace_bytes,
                                    void* rows, size_t rows_bytes, int rows_valid, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * a3  Batch.num_actions.  Replaces rollout/rollout.py:130-180 (num_actions / per_position_num_actions) over a whole
 * rollout in one launch (d->B = T * N steps; d->logits_dtype / logits_ld are ignored):
 *   cells_out[r] = without gates (every gate_ref < 0): the number of cells of step r with any valid mask entry;
 *                  with gates (a subaction mask is configured): the number of (cell, action plane) pairs with a valid
 *                  entry, a gated plane counting only where actions[r, cell, gate_ref[h]] == gate_val[h];
 *   picks_out[r] = the number of cells any pick_position head may choose (n_pick > 0; NULL otherwise) -- the
 *                  reference adds log(picks) where picks > 0 (rollout.py:143-149), which the caller does.
 * Exact (integer counts).
 */
int b200rl_gridnet_num_actions(const b200rl_gridnet_desc* d, const uint8_t* mask, const uint8_t* pick_mask,
                               const void* actions, int32_t* cells_out, int32_t* picks_out, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K5  rollout-time sampling: one action per head per cell (+ pick) from the masked logits and
 * its log-prob, in one launch.  Replaces shared/actor/gridnet.py:195-207 (sample) +
 * log_prob at shared/policy/actor_critic.py:311-314.  Gumbel-max over a counter-based RNG
 * (Philox4x32-10 keyed by seed, counter = (offset, sample, cell, head)): the distribution is
 * the reference's, the random stream is not torch.multinomial's.  The effective offset is
 * `offset + *offset_dev` when offset_dev is non-NULL (a device-side step counter: one captured
 * launch draws fresh numbers on every graph replay).  A head with no valid entry returns action 0
 * (the reference draws uniformly there; such a head has log-prob 0 and no effect on training).
 */
int b200rl_gridnet_sample(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                          const uint8_t* pick_mask, uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                          void* actions_out, void* pick_actions_out, float* logp,
                          int64_t* actions_wide_out /*nullable: the per-cell actions once more as int64 [B, HW, A], what a
                          host env is handed (actor_critic.py:315-318) -- no cast on the host*/,
                          b200rl_stream_t stream);
int b200rl_categorical_sample_f32(const float* logits, const uint8_t* mask, int64_t R, int64_t n, uint64_t seed,
                                  uint64_t offset, const int64_t* offset_dev, int64_t* actions_out, float* logp,
                                  b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K0  rollout-buffer step write.  Replaces the per-step numpy slice assignments of
 * rollout/sync_step_rollout.py:188-201 (obs[s] = ..., episode_starts[s] = ..., fold_in(...)).
 * For each of n_tensors fields, copies step_bytes[t] bytes from src[t] (this step's [N, ...] slice)
 * to dst[t] + (*step_dev % T) * step_bytes[t] (row s of the [T, N, ...] buffer).  The step index
 * lives on the device so that the whole env step -- policy forward, sampling, buffer write -- can
 * be captured once in a CUDA graph and replayed T times.
 */
int b200rl_rollout_store_step(const void* const* src_host, void* const* dst_host, const int64_t* step_bytes_host,
                              int n_tensors, const int64_t* step_dev, int64_t T, b200rl_stream_t stream);
/* ... and the carry-over of sync_step_rollout.py:202-212 (self.next_obs = ..., self.next_action_masks = ...) in the
 * same launch: where carry_host[t] is non-NULL, src[t] (which must be writable) is overwritten with carry_host[t]
 * -- the env's output for the NEXT step -- right after its current content went to the buffer row. */
int b200rl_rollout_store_step_carry(const void* const* src_host, void* const* dst_host, const int64_t* step_bytes_host,
                                    const void* const* carry_host, int n_tensors, const int64_t* step_dev, int64_t T,
                                    b200rl_stream_t stream);

/* ... and the rest of the step's bookkeeping in that launch (b200rl_rollout_store_step_fused):
 *  - carry_or_host[t] non-NULL (with carry_host[t]): src[t] <- carry[t] | carry_or[t], bytewise -- the episode-start
 *    flags of the next step are terminations | truncations (sync_step_rollout.py:204-206);
 *  - pack (nullable): field pack->field holds observations in the trunk's layout, src [N, HW, Cp] float32 (planes C..
 *    zero); its carry is the env's RAW observation [N, C, HW] float32, transposed into that layout on the way
 *    (the per-step permute copy of the packed-observation path);
 *  - advance != 0: *step_dev is incremented once every CTA has read it (the last CTA to finish does it; `ticket` is a
 *    zero-initialised device int32 the caller keeps, left at zero again). */
typedef struct b200rl_store_pack {
  int field;             /* index into src_host / dst_host / carry_host */
  int64_t N;             /* envs */
  int64_t C;             /* observation planes */
  int64_t HW;            /* map cells */
  int64_t Cp;            /* planes per packed cell row (>= C, <= 128) */
} b200rl_store_pack;
int b200rl_rollout_store_step_fused(const void* const* src_host, void* const* dst_host, const int64_t* step_bytes_host,
                                    const void* const* carry_host, const void* const* carry_or_host, int n_tensors,
                                    const b200rl_store_pack* pack /*nullable*/, int64_t* step_dev, int64_t T, int advance,
                                    int32_t* ticket /*nullable unless advance*/, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K6  running-moment normalisers.  Replaces wrappers/normalize.py:18-122 over
 * utils/running_mean_std.py:10-33.  State (mean[D], var[D], count[D] -- the reference's scalar count,
 * kept per feature) is float64 on the device; with `training` the batch moments over the N envs are
 * merged first (Chan), then
 *   obs:    out = clip((x - mean) / sqrt(var + epsilon), -clip, clip)
 *   reward: returns = returns * gamma + r; update(returns); out = clip(r / sqrt(var + epsilon), ...);
 *           returns[done] = 0          (returns: [N, V] float64 accumulator)
 */
int b200rl_running_norm_obs_f32(const float* x, int64_t N, int64_t D, double* mean, double* var, double* count,
                                int training, double epsilon, double clip, float* out, b200rl_stream_t stream);
int b200rl_running_norm_reward_f32(const float* rewards, const uint8_t* dones, int64_t N, int64_t V, double gamma,
                                   double* returns, double* mean, double* var, double* count, int training,
                                   double epsilon, double clip, float* out, b200rl_stream_t stream);

/* NormalizeReward(exponential_moving_mean_var=True) (wrappers/normalize.py:74-78): the variance the reward is
 * divided by is HybridMovingMeanVar.var (utils/running_mean_std.py:120-170) -- the running moments above blended
 * into exponential-moving ones (ExponentialMovingMeanVar.update, :79-96: weights alpha (1-alpha)^(N-1-n) over
 * the batch rows, first batch = plain batch moments) as count / window grows to 1, window = 2 / alpha - 1.
 * ema_* are float64 [V] device arrays, ema_init int32 [V] (0 before the first update).
 * per_env != 0 (V must be 1) reproduces what the reference computes for SCALAR rewards (shape == ()): its update
 * broadcasts weights[:, None] against the 1-D batch, so after the first update every env keeps its own moving
 * moments (mean_j = sum_i w_i x_j + (1 - sum w) mean_j); ema_mean / ema_sq / ema_var then hold N entries. */
int b200rl_running_norm_reward_ema_f32(const float* rewards, const uint8_t* dones, int64_t N, int64_t V, double gamma,
                                       double* returns, double* mean, double* var, double* count, double* ema_mean,
                                       double* ema_sq, double* ema_var, int* ema_init, double alpha, int per_env,
                                       int training, double epsilon, double clip, float* out, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K7  multi-head reward assembly.  Replaces wrappers/info_rewards_wrapper.py:39-57
 * (InfoRewardsWrapper.step): out[n] = concat(base[n, :V0], series_0[n], ..., series_{K-1}[n]) where a series
 * flagged in episode_end_host[k] is zeroed unless terminations[n] | truncations[n] (:45-53) and every
 * series is multiplied by multiplier_host[k] when multiplier_host is not NULL (:54-55).  series_host is a
 * HOST array of K device pointers to [N] float32 vectors (the leaves of the env's infos dict).
 */
int b200rl_reward_assemble_f32(const float* base, int64_t V0, const float* const* series_host, int K,
                               const uint8_t* terminations, const uint8_t* truncations,
                               const uint8_t* episode_end_host, const float* multiplier_host, float* out, int64_t N,
                               b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K8  channels-last glue between the convolutions of the GridNet encoder / decoder, float32.
 * Replaces, around the cuDNN convolutions of shared/encoder/gridnet_encoder.py:26-51 (conv -> MaxPool2d(3, 2, 1) ->
 * ReLU) and shared/actor/gridnet_decoder.py:36-53 (transposed conv -> ReLU), PyTorch's separate bias-add, max-pool,
 * ReLU, their backward kernels and the bias-gradient reduction on the NHWC tensors the path hands the trunk.
 * Forward bit-identical to the PyTorch sequence (same adds, torch's arg-max rule: first maximum of the kh, kw scan,
 * NaN propagates); backward deterministic (gather form, fixed order).
 *
 *   x [N, H, W, C], bias [C] (nullable: no bias), out [N, Ho, Wo, C], Ho = (H + 2 padding - kernel) / stride + 1;
 *   argmax [N, Ho, Wo, C] uint8 (nullable when no backward follows): kh * kernel + kw of the maximum, 255 where
 *   the ReLU zeroed the output.  relu == 0: plain bias + max-pool.
 *   _bwd: dx [N, H, W, C] is written everywhere (zero where no window routes a gradient); dbias [C] nullable;
 *   workspace: b200rl_nhwc_bias_grad_workspace_bytes(N * Ho * Wo, C) bytes (only read / written when dbias != NULL).
 *   b200rl_nhwc_bias_relu_*: out = relu(x + bias) over [rows, C] (relu == 0: out = x + bias, the logit head's bias);
 *   out may alias x, dx may alias dout; workspace: b200rl_nhwc_bias_grad_workspace_bytes(rows, C).
 *   _bwd with out == NULL is the relu == 0 case: d x is dout itself (dx is not written), dbias its column sums. */
size_t b200rl_nhwc_bias_grad_workspace_bytes(int64_t rows, int64_t C);
int b200rl_nhwc_bias_pool_relu_fwd(const float* x, const float* bias /*nullable*/, float* out, uint8_t* argmax /*nullable*/,
                                   int64_t N, int64_t H, int64_t W, int64_t C, int kernel, int stride, int padding,
                                   int relu, b200rl_stream_t stream);
int b200rl_nhwc_bias_pool_relu_bwd(const float* dout, const uint8_t* argmax, float* dx, float* dbias /*nullable*/,
                                   void* workspace, size_t workspace_bytes, int64_t N, int64_t H, int64_t W, int64_t C,
                                   int kernel, int stride, int padding, b200rl_stream_t stream);
int b200rl_nhwc_bias_relu_fwd(const float* x, const float* bias, float* out, int64_t rows, int64_t C, int relu,
                              b200rl_stream_t stream);
int b200rl_nhwc_bias_relu_bwd(const float* dout, const float* out /*nullable*/, float* dx, float* dbias /*nullable*/, void* workspace,
                              size_t workspace_bytes, int64_t rows, int64_t C, b200rl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K9  channels-last glue of the squeeze U-net (shared/policy/actor_critic_network/squeeze_unet.py:20-196 over the
 * SE-residual blocks of double_cone.py:18-86), float32 or bfloat16 maps (dtype = B200RL_F32 / B200RL_BF16), float32
 * biases and reductions.  With bfloat16 maps every value PyTorch materialises as a bfloat16 tensor under autocast is
 * rounded at the same point (forward bit-identical to the PyTorch chain).
 *
 *   b200rl_nhwc_bias_act_*: out = act(x + bias) over [rows, C], act 0 none / 1 ReLU / 2 GELU (torch's exact erf form);
 *     _bwd: dx = dout * act'(x + bias) from the convolution's bias-free output x (no pre-activation tensor is kept),
 *     dbias (nullable) = column sums of dx; workspace b200rl_nhwc_bias_act_workspace_bytes(rows, C).
 *   SE tail of a residual block, out = gelu(x + (y2 + b2) * gate[n, c]) over [N, HW, C]:
 *     b200rl_se_mean_sums       sums[n, c] = sum over hw of y2 (the caller adds b2 and divides: the squeeze mean);
 *     b200rl_se_tail_fwd        the gated residual sum + output activation in one pass;
 *     b200rl_se_tail_gate_grad  dgate[n, c] = sum over hw of dz * (y2 + b2), dz = dout * gelu'(z); dz is left in dx (it IS
 *                               the block input's gradient through the residual sum);
 *     b200rl_se_tail_bwd        dy2 = dz * gate + dmean[n, c] (dmean: the mean path's gradient, already / HW),
 *                               db2 (nullable) = column sums of dy2.
 *   The two linears of the gate ([N, C] x [C, C/16]) stay library GEMMs on the host side of the ABI. */
size_t b200rl_nhwc_bias_act_workspace_bytes(int64_t rows, int64_t C);
int b200rl_nhwc_bias_act_fwd(const void* x, const float* bias, void* out, int64_t rows, int64_t C, int act, int dtype,
                             b200rl_stream_t stream);
int b200rl_nhwc_bias_act_bwd(const void* dout, const void* x, const float* bias, void* dx, float* dbias /*nullable*/,
                             void* workspace, size_t workspace_bytes, int64_t rows, int64_t C, int act, int dtype,
                             b200rl_stream_t stream);
size_t b200rl_se_workspace_bytes(int64_t N, int64_t HW, int64_t C);  /* per-slab partials of the two reductions */
int b200rl_se_mean_sums(const void* y2, float* sums, void* workspace, size_t workspace_bytes, int64_t N, int64_t HW,
                        int64_t C, int dtype, b200rl_stream_t stream);
int b200rl_se_tail_fwd(const void* x, const void* y2, const float* b2, const void* gate, void* out, int64_t N, int64_t HW,
                       int64_t C, int dtype, b200rl_stream_t stream);
int b200rl_se_tail_gate_grad(const void* dout, const void* x, const void* y2, const float* b2, const void* gate,
                             float* dgate, void* dx /*out: dz*/, void* workspace, size_t workspace_bytes, int64_t N,
                             int64_t HW, int64_t C, int dtype, b200rl_stream_t stream);
int b200rl_se_tail_bwd(const void* dz, const void* gate, const float* dmean, void* dy2, float* db2 /*nullable*/,
                       void* workspace, size_t workspace_bytes, int64_t N, int64_t HW, int64_t C, int dtype,
                       b200rl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H */

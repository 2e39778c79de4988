// PPO per-sample stage as its own launch (distribution-level path) + the stats finaliser
// shared by every fused-loss kernel.  See ppo_terms.cuh for the arithmetic and citations.
#include <math.h>

#include "ppo_terms.cuh"

namespace b200rl {

int ppo_make_dev(const b200rl_ppo_args* a, long long B, void* workspace, size_t workspace_bytes, PpoDev* out) {
  B200RL_REQUIRE(a != nullptr, "ppo: args is null");
  B200RL_REQUIRE(B >= 1, "ppo: B=%lld", B);
  B200RL_REQUIRE(a->old_logp && a->adv && a->old_values && a->returns && a->new_values && a->stats_out,
                 "ppo: null tensor pointer");
  B200RL_REQUIRE(a->V >= 1 && a->adv_v >= 1, "ppo: V=%lld adv_v=%lld", (long long)a->V, (long long)a->adv_v);
  B200RL_UNSUPPORTED(a->V > B200RL_MAX_VALUE_HEADS || a->adv_v > B200RL_MAX_VALUE_HEADS,
                     "ppo: more than %d value heads", B200RL_MAX_VALUE_HEADS);
  B200RL_REQUIRE(a->adv_mode >= 0 && a->adv_mode <= 3, "ppo: adv_mode %d", a->adv_mode);
  B200RL_REQUIRE(a->adv_mode == 0 || a->moments, "ppo: adv_mode %d needs moments", a->adv_mode);
  B200RL_REQUIRE(a->adv_v == 1 || a->adv_weights_host, "ppo: adv_v > 1 needs multi_reward_weights");
  B200RL_REQUIRE(a->vf_coef_host != nullptr, "ppo: vf_coef is null");
  B200RL_REQUIRE(workspace && workspace_bytes >= b200rl_ppo_workspace_bytes(B, a->V), "ppo: workspace too small");
  PpoDev P{};
  P.old_logp = a->old_logp, P.adv = a->adv, P.moments = a->moments;
  P.adv_v = (int)a->adv_v, P.adv_mode = a->adv_mode, P.has_w = a->adv_weights_host != nullptr;
  for (int v = 0; v < P.adv_v; ++v) P.w[v] = a->adv_weights_host ? a->adv_weights_host[v] : 0.f;
  P.old_values = a->old_values, P.returns = a->returns, P.new_values = a->new_values, P.dvalues = a->dvalues;
  P.V = (int)a->V;
  P.clip = (float)a->clip_range;
  P.ratio_lo = (float)(1.0 - a->clip_range);
  P.ratio_hi = (float)(1.0 + a->clip_range);
  P.vclip = a->clip_range_vf < 0 ? -1.f : (float)a->clip_range_vf;
  for (int v = 0; v < P.V; ++v) P.vf_coef[v] = a->vf_coef_host[v];
  P.ent_coef = a->ent_coef, P.pi_coef = a->pi_coef, P.pi_coef_dev = nullptr;
  P.halving = a->vf_halving, P.loss_scale = a->loss_scale;
  B200RL_UNSUPPORTED(a->vf_loss < B200RL_VF_MSE || a->vf_loss > B200RL_VF_L1, "ppo: vf_loss=%d", a->vf_loss);
  P.vf_loss = a->vf_loss;
  P.teacher_logp = a->teacher_logp, P.teacher_coef = a->teacher_kl_coef;
  P.teacher_unbiased = a->teacher_unbiased, P.teacher_importance = a->teacher_importance;
  P.B = B;
  P.partials = static_cast<double*>(workspace);
  // the last 2 * MAX_VALUE_HEADS floats of the workspace hold the derived (mean, denominator) pairs
  P.norm = reinterpret_cast<const float*>(static_cast<const uint8_t*>(workspace) + b200rl_ppo_workspace_bytes(B, a->V) -
                                          2 * B200RL_MAX_VALUE_HEADS * sizeof(float));
  P.stats_out = a->stats_out;
  *out = P;
  return B200RL_OK;
}

__global__ void ppo_prepare_kernel(PpoDev P) {
  const int Vm = P.adv_mode == 3 ? 1 : P.adv_v;
  if ((int)threadIdx.x < Vm) ppo_norm_pair(P, Vm, threadIdx.x, const_cast<float*>(P.norm));
}

int ppo_launch_prepare(const PpoDev& P, cudaStream_t stream) {
  if (P.adv_mode == 0) return B200RL_OK;
  ppo_prepare_kernel<<<1, B200RL_MAX_VALUE_HEADS, 0, stream>>>(P);
  return check_launch("ppo_prepare");
}

// One block of 1024 threads over the per-sample (or per-block) partial rows; see ppo_finalize_block.
__global__ void __launch_bounds__(1024) ppo_finalize_kernel(PpoDev P, long long rows, int ent_d) {
  extern __shared__ double s_scratch[];  // [nwarps][ns]
  pdl_wait();  // launched as a programmatic dependent of the kernel that writes the partial rows
  ppo_finalize_block(P, rows, ent_d, s_scratch);
}

int ppo_launch_finalize(const PpoDev& P, long long rows, int ent_d, cudaStream_t stream) {
  const int ns = ppo_nstat(P.V);
  int threads = 1024;
  if (rows < threads) threads = (int)(((rows + 31) / 32) * 32);
  if (threads < ((ns + 31) / 32) * 32) threads = ((ns + 31) / 32) * 32;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1), cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)(threads / 32) * ns * sizeof(double), cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr, cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, ppo_finalize_kernel, P, rows, ent_d) != cudaSuccess) return check_launch("ppo_finalize");
  return check_launch("ppo_finalize");
}

// ------------------------------------------------------------------------------------------
// Scalar stage kernels (distribution-level path)
constexpr int kScalarBlock = 256;

// phase 0 of the KL cut-off: per-block sums of (ratio - 1) - logratio
__global__ void __launch_bounds__(kScalarBlock) ppo_kl_partial_kernel(PpoDev P, const float* new_logp) {
  __shared__ double scratch[32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double kl[1] = {0.0};
  if (i < P.B) {
    const float lr = new_logp[i] - P.old_logp[i];
    kl[0] = (double)((expf(lr) - 1.f) - lr);
  }
  block_sum<double, 1>(kl, scratch);
  if (threadIdx.x == 0) P.partials[blockIdx.x] = kl[0];
}
__global__ void ppo_kl_decide_kernel(PpoDev P, int blocks, float kl_cutoff, float* pi_coef_state) {
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int b = 0; b < blocks; ++b) a += P.partials[b];
    const float approx_kl = (float)(a / (double)P.B);
    if (approx_kl > kl_cutoff) *pi_coef_state = 0.f;  // sticky for the rest of the learn_epoch
  }
}

__global__ void __launch_bounds__(kScalarBlock)
    ppo_scalar_kernel(PpoDev P, const float* new_logp, const float* entropy, int ent_d, float* dlogp, float* dentropy) {
  __shared__ double scratch[5 * 32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = ppo_nstat(P.V);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < P.B) {
    PolicyTerms t = ppo_policy_terms(P, i, new_logp[i]);
    dlogp[i] = t.dlogp;
    const float de = ppo_dentropy(P, ent_d);
    float es = 0.f;
    for (int d = 0; d < ent_d; ++d) {
      es += entropy[i * ent_d + d];
      dentropy[i * ent_d + d] = de;
    }
    acc[0] = t.surrogate, acc[1] = es, acc[2] = t.kl, acc[3] = t.clipped, acc[4] = t.teacher;
  }
  block_sum<double, 5>(acc, scratch);
  double* row = P.partials + (long long)blockIdx.x * ns;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 5; ++k) row[k] = acc[k];
  }
  for (int v = 0; v < P.V; ++v) {
    double va[2] = {0.0, 0.0};
    if (i < P.B) {
      float2 r = ppo_value_terms(P, i, v);
      va[0] = r.x, va[1] = r.y;
    }
    block_sum<double, 2>(va, scratch);
    if (threadIdx.x == 0) row[kPolicyStats + v] = va[0], row[kPolicyStats + P.V + v] = va[1];
  }
}

}  // namespace b200rl

extern "C" size_t b200rl_ppo_workspace_bytes(int64_t B, int64_t V) {
  if (B < 1) B = 1;
  if (V < 1) V = 1;
  // one row of partial stats per sample (grid-per-sample kernels) + per-block rows for the
  // Gaussian log_std gradient (<= 64 action dims)
  return (size_t)B * (size_t)(5 + 2 * V) * sizeof(double) + ((size_t)B / 128 + 2) * 64 * sizeof(double) +
         2 * B200RL_MAX_VALUE_HEADS * sizeof(float);
}

extern "C" int b200rl_ppo_scalar_loss_f32(const float* new_logp, const float* entropy, int64_t ent_d, int64_t B,
                                          const b200rl_ppo_args* args, float kl_cutoff, float* pi_coef_state,
                                          float* dlogp, float* dentropy, void* workspace, size_t workspace_bytes,
                                          b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(new_logp && entropy && dlogp && dentropy, "ppo_scalar_loss: null pointer");
  B200RL_REQUIRE(ent_d >= 1, "ppo_scalar_loss: ent_d=%lld", (long long)ent_d);
  PpoDev P;
  int rc = ppo_make_dev(args, B, workspace, workspace_bytes, &P);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  rc = ppo_launch_prepare(P, s);
  if (rc) return rc;
  const int blocks = (int)((B + kScalarBlock - 1) / kScalarBlock);
  if (kl_cutoff >= 0.f) {
    B200RL_REQUIRE(pi_coef_state != nullptr, "ppo_scalar_loss: kl_cutoff needs pi_coef_state");
    ppo_kl_partial_kernel<<<blocks, kScalarBlock, 0, s>>>(P, new_logp);
    ppo_kl_decide_kernel<<<1, 32, 0, s>>>(P, blocks, kl_cutoff, pi_coef_state);
  }
  P.pi_coef_dev = pi_coef_state;  // may be null: then the immediate pi_coef applies
  ppo_scalar_kernel<<<blocks, kScalarBlock, 0, s>>>(P, new_logp, entropy, (int)ent_d, dlogp, dentropy);
  rc = check_launch("ppo_scalar_loss");
  if (rc) return rc;
  return ppo_launch_finalize(P, blocks, (int)ent_d, s);
}

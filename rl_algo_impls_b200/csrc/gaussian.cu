// K4b: diagonal Gaussian head -- log-prob / entropy forward and the fused PPO loss.
// Replaces shared/actor/gaussian.py:11-16,42-45 over torch.distributions.Normal:
//   log_prob = sum_d [ -(a - mu)^2 / (2 var) - log(std) - log(sqrt(2 pi)) ],  std = exp(log_std)
//   entropy  = 0.5 + 0.5 log(2 pi) + log(std)   per dim, NOT summed (the reference returns
//   [B, act_dim], so ppo.py:351's entropy.mean() averages over B * act_dim).
#include "ppo_terms.cuh"

namespace b200rl {

constexpr int kGaussBlock = 128;
constexpr int kGaussMaxD = 64;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;
constexpr float kHalfLog2Pi = 0.91893853320467274178f;  // 0.5 * log(2 pi) == log(sqrt(2 pi))

struct GaussParams {
  const float* mu;
  const float* log_std;
  const float* actions;
  long long B;
  int D;
  float* logp;
  float* entropy;
  float* dmu;
  float* dlog_std;
  double* ls_partials;  // [blocks][D]
};

__device__ __forceinline__ float gauss_logp(const GaussParams& p, long long i) {
  float lp = 0.f;
  for (int d = 0; d < p.D; ++d) {
    const float ls = p.log_std[d];
    const float std = expf(ls);
    const float var = std * std;
    const float diff = p.actions[i * p.D + d] - p.mu[i * p.D + d];
    lp += -(diff * diff) / (2.f * var) - logf(std) - kLogSqrt2Pi;
  }
  return lp;
}

__global__ void __launch_bounds__(kGaussBlock) gauss_fwd_kernel(const GaussParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  p.logp[i] = gauss_logp(p, i);
  for (int d = 0; d < p.D; ++d) p.entropy[i * p.D + d] = 0.5f + kHalfLog2Pi + logf(expf(p.log_std[d]));
}

__global__ void __launch_bounds__(kGaussBlock) gauss_ppo_kernel(const GaussParams p, const PpoDev P) {
  __shared__ double scratch[5 * 32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = ppo_nstat(P.V);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  float dlogp = 0.f;
  const bool live = i < p.B;
  if (live) {
    const float lp = gauss_logp(p, i);
    PolicyTerms t = ppo_policy_terms(P, i, lp);
    dlogp = t.dlogp;
    float ent_sum = 0.f;
    for (int d = 0; d < p.D; ++d) ent_sum += 0.5f + kHalfLog2Pi + logf(expf(p.log_std[d]));
    acc[0] = t.surrogate, acc[1] = ent_sum, acc[2] = t.kl, acc[3] = t.clipped, acc[4] = t.teacher;
    if (p.logp) p.logp[i] = lp;
  }
  block_sum<double, 5>(acc, scratch);
  double* row = P.partials + (long long)blockIdx.x * ns;
  if (threadIdx.x == 0)
    for (int k = 0; k < 5; ++k) row[k] = acc[k];
  for (int v = 0; v < P.V; ++v) {
    double va[2] = {0.0, 0.0};
    if (live) {
      float2 r = ppo_value_terms(P, i, v);
      va[0] = r.x, va[1] = r.y;
    }
    block_sum<double, 2>(va, scratch);
    if (threadIdx.x == 0) row[kPolicyStats + v] = va[0], row[kPolicyStats + P.V + v] = va[1];
  }
  // d logp / d mu_d = (a - mu) / var ;  d logp / d log_std_d = (a - mu)^2 / var - 1 ;
  // d entropy_d / d log_std_d = 1 for each of the B rows
  const float de = ppo_dentropy(P, p.D);
  for (int d = 0; d < p.D; ++d) {
    double g[1] = {0.0};
    if (live) {
      const float std = expf(p.log_std[d]);
      const float var = std * std;
      const float diff = p.actions[i * p.D + d] - p.mu[i * p.D + d];
      p.dmu[i * p.D + d] = dlogp * diff / var;
      g[0] = (double)(dlogp * (diff * diff / var - 1.f) + de);
    }
    block_sum<double, 1>(g, scratch);
    if (threadIdx.x == 0) p.ls_partials[(long long)blockIdx.x * p.D + d] = g[0];
  }
}

__global__ void gauss_logstd_final_kernel(const GaussParams p, int blocks) {
  const int d = threadIdx.x;
  if (d >= p.D) return;
  double a = 0.0;
  for (int b = 0; b < blocks; ++b) a += p.ls_partials[(long long)b * p.D + d];
  p.dlog_std[d] = (float)a;
}

}  // namespace b200rl

extern "C" int b200rl_gaussian_fwd_f32(const float* mu, const float* log_std, const float* actions, int64_t B,
                                       int64_t D, float* logp, float* entropy, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(mu && log_std && actions && logp && entropy, "gaussian_fwd: null pointer");
  B200RL_REQUIRE(B >= 0 && D >= 1, "gaussian_fwd: bad shape");
  B200RL_UNSUPPORTED(D > kGaussMaxD, "gaussian_fwd: D=%lld exceeds %d", (long long)D, kGaussMaxD);
  if (B == 0) return B200RL_OK;
  GaussParams p{mu, log_std, actions, B, (int)D, logp, entropy, nullptr, nullptr, nullptr};
  const unsigned grid = (unsigned)((B + kGaussBlock - 1) / kGaussBlock);
  gauss_fwd_kernel<<<grid, kGaussBlock, 0, (cudaStream_t)stream>>>(p);
  return check_launch("gaussian_fwd");
}

extern "C" int b200rl_ppo_gaussian_loss_f32(const float* mu, const float* log_std, const float* actions, int64_t B,
                                            int64_t D, const b200rl_ppo_args* args, float* dmu, float* dlog_std,
                                            void* workspace, size_t workspace_bytes, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(mu && log_std && actions && dmu && dlog_std, "ppo_gaussian_loss: null pointer");
  B200RL_REQUIRE(D >= 1, "ppo_gaussian_loss: bad shape");
  B200RL_UNSUPPORTED(D > kGaussMaxD, "ppo_gaussian_loss: D=%lld exceeds %d", (long long)D, kGaussMaxD);
  PpoDev P;
  int rc = ppo_make_dev(args, B, workspace, workspace_bytes, &P);
  if (rc) return rc;
  GaussParams p{mu, log_std, actions, B, (int)D, nullptr, nullptr, dmu, dlog_std, nullptr};
  // the log_std partials live after the stats partials inside the same workspace
  p.ls_partials = P.partials + (size_t)B * ppo_nstat(P.V);
  const unsigned grid = (unsigned)((B + kGaussBlock - 1) / kGaussBlock);
  cudaStream_t s = (cudaStream_t)stream;
  rc = ppo_launch_prepare(P, s);
  if (rc) return rc;
  gauss_ppo_kernel<<<grid, kGaussBlock, 0, s>>>(p, P);
  gauss_logstd_final_kernel<<<1, kGaussMaxD, 0, s>>>(p, (int)grid);
  rc = check_launch("ppo_gaussian_loss");
  if (rc) return rc;
  return ppo_launch_finalize(P, grid, (int)D, s);
}

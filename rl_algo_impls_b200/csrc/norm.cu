// K6: running-moment observation / reward normalisers on the device.
//
// Replaces wrappers/normalize.py:18-122 over utils/running_mean_std.py:10-33 (RunningMeanStd):
//   batch mean / variance over the N envs, Chan's parallel merge into the running (mean, var,
//   count) in float64, then  clip((x - mean) / sqrt(var + eps), -clip, clip)   (observations) or
//   returns = returns * gamma + r ; update(returns) ; clip(r / sqrt(var + eps), -clip, clip) ;
//   returns[done] = 0                                                        (rewards).
// One launch per env step: a CTA owns a tile of FEAT features (1 / 4 / 16 / 32: the smallest that covers min(D, 32))
// for ALL N rows, so it can reduce, merge and normalise its own columns without a grid-wide barrier; its other
// BLOCK / FEAT thread rows split the N envs (BLOCK = 1024 when a 256-thread CTA would walk more than 8 rows per
// thread: narrow observations over thousands of envs), partial moments fold through a shared-memory tree.  State lives in HBM as float64
// (count is kept per feature so that no CTA reads a scalar another CTA is updating); everything
// is asynchronous and address-stable, i.e. capturable in the rollout step's CUDA graph.
#include "common.cuh"

namespace b200rl {


struct NormParams {
  const float* x;      // [N, D] observations, or rewards
  float* out;          // [N, D]
  double* mean;        // [D]
  double* var;         // [D]
  double* count;       // [D]
  double* returns;     // [N, D] discounted-return accumulator (reward mode) or null
  const uint8_t* dones;  // [N] (reward mode) or null
  long long N, D;
  double gamma, epsilon, clip;
  int training;
  // exponential-moving moments next to the running ones (HybridMovingMeanVar, running_mean_std.py:56-170) or null
  double* ema_mean;     // [D]
  double* ema_sq;       // [D] moving mean of x^2
  double* ema_var;      // [D]
  int* ema_init;        // [D] 0 until the first update (kept per feature, like count)
  double alpha, window;
  // The reference's ExponentialMovingMeanVar.update broadcasts weights[:, None] against a 1-D batch (scalar
  // rewards, shape == ()), which turns its moving moments into PER-ENV vectors after the first update
  // (running_mean_std.py:88-96: mean_j = sum_i w_i x_j + (1 - sum w) mean_j).  per_row reproduces exactly that:
  // D == 1 and ema_mean / ema_sq / ema_var hold N entries.
  int per_row;
};

template <int kNormFeat, int kNormBlock>
__global__ void __launch_bounds__(kNormBlock) running_norm_kernel(const NormParams p) {
  constexpr int kNormRows = kNormBlock / kNormFeat;
  __shared__ double s_sum[kNormRows][kNormFeat], s_sq[kNormRows][kNormFeat];
  __shared__ double s_wsum[kNormRows][kNormFeat], s_wsq[kNormRows][kNormFeat];
  __shared__ double s_batch[5];  // per_row mode: batch mean, mean of squares, variance, count / window, running variance
  __shared__ int s_first;
  const bool ema = p.ema_mean != nullptr;
  __shared__ double s_mean[kNormFeat], s_inv[kNormFeat];
  const int fx = threadIdx.x % kNormFeat, ry = threadIdx.x / kNormFeat;
  const long long f = (long long)blockIdx.x * kNormFeat + fx;
  const bool live = f < p.D;
  const bool reward = p.returns != nullptr;

  // ---- 1. (reward mode) returns = returns * gamma + r ; batch moments of the tracked quantity --------
  double sum = 0.0, sq = 0.0, wsum = 0.0, wsq = 0.0;
  if (live && p.training) {
    const double decay = 1.0 - p.alpha;
    for (long long n = ry; n < p.N; n += kNormRows) {
      double v;
      if (reward) {
        v = p.returns[n * p.D + f] * p.gamma + (double)p.x[n * p.D + f];
        p.returns[n * p.D + f] = v;
      } else {
        v = (double)p.x[n * p.D + f];
      }
      sum += v;
      sq += v * v;
      if (ema) {  // weights alpha (1 - alpha)^(N-1-n): the batch is a time-ordered window (running_mean_std.py:88-90)
        const double w = p.alpha * pow(decay, (double)(p.N - 1 - n));
        wsum += w * v;
        wsq += w * (v * v);
      }
    }
  }
  s_sum[ry][fx] = sum, s_sq[ry][fx] = sq;
  s_wsum[ry][fx] = wsum, s_wsq[ry][fx] = wsq;
  __syncthreads();
  for (int half = kNormRows / 2; half > 0; half >>= 1) {  // fixed tree over the thread rows: row 0 ends with the totals
    if (ry < half) {
      s_sum[ry][fx] += s_sum[ry + half][fx], s_sq[ry][fx] += s_sq[ry + half][fx];
      if (ema) s_wsum[ry][fx] += s_wsum[ry + half][fx], s_wsq[ry][fx] += s_wsq[ry + half][fx];
    }
    __syncthreads();
  }

  // ---- 2. Chan merge into the running moments (running_mean_std.py:16-29) --------------------------
  if (ry == 0 && live) {
    double mean = p.mean[f], var = p.var[f];
    if (p.training) {
      const double a = s_sum[0][fx], b = s_sq[0][fx];
      const double count = p.count[f], n = (double)p.N;
      const double batch_mean = a / n;
      const double batch_var = fmax(0.0, b / n - batch_mean * batch_mean);
      const double delta = batch_mean - mean, total = count + n;
      mean += delta * n / total;
      const double m2 = var * count + batch_var * n + delta * delta * count * n / total;
      var = m2 / total;
      p.mean[f] = mean, p.var[f] = var, p.count[f] = total;
      if (ema && p.per_row) {
        s_first = !p.ema_init[0];
        s_batch[0] = batch_mean, s_batch[1] = b / n, s_batch[2] = batch_var;
        p.ema_init[0] = 1;
      } else if (ema) {
        double em, esq, ev;
        if (!p.ema_init[f]) {  // first batch: plain batch moments (running_mean_std.py:81-86)
          em = batch_mean, esq = b / n, ev = batch_var;
          p.ema_init[f] = 1;
        } else {
          const double wa = s_wsum[0][fx], wb = s_wsq[0][fx];
          const double keep = pow(1.0 - p.alpha, n);  // 1 - sum of the weights
          em = wa + keep * p.ema_mean[f];
          esq = wb + keep * p.ema_sq[f];
          ev = esq - em * em;
        }
        p.ema_mean[f] = em, p.ema_sq[f] = esq, p.ema_var[f] = ev;
      }
    }
    if (ema && p.per_row) {
      s_batch[3] = p.count[f] / p.window, s_batch[4] = var;
    } else if (ema) {  // HybridMovingMeanVar.mean / .var: running moments until `window` samples were seen, then the moving ones
      const double frac = p.count[f] / p.window;
      if (frac >= 1.0) {
        mean = p.ema_mean[f], var = p.ema_var[f];
      } else {
        mean = mean * (1.0 - frac) + p.ema_mean[f] * frac;
        var = var * (1.0 - frac) + p.ema_var[f] * frac;
      }
    }
    s_mean[fx] = mean;
    s_inv[fx] = 1.0 / sqrt(var + p.epsilon);
  }
  __syncthreads();

  // ---- 3. normalise this tile's columns ------------------------------------------------------------
  if (!live) return;
  const double mean = reward ? 0.0 : s_mean[fx];
  double inv = s_inv[fx];
  const double keep = ema && p.per_row ? pow(1.0 - p.alpha, (double)p.N) : 0.0;
  for (long long n = ry; n < p.N; n += kNormRows) {
    if (ema && p.per_row) {  // env n's own moving moments of its discounted return (D == 1)
      double em = p.ema_mean[n], esq = p.ema_sq[n], ev = p.ema_var[n];
      if (p.training) {
        if (s_first) {
          em = s_batch[0], esq = s_batch[1], ev = s_batch[2];
        } else {
          const double r = p.returns[n];
          em = (1.0 - keep) * r + keep * em;
          esq = (1.0 - keep) * (r * r) + keep * esq;
          ev = esq - em * em;
        }
        p.ema_mean[n] = em, p.ema_sq[n] = esq, p.ema_var[n] = ev;
      }
      const double frac = s_batch[3];
      inv = 1.0 / sqrt((frac >= 1.0 ? ev : s_batch[4] * (1.0 - frac) + ev * frac) + p.epsilon);
    }
    double v = ((double)p.x[n * p.D + f] - mean) * inv;
    v = fmin(fmax(v, -p.clip), p.clip);
    p.out[n * p.D + f] = (float)v;
    if (reward && p.dones[n]) p.returns[n * p.D + f] = 0.0;  // wrappers/normalize.py:91
  }
}

template <int FEAT>
static void launch_norm_feat(const NormParams& p, cudaStream_t stream) {
  const unsigned tiles = (unsigned)((p.D + FEAT - 1) / FEAT);
  // 256 threads unless each would walk more than 8 rows (narrow observations over thousands of envs)
  if (p.N > 8 * (256 / FEAT)) running_norm_kernel<FEAT, 1024><<<tiles, 1024, 0, stream>>>(p);
  else running_norm_kernel<FEAT, 256><<<tiles, 256, 0, stream>>>(p);
}

static int launch_norm(const NormParams& p, cudaStream_t stream) {
  if (p.D > 0x7fffffffLL) {
    set_error("running_norm: too many features (%lld)", p.D);
    return B200RL_EUNSUPPORTED;
  }
  if (p.D <= 1) launch_norm_feat<1>(p, stream);
  else if (p.D <= 4) launch_norm_feat<4>(p, stream);
  else if (p.D <= 16) launch_norm_feat<16>(p, stream);
  else launch_norm_feat<32>(p, stream);
  return check_launch("running_norm");
}

}  // namespace b200rl

extern "C" int b200rl_running_norm_obs_f32(const float* x, int64_t N, int64_t D, double* mean, double* var,
                                           double* count, int training, double epsilon, double clip, float* out,
                                           b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(x && mean && var && count && out, "running_norm_obs: null pointer");
  B200RL_REQUIRE(N >= 1 && D >= 1, "running_norm_obs: bad shape N=%lld D=%lld", (long long)N, (long long)D);
  NormParams p{x, out, mean, var, count, nullptr, nullptr, N, D, 0.0, epsilon, clip, training,
               nullptr, nullptr, nullptr, nullptr, 0.0, 1.0, 0};
  return launch_norm(p, (cudaStream_t)stream);
}

extern "C" int b200rl_running_norm_reward_f32(const float* rewards, const uint8_t* dones, int64_t N, int64_t V,
                                              double gamma, double* returns, double* mean, double* var, double* count,
                                              int training, double epsilon, double clip, float* out,
                                              b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rewards && dones && returns && mean && var && count && out, "running_norm_reward: null pointer");
  B200RL_REQUIRE(N >= 1 && V >= 1, "running_norm_reward: bad shape N=%lld V=%lld", (long long)N, (long long)V);
  NormParams p{rewards, out, mean, var, count, returns, dones, N, V, gamma, epsilon, clip, training,
               nullptr, nullptr, nullptr, nullptr, 0.0, 1.0, 0};
  return launch_norm(p, (cudaStream_t)stream);
}

extern "C" int b200rl_running_norm_reward_ema_f32(const float* rewards, const uint8_t* dones, int64_t N, int64_t V,
                                                  double gamma, double* returns, double* mean, double* var,
                                                  double* count, double* ema_mean, double* ema_sq, double* ema_var,
                                                  int* ema_init, double alpha, int per_env, int training,
                                                  double epsilon, double clip, float* out, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rewards && dones && returns && mean && var && count && out, "running_norm_reward_ema: null pointer");
  B200RL_REQUIRE(ema_mean && ema_sq && ema_var && ema_init, "running_norm_reward_ema: null moving-moment pointer");
  B200RL_REQUIRE(N >= 1 && V >= 1, "running_norm_reward_ema: bad shape N=%lld V=%lld", (long long)N, (long long)V);
  B200RL_REQUIRE(alpha > 0.0 && alpha < 1.0, "running_norm_reward_ema: alpha=%g must be in (0, 1)", alpha);
  B200RL_REQUIRE(!per_env || V == 1, "running_norm_reward_ema: per-env moving moments are the scalar-reward case (V=1)");
  NormParams p{rewards, out, mean, var, count, returns, dones, N, V, gamma, epsilon, clip, training,
               ema_mean, ema_sq, ema_var, ema_init, alpha, 2.0 / alpha - 1.0, per_env};
  return launch_norm(p, (cudaStream_t)stream);
}

// K7: multi-head reward assembly on the device.
//
// Replaces wrappers/info_rewards_wrapper.py:39-57 (InfoRewardsWrapper.step): the K per-env series an env
// reports in `infos` are stacked behind the env's own reward head(s), series flagged `episode_end` only
// count on the step that ends an episode (np.where(done | ~episode_end, r, 0)), then each series is
// scaled by its multiplier -- the [N, V0 + K] float32 reward row the rollout buffer stores for
// multi-head critics (Lux: 13 heads).  One launch per env step, a thread per output element; no
// synchronisation and only fixed addresses, so it sits inside the rollout step's CUDA graph.
#include "common.cuh"

namespace b200rl {

constexpr int kMaxSeries = 32;

struct RewardParams {
  const float* base;                 // [N, V0]
  const float* series[kMaxSeries];   // K x [N]
  const uint8_t* terminations;       // [N]
  const uint8_t* truncations;        // [N] or null
  float multiplier[kMaxSeries];
  uint32_t episode_end;              // bit k: series k only counts on episode-ending steps
  uint32_t scaled;                   // bit k: multiply series k (the reference skips the multiply when multiplier is None)
  float* out;                        // [N, V0 + K]
  long long N;
  int V0, K;
};

__global__ void __launch_bounds__(256) reward_assemble_kernel(const __grid_constant__ RewardParams p) {
  const int V = p.V0 + p.K;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N * V) return;
  const long long n = i / V;
  const int v = (int)(i - n * V);
  float r;
  if (v < p.V0) {
    r = p.base[n * p.V0 + v];
  } else {
    const int k = v - p.V0;
    r = __ldg(p.series[k] + n);
    if ((p.episode_end >> k) & 1u) {
      const bool done = p.terminations[n] || (p.truncations != nullptr && p.truncations[n]);
      if (!done) r = 0.f;
    }
    if ((p.scaled >> k) & 1u) r = __fmul_rn(r, p.multiplier[k]);
  }
  p.out[i] = r;
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_reward_assemble_f32(const float* base, int64_t V0, const float* const* series_host, int K,
                                          const uint8_t* terminations, const uint8_t* truncations,
                                          const uint8_t* episode_end_host, const float* multiplier_host, float* out,
                                          int64_t N, b200rl_stream_t stream) {
  B200RL_REQUIRE(N >= 0 && V0 >= 0 && K >= 0 && V0 + K >= 1, "reward_assemble: bad shape N=%lld V0=%lld K=%d",
                 (long long)N, (long long)V0, K);
  B200RL_UNSUPPORTED(K > kMaxSeries, "reward_assemble: K=%d series (max %d)", K, kMaxSeries);
  if (N == 0) return B200RL_OK;
  B200RL_REQUIRE(out != nullptr && (V0 == 0 || base != nullptr), "reward_assemble: null pointer");
  B200RL_REQUIRE(K == 0 || (series_host != nullptr && episode_end_host != nullptr), "reward_assemble: null series table");
  RewardParams p{};
  p.base = base, p.terminations = terminations, p.truncations = truncations, p.out = out;
  p.N = N, p.V0 = (int)V0, p.K = K;
  bool any_end = false;
  for (int k = 0; k < K; ++k) {
    B200RL_REQUIRE(series_host[k] != nullptr, "reward_assemble: series %d is null", k);
    p.series[k] = series_host[k];
    if (episode_end_host[k]) p.episode_end |= 1u << k, any_end = true;
    if (multiplier_host != nullptr) p.scaled |= 1u << k, p.multiplier[k] = multiplier_host[k];
  }
  B200RL_REQUIRE(!any_end || terminations != nullptr, "reward_assemble: episode_end series need the terminations flags");
  const long long total = N * (V0 + K);
  reward_assemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("reward_assemble");
}

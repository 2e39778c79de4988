// K1: GAE(lambda) reverse-time scan + returns over a time-major [T, N, V] rollout.
//
// Replaces the numpy loop of shared/gae.py:113-123 and the returns add of
// rollout/vec_rollout.py:88.  One thread owns VEC adjacent lanes (lane = n*V + v) and walks
// t = T-1 .. 0; a warp therefore reads 32*VEC*4 contiguous bytes of rewards / values per
// step (128-bit loads when the row length allows) and the loads of UNROLL steps are issued
// before the dependent f64 recurrence consumes them.  Memory-bound: 16*V+1 bytes per
// env-step, a handful of flops.
//
// Arithmetic is the reference's, bit for bit: the non-terminal factor is f64 (bool operand),
// a scalar gamma multiplies next_value in f32 (weak Python scalar), the carry is f64 and
// only the store rounds to f32.  Every operation uses an explicit _rn intrinsic so that
// ptxas cannot contract a multiply-add the reference did not.
#include "common.cuh"

namespace b200rl {

struct GaeParams {
  const float* rewards;
  const float* values;
  const uint8_t* episode_starts;
  const uint8_t* next_episode_starts;
  const float* next_values;
  float* advantages;
  float* returns;
  long long T, N, V, L;  // L = N * V lanes
  int gamma_is_scalar;
  double gamma[B200RL_MAX_VALUE_HEADS];
  double gamma_lambda[B200RL_MAX_VALUE_HEADS];  // gamma * gae_lambda, rounded once in f64 as numpy does
};

template <int VEC>
struct Lanes {
  float x[VEC];
};

template <int VEC>
__device__ __forceinline__ Lanes<VEC> load_lanes(const float* p) {
  Lanes<VEC> r;
  if constexpr (VEC == 4) {
    float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(p));
    r.x[0] = v.x, r.x[1] = v.y, r.x[2] = v.z, r.x[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.x[i] = __ldg(p + i);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void store_lanes(float* p, const Lanes<VEC>& r) {
  if constexpr (VEC == 4) {
    stg_stream_f4(reinterpret_cast<float4*>(p), make_float4(r.x[0], r.x[1], r.x[2], r.x[3]));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = r.x[i];
  }
}

// episode_starts flags of the VEC lanes at one time row.  V == 1 && VEC == 4: one 32-bit load.
// V > 1 && VEC == 4: four adjacent lanes (lane0 a multiple of 4) belong to at most two envs, the
// first lane's and the last lane's, so two byte loads cover them.
template <int VEC, bool V1>
__device__ __forceinline__ uint32_t load_starts(const uint8_t* row, const long long (&env)[VEC]) {
  if constexpr (V1 && VEC == 4) {
    return __ldg(reinterpret_cast<const unsigned int*>(row + env[0]));
  } else if constexpr (VEC == 4) {
    const uint32_t lo = __ldg(row + env[0]) != 0, hi = __ldg(row + env[3]) != 0;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < VEC; ++i) packed |= (env[i] == env[0] ? lo : hi) << (8 * i);
    return packed;
  } else {
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < VEC; ++i) packed |= (uint32_t)(__ldg(row + env[i]) != 0) << (8 * i);
    return packed;
  }
}

template <int VEC, bool V1, int UNROLL>
__global__ void __launch_bounds__(128) gae_scan_kernel(const GaeParams p) {
  const long long group = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lane0 = group * VEC;
  if (lane0 >= p.L) return;

  double g[VEC], gl[VEC];
  long long env[VEC];  // env of each lane: constant over time, so the division happens once
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    env[i] = V1 ? (lane0 + i) : (lane0 + i) / p.V;
    int h = V1 ? 0 : (int)((lane0 + i) - env[i] * p.V);
    g[i] = p.gamma[h];
    gl[i] = p.gamma_lambda[h];
  }
  const bool scalar_gamma = p.gamma_is_scalar != 0;

  Lanes<VEC> v_next = load_lanes<VEC>(p.next_values + lane0);
  uint32_t started_next = load_starts<VEC, V1>(p.next_episode_starts, env);
  double carry[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) carry[i] = 0.0;

  long long t = p.T - 1;
  while (t >= 0) {
    Lanes<VEC> r[UNROLL], v[UNROLL];
    uint32_t st[UNROLL];
    // issue every load of this block of steps before the recurrence touches them
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long tu = t - u;
      if (tu >= 0) {
        r[u] = load_lanes<VEC>(p.rewards + tu * p.L + lane0);
        v[u] = load_lanes<VEC>(p.values + tu * p.L + lane0);
        st[u] = load_starts<VEC, V1>(p.episode_starts + tu * p.N, env);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long tu = t - u;
      if (tu >= 0) {
        Lanes<VEC> adv, ret;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const double alive = ((started_next >> (8 * i)) & 0xffu) ? 0.0 : 1.0;  // 1.0 - episode_starts[t+1]
          double boot;
          if (scalar_gamma) {
            boot = __dmul_rn((double)__fmul_rn((float)g[i], v_next.x[i]), alive);
          } else {
            boot = __dmul_rn(__dmul_rn(g[i], (double)v_next.x[i]), alive);
          }
          const double delta = __dsub_rn(__dadd_rn((double)r[u].x[i], boot), (double)v[u].x[i]);
          carry[i] = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl[i], alive), carry[i]));
          adv.x[i] = __double2float_rn(carry[i]);
          ret.x[i] = __fadd_rn(adv.x[i], v[u].x[i]);
        }
        store_lanes<VEC>(p.advantages + tu * p.L + lane0, adv);
        if (p.returns) store_lanes<VEC>(p.returns + tu * p.L + lane0, ret);
        v_next = v[u];
        started_next = st[u];
      }
    }
    t -= UNROLL;
  }
}

template <int VEC, bool V1>
static int launch(const GaeParams& p, cudaStream_t stream) {
  const long long groups = (p.L + VEC - 1) / VEC;
  const int block = 128;
  const long long grid = (groups + block - 1) / block;
  if (grid > 0x7fffffffLL) {
    set_error("gae_scan: too many lanes (%lld)", p.L);
    return B200RL_EUNSUPPORTED;
  }
  gae_scan_kernel<VEC, V1, 4><<<(unsigned)grid, block, 0, stream>>>(p);
  return check_launch("gae_scan");
}

}  // namespace b200rl

extern "C" int b200rl_gae_scan_f32(const float* rewards, const float* values, const uint8_t* episode_starts,
                                   const uint8_t* next_episode_starts, const float* next_values,
                                   const double* gamma_host, const double* gae_lambda_host, int gamma_is_scalar,
                                   float* advantages, float* returns, int64_t T, int64_t N, int64_t V,
                                   b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rewards && values && episode_starts && next_episode_starts && next_values && advantages,
                 "gae_scan: null pointer");
  B200RL_REQUIRE(gamma_host && gae_lambda_host, "gae_scan: gamma / gae_lambda are required");
  B200RL_REQUIRE(T >= 0 && N >= 0 && V >= 1, "gae_scan: bad shape T=%lld N=%lld V=%lld", (long long)T, (long long)N,
                 (long long)V);
  B200RL_UNSUPPORTED(V > B200RL_MAX_VALUE_HEADS, "gae_scan: V=%lld exceeds %d value heads", (long long)V,
                     B200RL_MAX_VALUE_HEADS);
  if (T == 0 || N == 0) return B200RL_OK;
  GaeParams p;
  p.rewards = rewards, p.values = values, p.episode_starts = episode_starts;
  p.next_episode_starts = next_episode_starts, p.next_values = next_values;
  p.advantages = advantages, p.returns = returns;
  p.T = T, p.N = N, p.V = V, p.L = N * V;
  p.gamma_is_scalar = gamma_is_scalar;
  for (int v = 0; v < V; ++v) {
    p.gamma[v] = gamma_host[v];
    p.gamma_lambda[v] = gamma_host[v] * gae_lambda_host[v];
  }
  cudaStream_t s = (cudaStream_t)stream;
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec4 = (p.L % 4 == 0) && aligned16(rewards) && aligned16(values) && aligned16(next_values) &&
                    aligned16(advantages) && (returns == nullptr || aligned16(returns)) &&
                    (V != 1 || ((reinterpret_cast<uintptr_t>(episode_starts) & 3u) == 0 &&
                                (reinterpret_cast<uintptr_t>(next_episode_starts) & 3u) == 0));
  if (vec4) return V == 1 ? launch<4, true>(p, s) : launch<4, false>(p, s);
  return V == 1 ? launch<1, true>(p, s) : launch<1, false>(p, s);
}

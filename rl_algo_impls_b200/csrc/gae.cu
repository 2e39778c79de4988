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

// episode_starts flags of the VEC lanes at one time row, one byte per lane in a 32-bit word.
// V == 1 && VEC == 4: one 32-bit load.  V > 1 && VEC == 4: four adjacent lanes (lane0 a multiple of 4) belong
// to at most two envs, the first lane's and the last lane's, so two byte loads cover them; `pat_lo` / `pat_hi`
// hold 0x01 in the byte of every lane that belongs to the first / last env.
struct LaneEnvs {
  int lo, hi;               // env of the first / last lane
  uint32_t pat_lo, pat_hi;  // byte masks of the lanes of each
};
template <int VEC, bool V1>
__device__ __forceinline__ uint32_t load_starts(const uint8_t* row, const LaneEnvs& e) {
  if constexpr (V1 && VEC == 4) {
    return __ldg(reinterpret_cast<const unsigned int*>(row + e.lo));
  } else if constexpr (V1) {
    return (uint32_t)(__ldg(row + e.lo) != 0);
  } else {
    const uint32_t lo = __ldg(row + e.lo) != 0, hi = __ldg(row + e.hi) != 0;
    return lo * e.pat_lo | hi * e.pat_hi;
  }
}

// Occupancy is what hides HBM latency here: 8 CTAs of 128 threads per SM (<= 64 registers, no spills) with the
// loads of UNROLL = 2 steps in flight per thread measured 0.91 of the HBM peak on both roofline shapes; deeper
// unrolling at 4-6 CTAs / SM measured 0.72-0.83 (V = 13) and 0.90 (V = 1).  The per-lane constants are kept
// small for that budget: envs / heads are packed into a few 32-bit words and the per-head discount factors are
// read from shared memory where they are used instead of living in 16 double registers.
template <int VEC, bool V1, int UNROLL, int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS) gae_scan_kernel(const __grid_constant__ GaeParams p) {
  const long long group = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lane0 = group * VEC;

  // env / value head of each lane are constant over time, so the division happens once; packed small
  LaneEnvs envs{(int)lane0, (int)lane0, 0u, 0u};
  uint32_t heads = 0u;  // byte i: value head of lane i
  if (!V1) {
    envs.lo = (int)(lane0 / p.V), envs.hi = (int)((lane0 + VEC - 1) / p.V);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int e = (int)((lane0 + i) / p.V);
      heads |= (uint32_t)((lane0 + i) - (long long)e * p.V) << (8 * i);
      if (e == envs.lo) envs.pat_lo |= 1u << (8 * i);
      else envs.pat_hi |= 1u << (8 * i);
    }
  }
  const bool scalar_gamma = p.gamma_is_scalar != 0;
  // per-head discount factors: read where they are used (volatile shared loads) so that the compiler does not
  // park 4 x 2 doubles per thread in registers for the whole scan
  __shared__ double s_g[B200RL_MAX_VALUE_HEADS], s_gl[B200RL_MAX_VALUE_HEADS];
  if (!V1) {
    if (threadIdx.x < B200RL_MAX_VALUE_HEADS) s_g[threadIdx.x] = p.gamma[threadIdx.x], s_gl[threadIdx.x] = p.gamma_lambda[threadIdx.x];
    __syncthreads();
  }
  if (lane0 >= p.L) return;

  Lanes<VEC> v_next = load_lanes<VEC>(p.next_values + lane0);
  uint32_t started_next = load_starts<VEC, V1>(p.next_episode_starts, envs);
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
        st[u] = load_starts<VEC, V1>(p.episode_starts + tu * p.N, envs);
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
          const double g = V1 ? p.gamma[0] : *(volatile double*)&s_g[(heads >> (8 * i)) & 0xffu];
          const double gl = V1 ? p.gamma_lambda[0] : *(volatile double*)&s_gl[(heads >> (8 * i)) & 0xffu];
          double boot;
          if (scalar_gamma) {
            boot = __dmul_rn((double)__fmul_rn((float)g, v_next.x[i]), alive);
          } else {
            boot = __dmul_rn(__dmul_rn(g, (double)v_next.x[i]), alive);
          }
          const double delta = __dsub_rn(__dadd_rn((double)r[u].x[i], boot), (double)v[u].x[i]);
          carry[i] = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, alive), carry[i]));
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
  // A rollout that fills the machine (>= 8 CTAs on every SM) runs the occupancy-first configuration; smaller
  // ones are a serial chain of dependent round trips per thread and measured best with 4 steps in flight
  // (T=512, N=24: 145 us against 173-207 us for 2, 8 or 16).
  const long long full = (long long)device_info().sm_count * 8 * block;
  if (groups >= full) gae_scan_kernel<VEC, V1, 2, 8><<<(unsigned)grid, block, 0, stream>>>(p);
  else gae_scan_kernel<VEC, V1, 4, 4><<<(unsigned)grid, block, 0, stream>>>(p);
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

// ---- K1b: GAE over ragged trajectories ------------------------------------------------------------
// Replaces compute_advantages per Trajectory (rollout/trajectory.py:56-95: shapes [T_i(, V)], the
// episode start of step t+1 is dones[t], the step after the last one uses dones[-1] / next_values)
// and DiscreteSkipsTrajectoryBuilder.trajectory (rollout/discrete_skips_trajectory_builder.py:84-100:
// delta = r + gamma^k v' - v, adv = delta + gamma^k lambda adv', k = steps_elapsed[t]).
// Segments are concatenated along the first axis; one thread walks one (segment, value head) lane.
namespace b200rl {

struct SegParams {
  const float* rewards;
  const float* values;
  const uint8_t* episode_starts;  // [total] (standard) or null (skips)
  const int32_t* steps_elapsed;   // [total] (skips) or null
  const long long* offsets;       // [n_seg + 1]
  const uint8_t* next_starts;     // [n_seg]: standard: dones[-1]; skips: trajectory done
  const float* next_values;       // [n_seg, V]
  float* advantages;
  float* returns;
  long long n_seg, V;
  int gamma_is_scalar;
  double gamma[B200RL_MAX_VALUE_HEADS], lambda[B200RL_MAX_VALUE_HEADS];
};

__global__ void __launch_bounds__(128) gae_segments_kernel(const SegParams p) {
  const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= p.n_seg * p.V) return;
  const long long seg = lane / p.V;
  const int h = (int)(lane - seg * p.V);
  const long long begin = p.offsets[seg], end = p.offsets[seg + 1];
  const double g = p.gamma[h], lam = p.lambda[h], gl = __dmul_rn(g, lam);
  const bool skips = p.steps_elapsed != nullptr;
  double carry = 0.0;
  float v_next = p.next_values[seg * p.V + h];
  bool start_next = p.next_starts[seg] != 0;
  for (long long t = end - 1; t >= begin; --t) {
    const float r = p.rewards[t * p.V + h], v = p.values[t * p.V + h];
    double adv;
    if (skips) {
      // the step after a finished trajectory has value 0; every product is float64 (numpy scalar gamma^k)
      const double gk = pow(g, (double)p.steps_elapsed[t]);
      const double nv = (t == end - 1 && start_next) ? 0.0 : (double)v_next;
      const double delta = __dsub_rn(__dadd_rn((double)r, __dmul_rn(gk, nv)), (double)v);
      carry = __dadd_rn(delta, __dmul_rn(__dmul_rn(gk, lam), carry));
      adv = carry;
    } else {
      const double alive = start_next ? 0.0 : 1.0;
      const double boot = p.gamma_is_scalar ? __dmul_rn((double)__fmul_rn((float)g, v_next), alive)
                                            : __dmul_rn(__dmul_rn(g, (double)v_next), alive);
      const double delta = __dsub_rn(__dadd_rn((double)r, boot), (double)v);
      carry = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, alive), carry));
      adv = carry;
      start_next = p.episode_starts[t] != 0;
    }
    const float a32 = __double2float_rn(adv);
    p.advantages[t * p.V + h] = a32;
    if (p.returns) p.returns[t * p.V + h] = __fadd_rn(a32, v);
    v_next = v;
  }
}

}  // namespace b200rl

extern "C" int b200rl_gae_segments_f32(const float* rewards, const float* values, const uint8_t* episode_starts,
                                       const int32_t* steps_elapsed, const int64_t* seg_offsets,
                                       const uint8_t* next_episode_starts, const float* next_values,
                                       const double* gamma_host, const double* gae_lambda_host, int gamma_is_scalar,
                                       float* advantages, float* returns, int64_t n_segments, int64_t V,
                                       b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rewards && values && seg_offsets && next_episode_starts && next_values && advantages,
                 "gae_segments: null pointer");
  B200RL_REQUIRE((episode_starts != nullptr) != (steps_elapsed != nullptr),
                 "gae_segments: pass episode_starts (trajectory GAE) or steps_elapsed (discrete skips), not both");
  B200RL_REQUIRE(gamma_host && gae_lambda_host && n_segments >= 0 && V >= 1, "gae_segments: bad arguments");
  B200RL_UNSUPPORTED(V > B200RL_MAX_VALUE_HEADS, "gae_segments: V=%lld exceeds %d", (long long)V, B200RL_MAX_VALUE_HEADS);
  if (n_segments == 0) return B200RL_OK;
  SegParams p{};
  p.rewards = rewards, p.values = values, p.episode_starts = episode_starts, p.steps_elapsed = steps_elapsed;
  p.offsets = reinterpret_cast<const long long*>(seg_offsets), p.next_starts = next_episode_starts;
  p.next_values = next_values, p.advantages = advantages, p.returns = returns;
  p.n_seg = n_segments, p.V = V, p.gamma_is_scalar = gamma_is_scalar;
  for (int v = 0; v < V; ++v) p.gamma[v] = gamma_host[v], p.lambda[v] = gae_lambda_host[v];
  const long long lanes = n_segments * V;
  gae_segments_kernel<<<(unsigned)((lanes + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("gae_segments");
}

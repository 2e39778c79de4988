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

// ---- K1, cooperative form for rollouts that do not fill the machine -----------------------------------
// At the config shapes (T=512 x 24 lanes, T=64 x 4096, T=32 x 13,312) the lane-per-thread scan above is a chain of
// T / 4 dependent global round trips per thread -- 135 us at T=512 -- although only TWO operations per step actually
// depend on the previous step: carry = delta_t + c_t * carry.  delta_t = r_t + gamma v_{t+1} alive_{t+1} - v_t and
// c_t = gamma lambda alive_{t+1} depend on the inputs alone.  So a CTA takes VEC adjacent lanes and
//   1. all its threads compute (delta_t, c_t) for a chunk of time steps side by side, in float64, into shared memory
//      (every global load independent of every other);
//   2. VEC threads walk the chunk backwards out of shared memory: two dependent float64 operations per step;
//   3. all threads store advantages / returns of the chunk.
// Same operations in the same order per lane as the reference loop, so it stays BIT-EXACT -- nothing is
// re-associated (a parallel affine scan would have to be, and would only meet 1e-5).
constexpr int kCoopBlock = 128;
constexpr int kCoopChunk = 512;  // time steps per chunk: VEC * 512 * (8 + 8 + 4 + 4) bytes = 48 KB at VEC = 4

template <int VEC, bool V1>
__global__ void __launch_bounds__(kCoopBlock) gae_coop_kernel(const __grid_constant__ GaeParams p) {
  __shared__ double s_delta[kCoopChunk * VEC];
  __shared__ double s_c[kCoopChunk * VEC];
  __shared__ float s_v[kCoopChunk * VEC];
  __shared__ float s_adv[kCoopChunk * VEC];
  const int tid = threadIdx.x;
  const long long lane0 = (long long)blockIdx.x * VEC;
  LaneEnvs envs{(int)lane0, (int)lane0, 0u, 0u};
  uint32_t heads = 0u;
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
  double carry = 0.0;  // threads 0 .. VEC-1: the lane's running advantage
  for (long long hi = p.T; hi > 0; hi -= kCoopChunk) {
    const long long lo = hi > kCoopChunk ? hi - kCoopChunk : 0;
    // 1. (delta, c) of every step of the chunk, side by side
    for (long long t = lo + tid; t < hi; t += kCoopBlock) {
      const Lanes<VEC> r = load_lanes<VEC>(p.rewards + t * p.L + lane0);
      const Lanes<VEC> v = load_lanes<VEC>(p.values + t * p.L + lane0);
      const bool last = t + 1 == p.T;
      const Lanes<VEC> vn = load_lanes<VEC>(last ? p.next_values + lane0 : p.values + (t + 1) * p.L + lane0);
      const uint32_t started_next = load_starts<VEC, V1>(last ? p.next_episode_starts : p.episode_starts + (t + 1) * p.N, envs);
      const int row = (int)(t - lo) * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const double alive = ((started_next >> (8 * i)) & 0xffu) ? 0.0 : 1.0;  // 1.0 - episode_starts[t+1]
        const int h = V1 ? 0 : (int)((heads >> (8 * i)) & 0xffu);
        const double g = p.gamma[h], gl = p.gamma_lambda[h];
        const double boot = scalar_gamma ? __dmul_rn((double)__fmul_rn((float)g, vn.x[i]), alive)
                                         : __dmul_rn(__dmul_rn(g, (double)vn.x[i]), alive);
        s_delta[row + i] = __dsub_rn(__dadd_rn((double)r.x[i], boot), (double)v.x[i]);
        s_c[row + i] = __dmul_rn(gl, alive);
        s_v[row + i] = v.x[i];
      }
    }
    __syncthreads();
    // 2. the recurrence: one thread per lane, operands from shared memory, eight steps' loads ahead of their use
    if (tid < VEC && lane0 + tid < p.L) {
      int k = (int)(hi - lo) - 1;
      for (; k >= 7; k -= 8) {
        double d[8], c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = s_delta[(k - u) * VEC + tid], c[u] = s_c[(k - u) * VEC + tid];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          carry = __dadd_rn(d[u], __dmul_rn(c[u], carry));
          s_adv[(k - u) * VEC + tid] = __double2float_rn(carry);
        }
      }
      for (; k >= 0; --k) {
        carry = __dadd_rn(s_delta[k * VEC + tid], __dmul_rn(s_c[k * VEC + tid], carry));
        s_adv[k * VEC + tid] = __double2float_rn(carry);
      }
    }
    __syncthreads();
    // 3. stores
    for (long long t = lo + tid; t < hi; t += kCoopBlock) {
      const int row = (int)(t - lo) * VEC;
      Lanes<VEC> adv, ret;
#pragma unroll
      for (int i = 0; i < VEC; ++i) adv.x[i] = s_adv[row + i], ret.x[i] = __fadd_rn(adv.x[i], s_v[row + i]);
      if (VEC == 1 && lane0 >= p.L) break;
      store_lanes<VEC>(p.advantages + t * p.L + lane0, adv);
      if (p.returns) store_lanes<VEC>(p.returns + t * p.L + lane0, ret);
    }
    __syncthreads();  // the next chunk overwrites the shared arrays
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
  // A rollout that fills the machine (>= 8 CTAs on every SM) streams at the HBM roofline with the lane-per-thread
  // scan (occupancy-first configuration).  Smaller ones are latency-bound there (a serial chain of dependent global
  // round trips per thread): they take the cooperative form, as long as the problem is small enough to sit in L2.
  const long long full = (long long)device_info().sm_count * 8 * block;
  const long long bytes = p.T * p.L * 17;
  if (groups >= full) gae_scan_kernel<VEC, V1, 2, 8><<<(unsigned)grid, block, 0, stream>>>(p);
  else if (bytes <= (64ll << 20) && groups <= 0x7fffffffLL) gae_coop_kernel<VEC, V1><<<(unsigned)groups, kCoopBlock, 0, stream>>>(p);
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
// Segments are concatenated along the first axis; one CTA takes one segment.
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

// One CTA per trajectory, same three phases as gae_coop_kernel: (delta, c) of a chunk of steps for every value head side
// by side (rows [t, V] are contiguous: coalesced; the discrete-skips gamma ** k -- a float64 pow -- is computed here, in
// parallel), then one thread per head runs the two-operation recurrence out of shared memory, then all threads store.
// A thread per (trajectory, head) walking global memory step by step, as before, is a chain of dependent round trips
// with strided, uncoalesced accesses.
constexpr int kSegBlock = 128;
constexpr int kSegEntries = 2048;  // (step, head) pairs per chunk: 2048 * (8 + 8 + 4 + 4) bytes = 48 KB

__global__ void __launch_bounds__(kSegBlock) gae_segments_kernel(const __grid_constant__ SegParams p) {
  __shared__ double s_delta[kSegEntries];
  __shared__ double s_c[kSegEntries];
  __shared__ float s_v[kSegEntries];
  __shared__ float s_adv[kSegEntries];
  const int tid = threadIdx.x;
  const long long seg = blockIdx.x;
  const int V = (int)p.V;
  const long long begin = p.offsets[seg], end = p.offsets[seg + 1];
  const bool skips = p.steps_elapsed != nullptr;
  const bool seg_start_next = p.next_starts[seg] != 0;
  const int chunk = kSegEntries / V;  // steps per chunk
  double carry = 0.0;                 // threads 0 .. V-1
  for (long long hi = end; hi > begin; hi -= chunk) {
    const long long lo = hi - begin > chunk ? hi - chunk : begin;
    const int n = (int)(hi - lo) * V;
    for (int e = tid; e < n; e += kSegBlock) {
      const long long t = lo + e / V;
      const int h = e - (int)(t - lo) * V;
      const float r = p.rewards[t * V + h], v = p.values[t * V + h];
      const bool last = t == end - 1;
      const float v_next = last ? p.next_values[seg * V + h] : p.values[(t + 1) * V + h];
      const double g = p.gamma[h], lam = p.lambda[h];
      double delta, c;
      if (skips) {
        // the step after a finished trajectory has value 0; every product is float64 (numpy scalar gamma^k)
        const double gk = pow(g, (double)p.steps_elapsed[t]);
        const double nv = (last && seg_start_next) ? 0.0 : (double)v_next;
        delta = __dsub_rn(__dadd_rn((double)r, __dmul_rn(gk, nv)), (double)v);
        c = __dmul_rn(gk, lam);
      } else {
        const bool start_next = last ? seg_start_next : p.episode_starts[t + 1] != 0;
        const double alive = start_next ? 0.0 : 1.0;
        const double boot = p.gamma_is_scalar ? __dmul_rn((double)__fmul_rn((float)g, v_next), alive)
                                              : __dmul_rn(__dmul_rn(g, (double)v_next), alive);
        delta = __dsub_rn(__dadd_rn((double)r, boot), (double)v);
        c = __dmul_rn(__dmul_rn(g, lam), alive);
      }
      s_delta[e] = delta, s_c[e] = c, s_v[e] = v;
    }
    __syncthreads();
    if (tid < V) {
      for (int k = (int)(hi - lo) - 1; k >= 0; --k) {
        carry = __dadd_rn(s_delta[k * V + tid], __dmul_rn(s_c[k * V + tid], carry));
        s_adv[k * V + tid] = __double2float_rn(carry);
      }
    }
    __syncthreads();
    for (int e = tid; e < n; e += kSegBlock) {
      const float a32 = s_adv[e];
      p.advantages[lo * V + e] = a32;
      if (p.returns) p.returns[lo * V + e] = __fadd_rn(a32, s_v[e]);
    }
    __syncthreads();
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
  B200RL_UNSUPPORTED(n_segments > 0x7fffffffLL, "gae_segments: %lld trajectories in one call", (long long)n_segments);
  gae_segments_kernel<<<(unsigned)n_segments, kSegBlock, 0, (cudaStream_t)stream>>>(p);
  return check_launch("gae_segments");
}

// Per-sample PPO loss terms shared by every fused-loss kernel (device side).
//
// Replaces the eager torch sequence of ppo/ppo.py:307-318 (advantage normalisation),
// :326-361 (ratio, clipped surrogate, clipped value loss, entropy loss, approx-KL, total),
// :373-374 (loss / num_minibatches) and the autograd backward of all of it, plus the stats
// of :379-409.  Gradients are the analytic derivatives torch autograd produces, including
// its tie rules: torch.min / torch.max split the gradient evenly on ties and clamp passes
// the gradient on its closed interval.
#pragma once
#include "common.cuh"

namespace b200rl {

constexpr int kPolicyStats = 5;  // surrogate sum, entropy sum, kl sum, clipped count, teacher-KL sum

struct PpoDev {
  const float* old_logp;
  const float* adv;
  const double* moments;
  int adv_v;
  int adv_mode;
  int has_w;
  float w[B200RL_MAX_VALUE_HEADS];
  const float* old_values;
  const float* returns;
  const float* new_values;
  float* dvalues;
  int V;
  float clip;      // (float)clip_range, for the clipped-fraction comparison
  float ratio_lo;  // (float)(1.0 - clip_range), computed in double like Python does
  float ratio_hi;
  float vclip;  // < 0: value clipping off
  float vf_coef[B200RL_MAX_VALUE_HEADS];
  float ent_coef;
  float pi_coef;
  const float* pi_coef_dev;  // when non-null, overrides pi_coef (KL cut-off state on the device)
  int halving;
  int vf_loss;  // B200RL_VF_*
  float loss_scale;
  const float* teacher_logp;  // null: no teacher-KL term
  float teacher_coef;
  int teacher_unbiased, teacher_importance;
  long long B;
  const float* norm;  // [2 * Vm] (mean, std + 1e-8) derived from `moments` by ppo_launch_prepare
  double* partials;  // [rows][4 + 2V]
  float* stats_out;
};

__host__ __device__ inline int ppo_nstat(int V) { return kPolicyStats + 2 * V; }

// (A - mean) / (std + 1e-8) etc. for one sample, then the reward-weight contraction.  The float64
// moments were turned into float (mean, denominator) pairs once per launch (ppo_prepare_kernel),
// so the per-sample path has no double-precision division or square root on it.
// `norm` is P.norm (written by ppo_prepare_kernel) or a CTA-local copy made with ppo_norm_pair.
__device__ __forceinline__ float ppo_sample_advantage(const PpoDev& P, long long i, const float* norm) {
  const float* row = P.adv + i * P.adv_v;
  if (P.adv_mode == 3) {
    float a = row[0];
    if (P.has_w) {
      a = 0.f;
#pragma unroll 1
      for (int v = 0; v < P.adv_v; ++v) a = fmaf(row[v], P.w[v], a);
    }
    return (a - norm[0]) / norm[1];
  }
  float acc = 0.f;
#pragma unroll 1
  for (int v = 0; v < P.adv_v; ++v) {
    float a = row[v];
    if (P.adv_mode != 0) {
      const float denom = norm[P.adv_v + v];
      a = (P.adv_mode == 1) ? (a - norm[v]) / denom : a / denom;
    }
    if (!P.has_w) return a;  // adv_v == 1 (checked on the host)
    acc = fmaf(a, P.w[v], acc);
  }
  return acc;
}

struct PolicyTerms {
  float dlogp;      // d loss / d new_logp[i]
  float surrogate;  // min(ratio*A, clamp(ratio)*A)
  float kl;         // (ratio - 1) - logratio
  float clipped;    // |ratio - 1| > clip
  float teacher;    // w * f(teacher_logp - new_logp), 0 without a teacher
};

// new_logp arrives in float64 when the caller summed it that way (GridNet: hundreds of per-cell
// terms); the subtraction is exact in float64 and rounds once, like an f32 subtraction of f32 inputs.
__device__ __forceinline__ PolicyTerms ppo_policy_terms_at(const PpoDev& P, long long i, double new_logp, float A,
                                                           float old_logp) {
  const float logratio = (float)(new_logp - (double)old_logp);
  const float ratio = expf(logratio);
  const float cr = fminf(fmaxf(ratio, P.ratio_lo), P.ratio_hi);
  const float s1 = ratio * A, s2 = cr * A;
  const bool in_range = ratio >= P.ratio_lo && ratio <= P.ratio_hi;
  float dsurr;
  if (in_range) {
    dsurr = A;  // both branches of the tie carry grad/2 and clamp is transparent
  } else if (s1 < s2) {
    dsurr = A;
  } else if (s1 == s2) {
    dsurr = 0.5f * A;  // tie outside the clip range: only the unclipped half reaches ratio
  } else {
    dsurr = 0.f;
  }
  const float pi_coef = P.pi_coef_dev ? *P.pi_coef_dev : P.pi_coef;
  PolicyTerms t;
  t.surrogate = fminf(s1, s2);
  t.dlogp = -(pi_coef * P.loss_scale / (float)P.B) * dsurr * ratio;
  t.kl = (ratio - 1.f) - logratio;
  t.clipped = fabsf(ratio - 1.f) > P.clip ? 1.f : 0.f;
  t.teacher = 0.f;
  if (P.teacher_logp) {  // loss/teacher_kl_loss.py:35-50; the weight (ratio) carries gradient like in the reference
    const float d = (float)((double)P.teacher_logp[i] - new_logp);
    float f, df;  // f(d) and d f / d new_logp
    if (P.teacher_unbiased) {
      const float e = expf(d);
      f = (e - 1.f) - d, df = 1.f - e;
    } else {
      f = 0.5f * d * d, df = -d;
    }
    const float scale = P.teacher_coef * P.loss_scale / (float)P.B;
    if (P.teacher_importance) {
      t.teacher = ratio * f;
      t.dlogp += scale * ratio * (f + df);
    } else {
      t.teacher = f;
      t.dlogp += scale * df;
    }
  }
  return t;
}

// (the normalised advantage and the behaviour log-prob do not depend on the new log-prob: a kernel with a long
// per-sample reduction computes them ahead and calls ppo_policy_terms_at)
__device__ __forceinline__ PolicyTerms ppo_policy_terms(const PpoDev& P, long long i, double new_logp,
                                                        const float* norm) {
  return ppo_policy_terms_at(P, i, new_logp, ppo_sample_advantage(P, i, norm), P.old_logp[i]);
}

__device__ __forceinline__ PolicyTerms ppo_policy_terms(const PpoDev& P, long long i, double new_logp) {
  return ppo_policy_terms(P, i, new_logp, P.norm);
}

// (sum, sum of squares, count) in f64 -> (mean, unbiased std + 1e-8) in f32 for head v, exactly the values the
// reference's mb_adv.mean(0) / mb_adv.std(0) + 1e-8 feed into the division (ppo.py:307-316).  Vm = heads
// the moments were taken over (1 for the after-weighting mode).
__device__ __forceinline__ void ppo_norm_pair(const PpoDev& P, int Vm, int v, float* norm) {
  const double n = P.moments[2 * Vm], mean = P.moments[v] / n;
  const double var = fmax(0.0, (P.moments[Vm + v] - P.moments[v] * mean) / (n - 1.0));
  norm[v] = (float)mean;
  norm[Vm + v] = (float)sqrt(var) + 1e-8f;
}

// d loss / d entropy element (entropy_loss = -mean over B * ent_d elements)
__device__ __forceinline__ float ppo_dentropy(const PpoDev& P, int ent_d) {
  return -(P.ent_coef * P.loss_scale) / ((float)P.B * (float)ent_d);
}

// element of vf_loss_fn(x, y, reduction="none") and its derivative in x, e = x - y (ppo.py:186,331-339): the
// torch.nn.functional losses at their default delta / beta = 1, forward and autograd backward formulas
__device__ __forceinline__ void vf_element(int kind, float e, float& loss, float& grad) {
  const float a = fabsf(e);
  switch (kind) {
    case B200RL_VF_HUBER:
    case B200RL_VF_SMOOTH_L1:
      if (a < 1.f) {
        loss = 0.5f * e * e, grad = e;
      } else {
        loss = a - 0.5f, grad = e < 0.f ? -1.f : 1.f;
      }
      break;
    case B200RL_VF_L1:
      loss = a, grad = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
      break;
    default:
      loss = e * e, grad = 2.f * e;
  }
}

// value head v of sample i: writes dvalues[i, v]; returns (loss element, clipped indicator)
__device__ __forceinline__ float2 ppo_value_terms(const PpoDev& P, long long i, int v) {
  const long long o = i * P.V + v;
  const float nv = P.new_values[o], ov = P.old_values[o], rt = P.returns[o];
  float u, g;
  vf_element(P.vf_loss, nv - rt, u, g);
  float vl = u, clipped = 0.f;
  if (P.vclip >= 0.f) {
    const float d = nv - ov;
    const float c = ov + fminf(fmaxf(d, -P.vclip), P.vclip);
    float cl, gc;
    vf_element(P.vf_loss, c - rt, cl, gc);
    if (!(d >= -P.vclip && d <= P.vclip)) gc = 0.f;  // clamp passes no gradient outside the band
    if (u > cl) {
      vl = u;
    } else if (u < cl) {
      vl = cl, g = gc;
    } else {
      vl = u, g = 0.5f * g + 0.5f * gc;
    }
    clipped = fabsf(d) > P.vclip ? 1.f : 0.f;
  }
  const float half = P.halving ? 0.5f : 1.f;
  if (P.dvalues) P.dvalues[o] = P.vf_coef[v] * half * P.loss_scale / (float)P.B * g;
  return make_float2(vl, clipped);
}

// partials [rows][5 + 2V] -> stats_out, by ONE block (any size that is a multiple of 32).  A warp's lanes are laid out
// as (row slot, column): with cpl = the power of two >= ns columns per slot, a warp sums 32 / cpl rows at a time --
// every load of a sweep is independent and a row's columns are adjacent (coalesced) -- the row slots fold with
// shuffles, warps combine through `scratch` ([nwarps][ns] doubles of shared memory), thread c < ns finishes column
// c.  More than 32 columns (V > 13): a thread per row, 8 columns at a time.  Fixed order => deterministic.
__device__ __forceinline__ void ppo_finalize_block(const PpoDev& P, long long rows, int ent_d, double* scratch) {
  const int ns = ppo_nstat(P.V);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __syncthreads();  // scratch may alias shared memory that was in use
  if (ns <= 32) {
    int cpl = 1;
    while (cpl < ns) cpl <<= 1;
    const int rpw = 32 / cpl, rs = lane / cpl, c = lane - rs * cpl;
    double acc = 0.0;
    if (c < ns) {
      long long r = (long long)warp * rpw + rs;
      const long long step = (long long)nwarps * rpw;
      for (; r + 3 * step < rows; r += 4 * step) {  // four independent loads in flight
        const double a0 = P.partials[r * ns + c], a1 = P.partials[(r + step) * ns + c];
        const double a2 = P.partials[(r + 2 * step) * ns + c], a3 = P.partials[(r + 3 * step) * ns + c];
        acc += a0, acc += a1, acc += a2, acc += a3;
      }
      for (; r < rows; r += step) acc += P.partials[r * ns + c];
    }
    for (int o = cpl; o < 32; o <<= 1) {
      int lo = __double2loint(acc), hi = __double2hiint(acc);
      lo = __shfl_xor_sync(0xffffffffu, lo, o), hi = __shfl_xor_sync(0xffffffffu, hi, o);
      acc += __hiloint2double(hi, lo);
    }
    if (lane < cpl && c < ns) scratch[warp * ns + c] = acc;
  } else {
    for (int c0 = 0; c0 < ns; c0 += 8) {
      double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (long long r = tid; r < rows; r += blockDim.x) {
        const double* row = P.partials + r * ns + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < ns) a[j] += row[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j >= ns) break;
        double x = a[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          int lo = __double2loint(x), hi = __double2hiint(x);
          lo = __shfl_xor_sync(0xffffffffu, lo, o), hi = __shfl_xor_sync(0xffffffffu, hi, o);
          x += __hiloint2double(hi, lo);
        }
        if (lane == 0) scratch[warp * ns + c0 + j] = x;
      }
    }
  }
  __syncthreads();
  // column c: sum over the warps, then its batch mean in float64 (one division per column, side by side)
  __shared__ float mean[kPolicyStats + 2 * B200RL_MAX_VALUE_HEADS];
  if (tid < ns) {
    double a = 0.0;
    for (int w = 0; w < nwarps; ++w) a += scratch[w * ns + tid];
    const double B = (double)P.B;
    mean[tid] = tid == 0 ? (float)(-a / B) : (tid == 1 ? (float)(-a / (B * ent_d)) : (float)(a / B));
  }
  __syncthreads();
  if (tid == 0) {
    const float pi_coef = P.pi_coef_dev ? *P.pi_coef_dev : P.pi_coef;
    const float pi_loss = mean[0], ent_loss = mean[1];
    float total = pi_coef * pi_loss + P.ent_coef * ent_loss;
    float vsum = 0.f;
    for (int v = 0; v < P.V; ++v) {
      float vl = mean[kPolicyStats + v];
      if (P.halving) vl *= 0.5f;
      P.stats_out[5 + v] = vl;
      P.stats_out[5 + P.V + v] = mean[kPolicyStats + P.V + v];
      vsum += P.vf_coef[v] * vl;
    }
    total += vsum;
    const float teacher_loss = mean[4];
    if (P.teacher_logp) total += P.teacher_coef * teacher_loss;
    P.stats_out[5 + 2 * P.V] = teacher_loss;
    P.stats_out[0] = total * P.loss_scale;
    P.stats_out[1] = pi_loss;
    P.stats_out[2] = ent_loss;
    P.stats_out[3] = mean[2];
    P.stats_out[4] = mean[3];
  }
}

// moments (f64 sum, sum of squares, count) -> norm (f32 mean, std + 1e-8); no-op for adv_mode 0.
int ppo_launch_prepare(const PpoDev& P, cudaStream_t stream);

// partials [rows][4 + 2V] -> stats_out (see b200rl.h).  One block; fixed order => deterministic.
int ppo_launch_finalize(const PpoDev& P, long long rows, int ent_d, cudaStream_t stream);

// host: b200rl_ppo_args -> PpoDev.  Returns 0 or an error code (message set).
int ppo_make_dev(const b200rl_ppo_args* a, long long B, void* workspace, size_t workspace_bytes, PpoDev* out);

}  // namespace b200rl

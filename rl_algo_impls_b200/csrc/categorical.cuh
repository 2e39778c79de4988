// One masked categorical over n entries: the arithmetic of shared/actor/categorical.py:12-54
// on top of torch.distributions.Categorical (logits - logsumexp, gather, -sum(p*logp)).
//
// Mask semantics follow the reference exactly: masked logits are finfo.min, so they
// contribute exp(.) == 0 to the partition sum; a row with no valid entry normalises to all
// zeros (log_prob 0 for any action, entropy 0) and its gradient is blocked by torch.where.
#pragma once
#include "common.cuh"

namespace b200rl {

struct CatRow {
  float lse;      // logsumexp over the valid entries
  float logp;     // log-prob of the chosen action
  float entropy;  // -sum_valid p * logp
  bool any;       // at least one valid entry
};

// X(k) -> float logit, M(k) -> bool valid
template <typename X, typename M>
__device__ __forceinline__ CatRow cat_forward(X x, M valid, int n, int action) {
  CatRow r;
  float mx = -INFINITY;
  bool any = false;
  for (int k = 0; k < n; ++k)
    if (valid(k)) {
      any = true;
      mx = fmaxf(mx, x(k));
    }
  r.any = any;
  if (!r.any) {
    r.lse = 0.f, r.logp = 0.f, r.entropy = 0.f;
    return r;
  }
  float sum = 0.f;
  for (int k = 0; k < n; ++k)
    if (valid(k)) sum += expf(x(k) - mx);
  r.lse = mx + logf(sum);
  float ent = 0.f;
  for (int k = 0; k < n; ++k)
    if (valid(k)) {
      const float lp = x(k) - r.lse;
      ent -= lp * expf(lp);
    }
  r.entropy = ent;
  const bool a_ok = action >= 0 && action < n && valid(action);
  r.logp = (a_ok ? x(action) : kF32Lowest) - r.lse;
  return r;
}

// gradient of  dlogp * logp + dent * entropy  w.r.t. entry k of the raw logits
__device__ __forceinline__ float cat_grad(float xk, bool valid_k, bool is_action, const CatRow& r, float dlogp,
                                          float dent) {
  if (!valid_k || !r.any) return 0.f;
  const float lp = xk - r.lse;
  const float p = expf(lp);
  return dlogp * ((is_action ? 1.f : 0.f) - p) - dent * p * (lp + r.entropy);
}

}  // namespace b200rl

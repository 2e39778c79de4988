// K4a: (Masked)Categorical over [R, n] logits -- forward, backward, fused PPO loss, sampling.
// Replaces shared/actor/categorical.py:12-54 + torch.distributions.Categorical and, for the
// fused entry point, ppo/ppo.py:326-361 and their autograd backward.  One thread per row:
// n is a handful of actions (Discrete(2), Discrete(4)) for the configs that use this head.
#include "categorical.cuh"
#include "philox.cuh"
#include "ppo_terms.cuh"

namespace b200rl {

constexpr int kCatBlock = 256;

template <typename ActT>
__device__ __forceinline__ int load_action(const void* actions, long long i) {
  return (int)static_cast<const ActT*>(actions)[i];
}

struct CatParams {
  const float* logits;
  const uint8_t* mask;
  const void* actions;
  long long R;
  int n;
  float* logp;
  float* entropy;
  const float* dlogp;
  const float* dentropy;
  float* dlogits;
};

template <typename ActT>
__global__ void __launch_bounds__(kCatBlock) cat_fwd_kernel(const CatParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.R) return;
  const float* x = p.logits + i * p.n;
  const uint8_t* m = p.mask ? p.mask + i * p.n : nullptr;
  CatRow r = cat_forward([&](int k) { return x[k]; }, [&](int k) { return m ? m[k] != 0 : true; }, p.n,
                         load_action<ActT>(p.actions, i));
  p.logp[i] = r.logp;
  p.entropy[i] = r.entropy;
}

template <typename ActT>
__global__ void __launch_bounds__(kCatBlock) cat_bwd_kernel(const CatParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.R) return;
  const float* x = p.logits + i * p.n;
  const uint8_t* m = p.mask ? p.mask + i * p.n : nullptr;
  const int a = load_action<ActT>(p.actions, i);
  CatRow r = cat_forward([&](int k) { return x[k]; }, [&](int k) { return m ? m[k] != 0 : true; }, p.n, a);
  const float dl = p.dlogp[i], de = p.dentropy[i];
  for (int k = 0; k < p.n; ++k)
    p.dlogits[i * p.n + k] = cat_grad(x[k], m ? m[k] != 0 : true, k == a, r, dl, de);
}

template <typename ActT>
__global__ void __launch_bounds__(kCatBlock) cat_ppo_kernel(const CatParams p, const PpoDev P) {
  __shared__ double scratch[5 * 32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = ppo_nstat(P.V);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < p.R) {
    const float* x = p.logits + i * p.n;
    const uint8_t* m = p.mask ? p.mask + i * p.n : nullptr;
    const int a = load_action<ActT>(p.actions, i);
    CatRow r = cat_forward([&](int k) { return x[k]; }, [&](int k) { return m ? m[k] != 0 : true; }, p.n, a);
    PolicyTerms t = ppo_policy_terms(P, i, r.logp);
    const float de = ppo_dentropy(P, 1);
    for (int k = 0; k < p.n; ++k)
      p.dlogits[i * p.n + k] = cat_grad(x[k], m ? m[k] != 0 : true, k == a, r, t.dlogp, de);
    if (p.logp) p.logp[i] = r.logp;
    if (p.entropy) p.entropy[i] = r.entropy;
    acc[0] = t.surrogate, acc[1] = r.entropy, acc[2] = t.kl, acc[3] = t.clipped, acc[4] = t.teacher;
  }
  block_sum<double, 5>(acc, scratch);
  double* row = P.partials + (long long)blockIdx.x * ns;
  if (threadIdx.x == 0)
    for (int k = 0; k < 5; ++k) row[k] = acc[k];
  for (int v = 0; v < P.V; ++v) {
    double va[2] = {0.0, 0.0};
    if (i < p.R) {
      float2 r = ppo_value_terms(P, i, v);
      va[0] = r.x, va[1] = r.y;
    }
    block_sum<double, 2>(va, scratch);
    if (threadIdx.x == 0) row[kPolicyStats + v] = va[0], row[kPolicyStats + P.V + v] = va[1];
  }
}

// Gumbel-max: argmax_k (x_k + g_k) over the valid entries is a draw from softmax(x | valid).
__global__ void __launch_bounds__(kCatBlock)
    cat_sample_kernel(const CatParams p, uint64_t seed, uint64_t offset0, const long long* offset_dev,
                      long long* actions_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.R) return;
  const uint64_t offset = offset0 + (offset_dev ? (uint64_t)*offset_dev : 0ull);
  const float* x = p.logits + i * p.n;
  const uint8_t* m = p.mask ? p.mask + i * p.n : nullptr;
  bool any = false;
  for (int k = 0; k < p.n; ++k) any |= (m ? m[k] != 0 : true);
  int best = 0;
  float best_score = -INFINITY;
  for (int k0 = 0; k0 < p.n; k0 += 4) {
    const Philox4 r = philox4x32_10(seed, (uint64_t)i, offset + (uint64_t)(k0 >> 2));
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    for (int j = 0; j < 4 && k0 + j < p.n; ++j) {
      const int k = k0 + j;
      const bool ok = !any || (m ? m[k] != 0 : true);  // a fully masked row is uniform over all entries
      if (!ok) continue;
      const float g = -logf(-logf(u01(bits[j])));
      const float score = (any ? x[k] : 0.f) + g;
      if (score > best_score) best_score = score, best = k;
    }
  }
  actions_out[i] = best;
  if (p.logp) {
    CatRow r = cat_forward([&](int k) { return x[k]; }, [&](int k) { return m ? m[k] != 0 : true; }, p.n, best);
    p.logp[i] = r.logp;
  }
}

static int check_cat(const float* logits, const void* actions, int act_dtype, int64_t R, int64_t n, const char* who) {
  B200RL_REQUIRE(logits && actions, "%s: null pointer", who);
  B200RL_REQUIRE(R >= 0 && n >= 1 && n <= (1 << 20), "%s: bad shape R=%lld n=%lld", who, (long long)R, (long long)n);
  B200RL_UNSUPPORTED(act_dtype != B200RL_U8 && act_dtype != B200RL_I32 && act_dtype != B200RL_I64,
                     "%s: action dtype %d", who, act_dtype);
  return B200RL_OK;
}

#define DISPATCH_ACT(act_dtype, KERNEL, ...)                       \
  switch (act_dtype) {                                             \
    case B200RL_U8: KERNEL<uint8_t> __VA_ARGS__; break;            \
    case B200RL_I32: KERNEL<int32_t> __VA_ARGS__; break;           \
    default: KERNEL<int64_t> __VA_ARGS__; break;                   \
  }

}  // namespace b200rl

extern "C" int b200rl_categorical_fwd_f32(const float* logits, const uint8_t* mask, const void* actions,
                                          int act_dtype, int64_t R, int64_t n, float* logp, float* entropy,
                                          b200rl_stream_t stream) {
  using namespace b200rl;
  int rc = check_cat(logits, actions, act_dtype, R, n, "categorical_fwd");
  if (rc) return rc;
  B200RL_REQUIRE(logp && entropy, "categorical_fwd: null output");
  if (R == 0) return B200RL_OK;
  CatParams p{logits, mask, actions, R, (int)n, logp, entropy, nullptr, nullptr, nullptr};
  const unsigned grid = (unsigned)((R + kCatBlock - 1) / kCatBlock);
  DISPATCH_ACT(act_dtype, cat_fwd_kernel, <<<grid, kCatBlock, 0, (cudaStream_t)stream>>>(p));
  return check_launch("categorical_fwd");
}

extern "C" int b200rl_categorical_bwd_f32(const float* logits, const uint8_t* mask, const void* actions,
                                          int act_dtype, int64_t R, int64_t n, const float* dlogp,
                                          const float* dentropy, float* dlogits, b200rl_stream_t stream) {
  using namespace b200rl;
  int rc = check_cat(logits, actions, act_dtype, R, n, "categorical_bwd");
  if (rc) return rc;
  B200RL_REQUIRE(dlogp && dentropy && dlogits, "categorical_bwd: null pointer");
  if (R == 0) return B200RL_OK;
  CatParams p{logits, mask, actions, R, (int)n, nullptr, nullptr, dlogp, dentropy, dlogits};
  const unsigned grid = (unsigned)((R + kCatBlock - 1) / kCatBlock);
  DISPATCH_ACT(act_dtype, cat_bwd_kernel, <<<grid, kCatBlock, 0, (cudaStream_t)stream>>>(p));
  return check_launch("categorical_bwd");
}

extern "C" int b200rl_ppo_categorical_loss_f32(const float* logits, const uint8_t* mask, const void* actions,
                                               int act_dtype, int64_t B, int64_t n, const b200rl_ppo_args* args,
                                               float* dlogits, void* workspace, size_t workspace_bytes,
                                               b200rl_stream_t stream) {
  using namespace b200rl;
  int rc = check_cat(logits, actions, act_dtype, B, n, "ppo_categorical_loss");
  if (rc) return rc;
  B200RL_REQUIRE(dlogits != nullptr, "ppo_categorical_loss: dlogits is null");
  PpoDev P;
  rc = ppo_make_dev(args, B, workspace, workspace_bytes, &P);
  if (rc) return rc;
  CatParams p{logits, mask, actions, B, (int)n, nullptr, nullptr, nullptr, nullptr, dlogits};
  const unsigned grid = (unsigned)((B + kCatBlock - 1) / kCatBlock);
  cudaStream_t s = (cudaStream_t)stream;
  rc = ppo_launch_prepare(P, s);
  if (rc) return rc;
  DISPATCH_ACT(act_dtype, cat_ppo_kernel, <<<grid, kCatBlock, 0, s>>>(p, P));
  rc = check_launch("ppo_categorical_loss");
  if (rc) return rc;
  return ppo_launch_finalize(P, grid, 1, s);
}

extern "C" int b200rl_categorical_sample_f32(const float* logits, const uint8_t* mask, int64_t R, int64_t n,
                                             uint64_t seed, uint64_t offset, const int64_t* offset_dev,
                                             int64_t* actions_out, float* logp, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(logits && actions_out, "categorical_sample: null pointer");
  B200RL_REQUIRE(R >= 0 && n >= 1, "categorical_sample: bad shape");
  if (R == 0) return B200RL_OK;
  CatParams p{logits, mask, nullptr, R, (int)n, logp, nullptr, nullptr, nullptr, nullptr};
  const unsigned grid = (unsigned)((R + kCatBlock - 1) / kCatBlock);
  cat_sample_kernel<<<grid, kCatBlock, 0, (cudaStream_t)stream>>>(
      p, seed, offset, reinterpret_cast<const long long*>(offset_dev), reinterpret_cast<long long*>(actions_out));
  return check_launch("categorical_sample");
}

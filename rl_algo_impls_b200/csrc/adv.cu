// K2: per-minibatch advantage moments (f64, deterministic) and normalisation.
// Replaces mb_adv.mean(0) / mb_adv.std(0) / sub / div / @ multi_reward_weights of
// ppo/ppo.py:307-318.  The moments are (sum, sum of squares, count) so that ranks of a
// data-parallel job can all-reduce them and normalise with the global-minibatch statistics.
#include "common.cuh"

namespace b200rl {

struct AdvParams {
  const float* adv;
  const long long* idx;
  long long B;
  int V;
  int mode;
  int has_w;
  float w[B200RL_MAX_VALUE_HEADS];
  double* moments;
  double* partials;  // [blocks][2 * Vm]
  const double* moments_in;
  float* out;
  int out_v;
};

constexpr int kAdvBlock = 256;

// threads per block used for the per-head pass: a multiple of V so that a thread's head is fixed
__host__ __device__ inline int adv_tpb(int Vm) { return (kAdvBlock / Vm) * Vm; }

__global__ void __launch_bounds__(kAdvBlock) adv_moments_partial_kernel(const AdvParams p) {
  __shared__ double s_sum[kAdvBlock], s_sq[kAdvBlock];
  const int Vm = p.mode == 3 ? 1 : p.V;
  const int tpb = adv_tpb(Vm);
  const int tid = threadIdx.x;
  double sum = 0.0, sq = 0.0;
  if (tid < tpb) {
    const long long total = p.B * Vm;
    const long long stride = (long long)gridDim.x * tpb;
    for (long long e = (long long)blockIdx.x * tpb + tid; e < total; e += stride) {
      const long long i = e / Vm;
      const int v = (int)(e - i * Vm);
      const long long row = p.idx ? p.idx[i] : i;
      float a;
      if (p.mode == 3) {
        const float* r = p.adv + row * p.V;
        a = r[0];
        if (p.has_w) {
          a = 0.f;
          for (int k = 0; k < p.V; ++k) a = fmaf(r[k], p.w[k], a);
        }
      } else {
        a = p.adv[row * p.V + v];
      }
      sum += (double)a;
      sq += (double)a * (double)a;
    }
  }
  s_sum[tid] = sum, s_sq[tid] = sq;
  __syncthreads();
  if (tid < Vm) {  // fixed order over the threads that own head `tid`
    double a = 0.0, b = 0.0;
    for (int t = tid; t < tpb; t += Vm) a += s_sum[t], b += s_sq[t];
    p.partials[(long long)blockIdx.x * 2 * Vm + tid] = a;
    p.partials[(long long)blockIdx.x * 2 * Vm + Vm + tid] = b;
  }
}

__global__ void adv_moments_final_kernel(const AdvParams p, int blocks) {
  const int Vm = p.mode == 3 ? 1 : p.V;
  const int k = threadIdx.x;
  if (k < 2 * Vm) {
    double a = 0.0;
    for (int b = 0; b < blocks; ++b) a += p.partials[(long long)b * 2 * Vm + k];
    p.moments[k] = a;
  }
  if (k == 0) p.moments[2 * Vm] = (double)p.B;
}

__global__ void __launch_bounds__(kAdvBlock) adv_normalize_kernel(const AdvParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B) return;
  const long long row = p.idx ? p.idx[i] : i;
  const float* r = p.adv + row * p.V;
  const double* M = p.moments_in;
  if (p.mode == 3) {
    float a = r[0];
    if (p.has_w) {
      a = 0.f;
      for (int k = 0; k < p.V; ++k) a = fmaf(r[k], p.w[k], a);
    }
    const double n = M[2], mean = M[0] / n;
    const double var = fmax(0.0, (M[1] - M[0] * mean) / (n - 1.0));
    p.out[i] = (a - (float)mean) / ((float)sqrt(var) + 1e-8f);
    return;
  }
  float acc = 0.f;
  for (int v = 0; v < p.V; ++v) {
    float a = r[v];
    if (p.mode != 0) {
      const double n = M[2 * p.V], mean = M[v] / n;
      const double var = fmax(0.0, (M[p.V + v] - M[v] * mean) / (n - 1.0));
      const float denom = (float)sqrt(var) + 1e-8f;
      a = (p.mode == 1) ? (a - (float)mean) / denom : a / denom;
    }
    if (p.out_v == p.V) {
      p.out[i * p.V + v] = a;
    } else {
      acc = p.has_w ? fmaf(a, p.w[v], acc) : a;
    }
  }
  if (p.out_v != p.V) p.out[i] = acc;
}

static int adv_blocks(long long B, int Vm) {
  const long long per_block = adv_tpb(Vm) * 8LL;
  long long blocks = (B * Vm + per_block - 1) / per_block;
  const long long cap = 4LL * device_info().sm_count;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static int fill(AdvParams* p, const float* adv, const int64_t* idx, int64_t B, int64_t V, int mode,
                const float* weights_host, const char* who) {
  B200RL_REQUIRE(adv != nullptr, "%s: adv is null", who);
  B200RL_REQUIRE(B >= 1 && V >= 1, "%s: bad shape B=%lld V=%lld", who, (long long)B, (long long)V);
  B200RL_REQUIRE(mode >= 0 && mode <= 3, "%s: adv_mode %d", who, mode);
  B200RL_UNSUPPORTED(V > B200RL_MAX_VALUE_HEADS, "%s: V=%lld exceeds %d", who, (long long)V, B200RL_MAX_VALUE_HEADS);
  p->adv = adv, p->idx = reinterpret_cast<const long long*>(idx), p->B = B, p->V = (int)V, p->mode = mode;
  p->has_w = weights_host != nullptr;
  for (int v = 0; v < V; ++v) p->w[v] = weights_host ? weights_host[v] : 0.f;
  return B200RL_OK;
}

}  // namespace b200rl

extern "C" size_t b200rl_adv_moments_workspace_bytes(int64_t B, int64_t V) {
  (void)B;
  if (V < 1) V = 1;
  return (size_t)(4 * b200rl::device_info().sm_count) * 2 * (size_t)V * sizeof(double);
}

extern "C" int b200rl_adv_moments_f64(const float* adv, const int64_t* idx, int64_t B, int64_t V, int adv_mode,
                                      const float* weights_host, double* moments, void* workspace,
                                      size_t workspace_bytes, b200rl_stream_t stream) {
  using namespace b200rl;
  AdvParams p{};
  int rc = fill(&p, adv, idx, B, V, adv_mode, weights_host, "adv_moments");
  if (rc) return rc;
  B200RL_REQUIRE(moments && workspace, "adv_moments: null output / workspace");
  B200RL_REQUIRE(workspace_bytes >= b200rl_adv_moments_workspace_bytes(B, V), "adv_moments: workspace too small");
  const int Vm = adv_mode == 3 ? 1 : (int)V;
  const int blocks = adv_blocks(B, Vm);
  p.moments = moments, p.partials = static_cast<double*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  adv_moments_partial_kernel<<<blocks, kAdvBlock, 0, s>>>(p);
  adv_moments_final_kernel<<<1, 64, 0, s>>>(p, blocks);
  return check_launch("adv_moments");
}

extern "C" int b200rl_adv_normalize_f32(const float* adv, const int64_t* idx, int64_t B, int64_t V, int adv_mode,
                                        const float* weights_host, const double* moments, float* out,
                                        int64_t out_v, b200rl_stream_t stream) {
  using namespace b200rl;
  AdvParams p{};
  int rc = fill(&p, adv, idx, B, V, adv_mode, weights_host, "adv_normalize");
  if (rc) return rc;
  B200RL_REQUIRE(out != nullptr, "adv_normalize: out is null");
  B200RL_REQUIRE(adv_mode == 0 || moments != nullptr, "adv_normalize: moments required for mode %d", adv_mode);
  B200RL_REQUIRE(out_v == 1 || out_v == V, "adv_normalize: out_v must be 1 or V");
  B200RL_REQUIRE(!(out_v == 1 && V > 1 && !weights_host), "adv_normalize: V > 1 needs weights to contract to [B]");
  B200RL_REQUIRE(!(adv_mode == 3 && out_v != 1), "adv_normalize: mode 3 yields [B]");
  p.moments_in = moments, p.out = out, p.out_v = (int)out_v;
  const long long blocks = (B + kAdvBlock - 1) / kAdvBlock;
  adv_normalize_kernel<<<(unsigned)blocks, kAdvBlock, 0, (cudaStream_t)stream>>>(p);
  return check_launch("adv_normalize");
}

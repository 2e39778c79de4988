// K9: the channels-last glue of the squeeze U-net (Lux / MicroRTS), float32 and bfloat16 -- what sits between and
// after the cuDNN convolutions of shared/policy/actor_critic_network/squeeze_unet.py:20-196 over the SE-residual blocks
// of double_cone.py:18-86:
//
//   bias + GELU                      after every convolution that feeds a GELU (stem, down / up convolutions, the first
//                                    convolution of a residual block, the critic's strided convolutions);
//   SE tail                          out = gelu(x + (y2 + b2) * s[n, c]),  s = sigmoid(W2 gelu(W1 mean_hw(y2 + b2))):
//                                    the per-(sample, channel) mean, the gated residual sum and the block's output
//                                    activation (the two tiny linears stay cuBLAS calls on [N, C]).
//
// PyTorch runs these as a strided broadcast add, GELU, a mean reduction, a broadcast multiply, an add and a GELU --
// each a full pass (or two) over a [N, C, H, W] map -- plus their backward kernels and the bias / gate reductions:
// 68 % of the GPU time of a C5 minibatch (torch.profiler, B = 128: 40 of 59 ms in elementwise / reduce kernels against
// 16 ms of convolutions).  Here: one pass forward per op group, the backward as one reduction pass + one elementwise
// pass, reductions in float32 with a fixed order.
//
// Arithmetic: element type T (float or bfloat16) in memory, float32 in registers.  With T = bfloat16 every value that
// PyTorch materialises as a bfloat16 tensor under autocast (y + b, (y + b) * s, x + ..., the GELU output) is rounded to
// bfloat16 at the same point, so the forward reproduces the PyTorch chain bit for bit; with T = float the same float32
// operations run in the same order (no contraction across PyTorch's op boundaries).  GELU is torch's exact form
// x * 0.5 * (1 + erf(x / sqrt 2)) and its derivative cdf + x * pdf.
#include "common.cuh"

namespace b200rl {

constexpr int kActBlock = 256;

template <typename T>
struct Elems {
  static constexpr int kVec = 16 / (int)sizeof(T);
};

template <typename T>
__device__ __forceinline__ float round_to(float v) {
  return to_f32(from_f32<T>(v));
}

template <typename T, int VEC>
__device__ __forceinline__ void load_elems(const T* p, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = to_f32(p[0]);
  } else {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = to_f32(e[j]);
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void store_elems(T* p, const float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    p[0] = from_f32<T>(v[0]);
  } else {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int j = 0; j < VEC; ++j) e[j] = from_f32<T>(v[j]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
}

// torch's exact GELU and its derivative (aten/native/cuda/ActivationGeluKernel.cu), float32 math
__device__ __forceinline__ float gelu_f(float x) { return x * 0.5f * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  constexpr float kBeta = 1.12837916709551257390f * 0.70710678118654752440f * 0.5f;
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = expf(-0.5f * x * x) * kBeta;
  return cdf + x * pdf;
}

enum Act { kActNone = 0, kActRelu = 1, kActGelu = 2 };

template <int ACT>
__device__ __forceinline__ float act_f(float v) {
  if constexpr (ACT == kActRelu) return (v > 0.f || v != v) ? v : 0.f;
  if constexpr (ACT == kActGelu) return gelu_f(v);
  return v;
}

// ---- bias + activation ------------------------------------------------------------------------------------------
// out = act(rT(x + b)) over [rows, C];  backward: dx = dout * act'(rT(x + b)), x = the convolution's (bias-free) output
template <typename T, int VEC, int ACT>
__global__ void __launch_bounds__(kActBlock) bias_act_fwd_kernel(const T* x, const float* bias, T* out, long long n_vec, int Cv) {
  const long long i = (long long)blockIdx.x * kActBlock + threadIdx.x;
  if (i >= n_vec) return;
  float v[VEC];
  load_elems<T, VEC>(x + i * VEC, v);
  const int c = (int)(i % Cv) * VEC;
#pragma unroll
  for (int j = 0; j < VEC; ++j) v[j] = act_f<ACT>(round_to<T>(v[j] + round_to<T>(bias[c + j])));
  store_elems<T, VEC>(out + i * VEC, v);
}

// Grid-stride; with SUMS (kActBlock % Cv == 0, so a thread keeps its channel vector across iterations) the CTA also
// leaves the column sums of the (rounded) gradients it wrote in partial[blockIdx.x][C]: the bias gradient needs no pass
// of its own over dx.
template <typename T, int VEC, int ACT, bool SUMS>
__global__ void __launch_bounds__(kActBlock) bias_act_bwd_kernel(const T* dout, const T* x, const float* bias, T* dx,
                                                                 long long n_vec, int Cv, float* partial) {
  __shared__ float s_acc[SUMS ? kActBlock * 8 : 1];
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  const long long stride = (long long)gridDim.x * kActBlock;
  for (long long i = (long long)blockIdx.x * kActBlock + threadIdx.x; i < n_vec; i += stride) {
    float g[VEC], v[VEC];
    load_elems<T, VEC>(dout + i * VEC, g);
    load_elems<T, VEC>(x + i * VEC, v);
    const int c = (int)(i % Cv) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float pre = round_to<T>(v[j] + round_to<T>(bias[c + j]));
      if constexpr (ACT == kActGelu) g[j] = g[j] * gelu_grad_f(pre);
      if constexpr (ACT == kActRelu) g[j] = pre <= 0.f ? 0.f : g[j];
      if constexpr (SUMS) acc[j] += round_to<T>(g[j]);
    }
    store_elems<T, VEC>(dx + i * VEC, g);
  }
  if constexpr (SUMS) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) s_acc[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    if ((int)threadIdx.x < Cv) {
      for (int q = 1; q < kActBlock / Cv; ++q)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += s_acc[(q * Cv + threadIdx.x) * 8 + j];
#pragma unroll
      for (int j = 0; j < VEC; ++j) partial[(long long)blockIdx.x * Cv * VEC + threadIdx.x * VEC + j] = acc[j];
    }
  }
}

// ---- column sums of a [rows, C] map of T (bias gradients), float32, two stages, fixed order -------------------------
struct ColT {
  const void* g;
  float* partial;  // [slabs][C]
  long long rows, rows_per_slab;
  int C, tx;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kActBlock) colsum_t_partial_kernel(const ColT G) {
  __shared__ float s_acc[kActBlock * 8];
  const int Cv = G.C / VEC;
  const int tx = G.tx, ty = kActBlock / tx;
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int cv = blockIdx.x * tx + cx;
  const long long r0 = (long long)blockIdx.y * G.rows_per_slab;
  const long long r1 = r0 + G.rows_per_slab < G.rows ? r0 + G.rows_per_slab : G.rows;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (cv < Cv)
    for (long long r = r0 + ry; r < r1; r += ty) {
      float g[VEC];
      load_elems<T, VEC>(static_cast<const T*>(G.g) + r * G.C + (long long)cv * VEC, g);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += g[j];
    }
#pragma unroll
  for (int j = 0; j < VEC; ++j) s_acc[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if (ry == 0 && cv < Cv) {
    for (int q = 1; q < ty; ++q)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += s_acc[(q * tx + cx) * 8 + j];
#pragma unroll
    for (int j = 0; j < VEC; ++j) G.partial[(long long)blockIdx.y * G.C + (long long)cv * VEC + j] = acc[j];
  }
}

// out[c] (+)= sum over slabs; a CTA owns 32 channels, thread (cx, sy) adds slabs sy, sy + 8, ...
__global__ void __launch_bounds__(kActBlock) colsum_t_final_kernel(const float* partial, float* out, int slabs, int C,
                                                                   const float* add /*nullable [C]*/) {
  __shared__ float s_part[kActBlock];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (c < C)
    for (int s = sy; s < slabs; s += kActBlock / 32) acc += partial[(long long)s * C + c];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (sy == 0 && c < C) {
    for (int q = 1; q < kActBlock / 32; ++q) acc += s_part[q * 32 + cx];
    out[c] = add ? acc + add[c] : acc;
  }
}

static void col_plan(long long rows, int C, int vec, int* tx_out, int* slabs_out) {
  const int Cv = C / vec;
  int tx = 1;
  while (tx < Cv && tx < 32) tx *= 2;
  const int tiles = (Cv + tx - 1) / tx;
  const int ty = kActBlock / tx;
  long long want = (long long)device_info().sm_count * 2 / tiles;
  const long long cap = (rows + ty * 8 - 1) / (ty * 8);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 4096) want = 4096;
  *tx_out = tx, *slabs_out = (int)want;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static int launch_colsum_t(const T* g, float* out, void* ws, size_t ws_bytes, long long rows, int C, const float* add,
                           cudaStream_t stream, const char* who) {
  B200RL_REQUIRE(ws != nullptr, "%s: workspace is null", who);
  float* partial = static_cast<float*>(ws);
  partial += ((16 - (reinterpret_cast<uintptr_t>(partial) & 15)) & 15) / sizeof(float);
  const size_t usable = ws_bytes - (size_t)(reinterpret_cast<uint8_t*>(partial) - static_cast<uint8_t*>(ws));
  constexpr int V = Elems<T>::kVec;
  const bool vec = C % V == 0 && al16(g);
  int tx, slabs;
  col_plan(rows, C, vec ? V : 1, &tx, &slabs);
  B200RL_REQUIRE(usable >= (size_t)slabs * C * sizeof(float), "%s: workspace too small (%zu < %zu bytes)", who, usable,
                 (size_t)slabs * C * sizeof(float));
  const int Cv = C / (vec ? V : 1);
  ColT G{g, partial, rows, (rows + slabs - 1) / slabs, C, tx};
  const dim3 grid((unsigned)((Cv + tx - 1) / tx), (unsigned)slabs);
  if (vec) colsum_t_partial_kernel<T, V><<<grid, kActBlock, 0, stream>>>(G);
  else colsum_t_partial_kernel<T, 1><<<grid, kActBlock, 0, stream>>>(G);
  int rc = check_launch(who);
  if (rc) return rc;
  colsum_t_final_kernel<<<(C + 31) / 32, kActBlock, 0, stream>>>(partial, out, slabs, C, add);
  return check_launch(who);
}

// ---- SE tail ------------------------------------------------------------------------------------------------------
// Per-(sample, channel) sums over the HW rows of a [N, HW, C] map, float32: thread (cx, ry) of CTA (channel tile, n)
// walks rows ry, ry + TY, ...; partials fold in ry order.  MODE 0: sum of v (the SE mean's numerator; the caller adds
// the bias and divides).  MODE 1: sum of dz * u with u = rT(y2 + b2), z = rT(x + rT(u * s)), dz = dout * gelu'(z)
// (the gradient of the gate s).
struct SeDev {
  const void* x;     // block input [N, HW, C]
  const void* y2;    // second convolution's output, bias-free
  const float* b2;   // [C]
  const void* s;     // gate [N, C] (T)
  const void* dout;  // [N, HW, C]
  void* out;         // forward: block output; backward: dx
  void* dy2;         // backward
  const float* dmean;  // backward: d loss / d mean [N, C] float32, already divided by HW
  float* sums;       // MODE 0 / 1 output [N, C] (one slab) or partials [N, slabs, C]
  long long N, HW;
  int C, tx;
  int slabs;         // the HW rows of a sample are cut into this many slabs, one CTA each (grid.z)
  long long rows_per_slab;
};

template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(kActBlock) se_rowsum_kernel(const SeDev G) {
  __shared__ float s_acc[kActBlock * 8];
  const int Cv = G.C / VEC;
  const int tx = G.tx, ty = kActBlock / tx;
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int cv = blockIdx.x * tx + cx;
  const long long n = blockIdx.y;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (cv < Cv) {
    const int c = cv * VEC;
    float b[VEC], sg[VEC];
    if constexpr (MODE == 1) {
      load_elems<T, VEC>(static_cast<const T*>(G.s) + n * G.C + c, sg);
#pragma unroll
      for (int j = 0; j < VEC; ++j) b[j] = round_to<T>(G.b2[c + j]);
    }
    const long long r_lo = (long long)blockIdx.z * G.rows_per_slab;
    const long long r_hi = r_lo + G.rows_per_slab < G.HW ? r_lo + G.rows_per_slab : G.HW;
    for (long long r = r_lo + ry; r < r_hi; r += ty) {
      const long long o = (n * G.HW + r) * G.C + c;
      float v[VEC];
      load_elems<T, VEC>(static_cast<const T*>(G.y2) + o, v);
      if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += v[j];
      } else {
        float xv[VEC], g[VEC];
        load_elems<T, VEC>(static_cast<const T*>(G.x) + o, xv);
        load_elems<T, VEC>(static_cast<const T*>(G.dout) + o, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float u = round_to<T>(v[j] + b[j]);
          const float z = round_to<T>(xv[j] + round_to<T>(u * sg[j]));
          g[j] = g[j] * gelu_grad_f(z);  // dz: the gradient of the block's input through the residual sum, as it is
          acc[j] += g[j] * u;
        }
        store_elems<T, VEC>(static_cast<T*>(G.out) + o, g);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) s_acc[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if (ry == 0 && cv < Cv) {
    for (int q = 1; q < ty; ++q)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += s_acc[(q * tx + cx) * 8 + j];
#pragma unroll
    for (int j = 0; j < VEC; ++j) G.sums[(n * G.slabs + blockIdx.z) * G.C + (long long)cv * VEC + j] = acc[j];
  }
}

// out[n, c] = sum over the slabs, in slab order
__global__ void __launch_bounds__(kActBlock) se_fold_kernel(const float* partial, float* out, long long NC, int C, int slabs) {
  const long long i = (long long)blockIdx.x * kActBlock + threadIdx.x;
  if (i >= NC) return;
  const long long n = i / C;
  const int c = (int)(i - n * C);
  float acc = 0.f;
  for (int z = 0; z < slabs; ++z) acc += partial[(n * slabs + z) * C + c];
  out[i] = acc;
}

// forward: out = rT(gelu(rT(x + rT(rT(y2 + b2) * s))));  backward (second pass): dy2 = rT(dz * s + dmean) from the dz the
// gate-gradient pass left in dx
template <typename T, int VEC, bool BWD>
__global__ void __launch_bounds__(kActBlock) se_apply_kernel(const SeDev G) {
  const int Cv = G.C / VEC;
  const long long n_vec = G.N * G.HW * Cv;
  const long long i = (long long)blockIdx.x * kActBlock + threadIdx.x;
  if (i >= n_vec) return;
  const int c = (int)(i % Cv) * VEC;
  const long long n = (i / Cv) / G.HW;
  float sg[VEC];
  load_elems<T, VEC>(static_cast<const T*>(G.s) + n * G.C + c, sg);
  if constexpr (!BWD) {
    float xv[VEC], v[VEC], z[VEC];
    load_elems<T, VEC>(static_cast<const T*>(G.x) + i * VEC, xv);
    load_elems<T, VEC>(static_cast<const T*>(G.y2) + i * VEC, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float u = round_to<T>(v[j] + round_to<T>(G.b2[c + j]));
      z[j] = gelu_f(round_to<T>(xv[j] + round_to<T>(u * sg[j])));
    }
    store_elems<T, VEC>(static_cast<T*>(G.out) + i * VEC, z);
  } else {
    float g[VEC];
    load_elems<T, VEC>(static_cast<const T*>(G.dout) + i * VEC, g);  // dz
#pragma unroll
    for (int j = 0; j < VEC; ++j) g[j] = g[j] * sg[j] + G.dmean[n * G.C + c + j];
    store_elems<T, VEC>(static_cast<T*>(G.dy2) + i * VEC, g);
  }
}

template <typename T>
static int bias_act_fwd_t(const T* x, const float* bias, T* out, long long rows, int C, int act, cudaStream_t st) {
  constexpr int V = Elems<T>::kVec;
  const bool vec = C % V == 0 && al16(x) && al16(out);
  const long long n_vec = rows * C / (vec ? V : 1);
  B200RL_UNSUPPORTED((n_vec + kActBlock - 1) / kActBlock > 0x7fffffffLL, "nhwc_bias_act_fwd: tensor too large");
  const unsigned grid = (unsigned)((n_vec + kActBlock - 1) / kActBlock);
  const int Cv = C / (vec ? V : 1);
#define B200RL_LAUNCH_ACT(KERNEL, ...)                                                                     \
  do {                                                                                                     \
    if (vec) {                                                                                             \
      if (act == kActGelu) KERNEL<T, V, kActGelu><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);                \
      else if (act == kActRelu) KERNEL<T, V, kActRelu><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);           \
      else KERNEL<T, V, kActNone><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);                                \
    } else {                                                                                               \
      if (act == kActGelu) KERNEL<T, 1, kActGelu><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);                \
      else if (act == kActRelu) KERNEL<T, 1, kActRelu><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);           \
      else KERNEL<T, 1, kActNone><<<grid, kActBlock, 0, st>>>(__VA_ARGS__);                                \
    }                                                                                                      \
  } while (0)
  B200RL_LAUNCH_ACT(bias_act_fwd_kernel, x, bias, out, n_vec, Cv);
#undef B200RL_LAUNCH_ACT
  return check_launch("nhwc_bias_act_fwd");
}

template <typename T>
static int bias_act_bwd_t(const T* dout, const T* x, const float* bias, T* dx, float* dbias, void* ws, size_t ws_bytes,
                          long long rows, int C, int act, cudaStream_t st) {
  constexpr int V = Elems<T>::kVec;
  const bool vec = C % V == 0 && al16(dout) && al16(x) && al16(dx);
  const int Cv = C / (vec ? V : 1);
  const long long n_vec = rows * Cv;
  const long long full = (n_vec + kActBlock - 1) / kActBlock;
  B200RL_UNSUPPORTED(full > 0x7fffffffLL, "nhwc_bias_act_bwd: tensor too large");
  // the bias gradient rides on the same pass when a thread's channel vector is fixed across its iterations
  const bool sums = dbias != nullptr && vec && Cv <= kActBlock && kActBlock % Cv == 0;
  float* partial = nullptr;
  unsigned grid = (unsigned)full;
  if (sums) {
    const long long cap = (long long)device_info().sm_count * 4;
    grid = (unsigned)(full < cap ? full : cap);
    B200RL_REQUIRE(ws != nullptr, "nhwc_bias_act_bwd: workspace is null");
    partial = static_cast<float*>(ws);
    partial += ((16 - (reinterpret_cast<uintptr_t>(partial) & 15)) & 15) / sizeof(float);
    const size_t usable = ws_bytes - (size_t)(reinterpret_cast<uint8_t*>(partial) - static_cast<uint8_t*>(ws));
    B200RL_REQUIRE(usable >= (size_t)grid * C * sizeof(float), "nhwc_bias_act_bwd: workspace too small (%zu < %zu bytes)",
                   usable, (size_t)grid * C * sizeof(float));
  }
#define B200RL_LAUNCH_BWD(VV, SS)                                                                                          \
  do {                                                                                                                     \
    if (act == kActGelu) bias_act_bwd_kernel<T, VV, kActGelu, SS><<<grid, kActBlock, 0, st>>>(dout, x, bias, dx, n_vec, Cv, partial); \
    else if (act == kActRelu) bias_act_bwd_kernel<T, VV, kActRelu, SS><<<grid, kActBlock, 0, st>>>(dout, x, bias, dx, n_vec, Cv, partial); \
    else bias_act_bwd_kernel<T, VV, kActNone, SS><<<grid, kActBlock, 0, st>>>(dout, x, bias, dx, n_vec, Cv, partial);     \
  } while (0)
  if (sums) B200RL_LAUNCH_BWD(V, true);
  else if (vec) B200RL_LAUNCH_BWD(V, false);
  else B200RL_LAUNCH_BWD(1, false);
#undef B200RL_LAUNCH_BWD
  int rc = check_launch("nhwc_bias_act_bwd");
  if (rc || !dbias) return rc;
  if (sums) {
    colsum_t_final_kernel<<<(C + 31) / 32, kActBlock, 0, st>>>(partial, dbias, (int)grid, C, nullptr);
    return check_launch("nhwc_bias_act_bwd (bias gradient)");
  }
  return launch_colsum_t<T>(dx, dbias, ws, ws_bytes, rows, C, nullptr, st, "nhwc_bias_act_bwd (bias gradient)");
}

// slabs per sample: enough CTAs to fill the machine (~4 per SM over all samples and channel tiles), >= 8 rows per thread
static int se_slabs(long long N, long long HW, int C) {
  const int Cv = C % 8 == 0 ? C / 8 : (C % 4 == 0 ? C / 4 : C);
  int tx = 1;
  while (tx < Cv && tx < 32) tx *= 2;
  const long long tiles = (Cv + tx - 1) / tx, ty = kActBlock / tx;
  long long want = ((long long)device_info().sm_count * 4 + N * tiles - 1) / (N * tiles > 0 ? N * tiles : 1);
  const long long cap = (HW + ty * 8 - 1) / (ty * 8);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

template <typename T>
static int se_launch(SeDev& G, int mode /*0 mean sums, 1 gate gradient, 2 forward, 3 backward*/, cudaStream_t st,
                     void* ws = nullptr, size_t ws_bytes = 0) {
  constexpr int V = Elems<T>::kVec;
  const bool vec = G.C % V == 0 && al16(G.y2) && (!G.x || al16(G.x)) && (!G.s || al16(G.s)) && (!G.dout || al16(G.dout)) &&
                   (!G.out || al16(G.out)) && (!G.dy2 || al16(G.dy2));
  const int Cv = G.C / (vec ? V : 1);
  if (mode <= 1) {
    int tx = 1;
    while (tx < Cv && tx < 32) tx *= 2;
    G.tx = tx;
    B200RL_UNSUPPORTED(G.N > 65535, "se_tail: N=%lld samples exceed 65535", G.N);
    G.slabs = se_slabs(G.N, G.HW, G.C);
    G.rows_per_slab = (G.HW + G.slabs - 1) / G.slabs;
    float* out = G.sums;
    if (G.slabs > 1) {  // partials in the workspace, folded in slab order below
      B200RL_REQUIRE(ws != nullptr && ws_bytes >= (size_t)G.N * G.slabs * G.C * sizeof(float) + 16,
                     "se_tail: workspace too small (%zu bytes)", ws_bytes);
      float* partial = static_cast<float*>(ws);
      partial += ((16 - (reinterpret_cast<uintptr_t>(partial) & 15)) & 15) / sizeof(float);
      G.sums = partial;
    }
    const dim3 grid((unsigned)((Cv + tx - 1) / tx), (unsigned)G.N, (unsigned)G.slabs);
    if (vec && mode == 0) se_rowsum_kernel<T, V, 0><<<grid, kActBlock, 0, st>>>(G);
    else if (vec) se_rowsum_kernel<T, V, 1><<<grid, kActBlock, 0, st>>>(G);
    else if (mode == 0) se_rowsum_kernel<T, 1, 0><<<grid, kActBlock, 0, st>>>(G);
    else se_rowsum_kernel<T, 1, 1><<<grid, kActBlock, 0, st>>>(G);
    int rc = check_launch("se_tail (row sums)");
    if (rc || G.slabs == 1) return rc;
    const long long NC = G.N * G.C;
    se_fold_kernel<<<(unsigned)((NC + kActBlock - 1) / kActBlock), kActBlock, 0, st>>>(G.sums, out, NC, G.C, G.slabs);
    return check_launch("se_tail (fold)");
  }
  const long long n_vec = G.N * G.HW * Cv;
  B200RL_UNSUPPORTED((n_vec + kActBlock - 1) / kActBlock > 0x7fffffffLL, "se_tail: tensor too large");
  const unsigned grid = (unsigned)((n_vec + kActBlock - 1) / kActBlock);
  if (vec && mode == 2) se_apply_kernel<T, V, false><<<grid, kActBlock, 0, st>>>(G);
  else if (vec) se_apply_kernel<T, V, true><<<grid, kActBlock, 0, st>>>(G);
  else if (mode == 2) se_apply_kernel<T, 1, false><<<grid, kActBlock, 0, st>>>(G);
  else se_apply_kernel<T, 1, true><<<grid, kActBlock, 0, st>>>(G);
  return check_launch("se_tail (apply)");
}

}  // namespace b200rl

extern "C" size_t b200rl_nhwc_bias_act_workspace_bytes(int64_t rows, int64_t C) {
  if (rows < 1 || C < 1) return 32;
  int tx, best = 0;
  for (int vec : {8, 4, 1}) {
    if (C % vec) continue;
    int slabs;
    b200rl::col_plan(rows, (int)C, vec, &tx, &slabs);
    best = slabs > best ? slabs : best;
  }
  const int fused = b200rl::device_info().sm_count * 4;  // the backward's own per-CTA partials
  best = fused > best ? fused : best;
  return (size_t)best * (size_t)C * sizeof(float) + 32;
}

extern "C" int b200rl_nhwc_bias_act_fwd(const void* x, const float* bias, void* out, int64_t rows, int64_t C, int act,
                                        int dtype, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rows >= 0 && C >= 1 && C <= (1 << 20), "nhwc_bias_act_fwd: bad shape");
  B200RL_REQUIRE(act >= kActNone && act <= kActGelu, "nhwc_bias_act_fwd: activation %d", act);
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "nhwc_bias_act_fwd: dtype %d", dtype);
  if (rows == 0) return B200RL_OK;
  B200RL_REQUIRE(x && bias && out, "nhwc_bias_act_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200RL_BF16)
    return bias_act_fwd_t<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), bias, static_cast<__nv_bfloat16*>(out), rows, (int)C, act, st);
  return bias_act_fwd_t<float>(static_cast<const float*>(x), bias, static_cast<float*>(out), rows, (int)C, act, st);
}

extern "C" int b200rl_nhwc_bias_act_bwd(const void* dout, const void* x, const float* bias, void* dx, float* dbias,
                                        void* workspace, size_t workspace_bytes, int64_t rows, int64_t C, int act, int dtype,
                                        b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rows >= 0 && C >= 1 && C <= (1 << 20), "nhwc_bias_act_bwd: bad shape");
  B200RL_REQUIRE(act >= kActNone && act <= kActGelu, "nhwc_bias_act_bwd: activation %d", act);
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "nhwc_bias_act_bwd: dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), st);
    return check_launch("nhwc_bias_act_bwd");
  }
  B200RL_REQUIRE(dout && x && bias && dx, "nhwc_bias_act_bwd: null pointer");
  if (dtype == B200RL_BF16)
    return bias_act_bwd_t<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(x), bias,
                                         static_cast<__nv_bfloat16*>(dx), dbias, workspace, workspace_bytes, rows, (int)C, act, st);
  return bias_act_bwd_t<float>(static_cast<const float*>(dout), static_cast<const float*>(x), bias, static_cast<float*>(dx),
                               dbias, workspace, workspace_bytes, rows, (int)C, act, st);
}

extern "C" size_t b200rl_se_workspace_bytes(int64_t N, int64_t HW, int64_t C) {
  if (N < 1 || HW < 1 || C < 1) return 32;
  return (size_t)N * 64 * (size_t)C * sizeof(float) + 32;  // at most 64 slabs per sample
}

extern "C" int b200rl_se_mean_sums(const void* y2, float* sums, void* workspace, size_t workspace_bytes, int64_t N,
                                   int64_t HW, int64_t C, int dtype, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(N >= 0 && HW >= 1 && C >= 1 && C <= (1 << 20), "se_mean_sums: bad shape");
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "se_mean_sums: dtype %d", dtype);
  if (N == 0) return B200RL_OK;
  B200RL_REQUIRE(y2 && sums, "se_mean_sums: null pointer");
  SeDev G{};
  G.y2 = y2, G.sums = sums, G.N = N, G.HW = HW, G.C = (int)C;
  return dtype == B200RL_BF16 ? se_launch<__nv_bfloat16>(G, 0, (cudaStream_t)stream, workspace, workspace_bytes)
                              : se_launch<float>(G, 0, (cudaStream_t)stream, workspace, workspace_bytes);
}

extern "C" int b200rl_se_tail_fwd(const void* x, const void* y2, const float* b2, const void* gate, void* out, int64_t N,
                                  int64_t HW, int64_t C, int dtype, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(N >= 0 && HW >= 1 && C >= 1 && C <= (1 << 20), "se_tail_fwd: bad shape");
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "se_tail_fwd: dtype %d", dtype);
  if (N == 0) return B200RL_OK;
  B200RL_REQUIRE(x && y2 && b2 && gate && out, "se_tail_fwd: null pointer");
  SeDev G{};
  G.x = x, G.y2 = y2, G.b2 = b2, G.s = gate, G.out = out, G.N = N, G.HW = HW, G.C = (int)C;
  return dtype == B200RL_BF16 ? se_launch<__nv_bfloat16>(G, 2, (cudaStream_t)stream) : se_launch<float>(G, 2, (cudaStream_t)stream);
}

extern "C" int b200rl_se_tail_gate_grad(const void* dout, const void* x, const void* y2, const float* b2, const void* gate,
                                        float* dgate, void* dx, void* workspace, size_t workspace_bytes, int64_t N, int64_t HW,
                                        int64_t C, int dtype, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(N >= 0 && HW >= 1 && C >= 1 && C <= (1 << 20), "se_tail_gate_grad: bad shape");
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "se_tail_gate_grad: dtype %d", dtype);
  if (N == 0) return B200RL_OK;
  B200RL_REQUIRE(dout && x && y2 && b2 && gate && dgate && dx, "se_tail_gate_grad: null pointer");
  SeDev G{};
  G.dout = dout, G.x = x, G.y2 = y2, G.b2 = b2, G.s = gate, G.sums = dgate, G.out = dx, G.N = N, G.HW = HW, G.C = (int)C;
  return dtype == B200RL_BF16 ? se_launch<__nv_bfloat16>(G, 1, (cudaStream_t)stream, workspace, workspace_bytes)
                              : se_launch<float>(G, 1, (cudaStream_t)stream, workspace, workspace_bytes);
}

extern "C" int b200rl_se_tail_bwd(const void* dz, const void* gate, const float* dmean, void* dy2, float* db2, void* workspace,
                                  size_t workspace_bytes, int64_t N, int64_t HW, int64_t C, int dtype, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(N >= 0 && HW >= 1 && C >= 1 && C <= (1 << 20), "se_tail_bwd: bad shape");
  B200RL_UNSUPPORTED(dtype != B200RL_F32 && dtype != B200RL_BF16, "se_tail_bwd: dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    if (db2) cudaMemsetAsync(db2, 0, (size_t)C * sizeof(float), st);
    return check_launch("se_tail_bwd");
  }
  B200RL_REQUIRE(dz && gate && dmean && dy2, "se_tail_bwd: null pointer");
  SeDev G{};
  G.dout = dz, G.s = gate, G.dmean = dmean, G.dy2 = dy2, G.y2 = dy2;  // (y2 only feeds the alignment test of the launcher)
  G.N = N, G.HW = HW, G.C = (int)C;
  int rc = dtype == B200RL_BF16 ? se_launch<__nv_bfloat16>(G, 3, st) : se_launch<float>(G, 3, st);
  if (rc || !db2) return rc;
  if (dtype == B200RL_BF16)
    return launch_colsum_t<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(dy2), db2, workspace, workspace_bytes, N * HW, (int)C,
                                          nullptr, st, "se_tail_bwd (bias gradient)");
  return launch_colsum_t<float>(static_cast<const float*>(dy2), db2, workspace, workspace_bytes, N * HW, (int)C, nullptr, st,
                                "se_tail_bwd (bias gradient)");
}

// K8: the channels-last glue between the convolutions of the GridNet encoder / decoder -- bias + max-pool + ReLU
// and bias + ReLU, forward and backward, float32.
//
// Replaces, between the cuDNN convolutions of shared/encoder/gridnet_encoder.py:26-51 (conv -> MaxPool2d(3, 2, 1) ->
// ReLU, four times) and shared/actor/gridnet_decoder.py:36-53 (transposed conv -> ReLU, three times), what PyTorch
// runs as separate launches on the channels-last tensors the path hands the trunk: the bias add (a strided
// elementwise kernel), max_pool_forward_nhwc / max_pool_backward_nhwc, the ReLU clamp / threshold_backward and the
// bias-gradient reduction.  Measured in a C4 step (profiles/r02/launches_bench_C4_summary.csv): 35 % of the step's
// GPU time in those kernels, 8.7-14 us per pooling launch on a 24-sample rollout tensor.
//
// Same arithmetic as the PyTorch sequence, in the same order per element: v = x + bias (one float32 add), window
// maximum with torch's rule (`v > max || isnan(v)` scanning kh, then kw: the first maximum wins, NaN propagates),
// ReLU as `v > 0 ? v : 0` (NaN kept) -> the forward is bit-identical.  The backward routes dout to the arg-max
// position where the output is positive; an input position gathers its <= ceil(k/s)^2 windows in ascending (ho, wo)
// order (no atomics, deterministic; torch accumulates the same terms in its own order).  d bias = column sums of
// the routed gradient over (n, ho, wo): per-slab partial sums, then one pass over the slabs in slab order.
//
// Layout: x [N, H, W, C], out / argmax / dout [N, Ho, Wo, C], C innermost; 128-bit accesses when C % 4 == 0.
// All kernels are HBM streaming kernels: one read of x (9/4 window overlap served by L1/L2) + one write of out.
#include "common.cuh"

namespace b200rl {

constexpr int kPoolBlock = 256;
constexpr uint8_t kDead = 255;  // argmax code of an output the ReLU zeroed: no gradient

struct PoolDev {
  const float* x;
  const float* bias;
  float* out;
  uint8_t* argmax;
  const float* dout;
  float* dx;
  long long N;
  int H, W, C, Ho, Wo, k, s, p;
  int relu;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *p = v[0];
}
template <int VEC>
__device__ __forceinline__ void load_codes(const uint8_t* p, uint8_t (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const uchar4 t = *reinterpret_cast<const uchar4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void store_codes(uint8_t* p, const uint8_t (&v)[VEC]) {
  if constexpr (VEC == 4) *reinterpret_cast<uchar4*>(p) = make_uchar4(v[0], v[1], v[2], v[3]);
  else *p = v[0];
}

// ---- forward: one thread per (n, ho, wo, VEC channels) --------------------------------------------
// IDX: unsigned when every element index fits 31 bits (64-bit divisions and address arithmetic were a third of the
// instructions), long long otherwise
template <int VEC, int K, typename IDX>  // K > 0: compile-time window (3: the encoder's), 0: runtime G.k
__global__ void __launch_bounds__(kPoolBlock) pool_fwd_kernel(const PoolDev G) {
  const IDX Cv = (IDX)(G.C / VEC);
  const IDX total = (IDX)G.N * (IDX)G.Ho * (IDX)G.Wo * Cv;
  const IDX i = (IDX)blockIdx.x * kPoolBlock + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % Cv);
  IDX t = i / Cv;
  const int wo = (int)(t % (IDX)G.Wo);
  t /= (IDX)G.Wo;
  const int ho = (int)(t % (IDX)G.Ho);
  const IDX n = t / (IDX)G.Ho;
  const int k = K > 0 ? K : G.k;
  const int h0 = ho * G.s - G.p, w0 = wo * G.s - G.p;
  const int c = cv * VEC;
  float b[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) b[j] = 0.f;
  if (G.bias) load_vec<VEC>(G.bias + c, b);
  float m[VEC];
  uint8_t am[VEC];
  const int kh_first = h0 < 0 ? -h0 : 0, kw_first = w0 < 0 ? -w0 : 0;  // torch starts the arg-max at the first in-bounds entry
#pragma unroll
  for (int j = 0; j < VEC; ++j) m[j] = -INFINITY, am[j] = (uint8_t)(kh_first * k + kw_first);
  const float* xn = G.x + n * (IDX)G.H * (IDX)G.W * (IDX)G.C + c;
  if constexpr (K > 0) {
    // compile-time window: every load of the window is issued before the first compare (out-of-range taps re-read a
    // clamped position and are skipped by the scan), instead of one round trip per tap
    float v[K * K][VEC];
    bool ok[K * K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int h = h0 + kh, w = w0 + kw;
        ok[kh * K + kw] = h >= 0 && h < G.H && w >= 0 && w < G.W;
        const int hc = h < 0 ? 0 : (h >= G.H ? G.H - 1 : h), wc = w < 0 ? 0 : (w >= G.W ? G.W - 1 : w);
        load_vec<VEC>(xn + ((IDX)hc * (IDX)G.W + (IDX)wc) * (IDX)G.C, v[kh * K + kw]);
      }
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      if (!ok[t]) continue;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float vb = v[t][j] + b[j];
        if (vb > m[j] || vb != vb) m[j] = vb, am[j] = (uint8_t)t;
      }
    }
  } else {
    for (int kh = 0; kh < k; ++kh) {
      const int h = h0 + kh;
      if (h < 0 || h >= G.H) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int w = w0 + kw;
        if (w < 0 || w >= G.W) continue;
        float v[VEC];
        load_vec<VEC>(xn + ((long long)h * G.W + w) * G.C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float vb = v[j] + b[j];
          if (vb > m[j] || vb != vb) m[j] = vb, am[j] = (uint8_t)(kh * k + kw);
        }
      }
    }
  }
  if (G.relu) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const bool live = m[j] > 0.f || m[j] != m[j];
      if (!live) m[j] = 0.f, am[j] = kDead;
    }
  }
  const IDX o = ((n * (IDX)G.Ho + (IDX)ho) * (IDX)G.Wo + (IDX)wo) * (IDX)G.C + (IDX)c;
  store_vec<VEC>(G.out + o, m);
  if (G.argmax) store_codes<VEC>(G.argmax + o, am);
}

// ---- backward, kernel 3 / stride 2 / padding 1 (the encoder's): one thread per 2 x 2 input patch x VEC channels ------
// Rows 2m, 2m+1 and columns 2q, 2q+1 are covered by windows ho in {m, m+1}, wo in {q, q+1} only (an even row by window
// m alone), so the patch's four pixels share four (code, gradient) loads -- the gather form below reads four per pixel.
// Same terms in the same (ho, wo) order per pixel: bit-identical to the gather form.
template <int VEC, typename IDX>
__global__ void __launch_bounds__(kPoolBlock) pool_bwd_k3s2_kernel(const PoolDev G) {
  const IDX Cv = (IDX)(G.C / VEC);
  const IDX Hm = (IDX)((G.H + 1) / 2), Wm = (IDX)((G.W + 1) / 2);
  const IDX total = (IDX)G.N * Hm * Wm * Cv;
  const IDX i = (IDX)blockIdx.x * kPoolBlock + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % Cv);
  IDX t = i / Cv;
  const int q = (int)(t % Wm);
  t /= Wm;
  const int m = (int)(t % Hm);
  const IDX n = t / Hm;
  const int c = cv * VEC;
  uint8_t am[4][VEC];
  float g[4][VEC];
  bool ok[4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      ok[a * 2 + b] = m + a < G.Ho && q + b < G.Wo;
      const int ho = m + a < G.Ho ? m + a : G.Ho - 1, wo = q + b < G.Wo ? q + b : G.Wo - 1;
      const IDX o = ((n * (IDX)G.Ho + (IDX)ho) * (IDX)G.Wo + (IDX)wo) * (IDX)G.C + (IDX)c;
      load_codes<VEC>(G.argmax + o, am[a * 2 + b]);
      load_vec<VEC>(G.dout + o, g[a * 2 + b]);
    }
#pragma unroll
  for (int dh = 0; dh < 2; ++dh)
#pragma unroll
    for (int dw = 0; dw < 2; ++dw) {
      const int h = 2 * m + dh, w = 2 * q + dw;
      if (h >= G.H || w >= G.W) continue;
      float acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
      for (int a = 0; a <= dh; ++a)
#pragma unroll
        for (int b = 0; b <= dw; ++b) {
          if (!ok[a * 2 + b]) continue;
          const uint8_t code = (uint8_t)((dh + 1 - 2 * a) * 3 + (dw + 1 - 2 * b));
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (am[a * 2 + b][j] == code) acc[j] += g[a * 2 + b][j];
        }
      store_vec<VEC>(G.dx + ((n * (IDX)G.H + (IDX)h) * (IDX)G.W + (IDX)w) * (IDX)G.C + (IDX)c, acc);
    }
}

// ---- backward: one thread per (n, h, w, VEC channels) gathers the windows that cover it ------------------
template <int VEC>
__global__ void __launch_bounds__(kPoolBlock) pool_bwd_kernel(const PoolDev G) {
  const int Cv = G.C / VEC;
  const long long total = G.N * G.H * G.W * Cv;
  const long long i = (long long)blockIdx.x * kPoolBlock + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % Cv);
  long long t = i / Cv;
  const int w = (int)(t % G.W);
  t /= G.W;
  const int h = (int)(t % G.H);
  const long long n = t / G.H;
  const int c = cv * VEC;
  // windows ho with ho*s - p <= h <= ho*s - p + k - 1
  const int hp = h + G.p, wp = w + G.p;
  int ho_lo = hp - G.k + 1 <= 0 ? 0 : (hp - G.k + 1 + G.s - 1) / G.s;
  int wo_lo = wp - G.k + 1 <= 0 ? 0 : (wp - G.k + 1 + G.s - 1) / G.s;
  int ho_hi = hp / G.s, wo_hi = wp / G.s;
  if (ho_hi > G.Ho - 1) ho_hi = G.Ho - 1;
  if (wo_hi > G.Wo - 1) wo_hi = G.Wo - 1;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  const int nh = ho_hi - ho_lo + 1, nw = wo_hi - wo_lo + 1;
  if (nh <= 2 && nw <= 2) {
    // at most 2 x 2 covering windows (kernel <= 2 * stride: the encoder's 3 / 2): their codes and gradients are all
    // loaded before the first compare (absent windows re-read a clamped one and are skipped), same order of the sum
    uint8_t am[4][VEC];
    float g[4][VEC];
    bool ok[4];
    uint8_t code[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        ok[a * 2 + b] = a < nh && b < nw;
        int ho = ho_lo + a, wo = wo_lo + b;
        ho = ho > G.Ho - 1 ? G.Ho - 1 : ho, wo = wo > G.Wo - 1 ? G.Wo - 1 : wo;
        code[a * 2 + b] = (uint8_t)((hp - ho * G.s) * G.k + (wp - wo * G.s));
        const long long o = ((n * G.Ho + ho) * G.Wo + wo) * G.C + c;
        load_codes<VEC>(G.argmax + o, am[a * 2 + b]);
        load_vec<VEC>(G.dout + o, g[a * 2 + b]);
      }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!ok[t]) continue;
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (am[t][j] == code[t]) acc[j] += g[t][j];
    }
    store_vec<VEC>(G.dx + ((n * G.H + h) * G.W + w) * G.C + c, acc);
    return;
  }
  for (int ho = ho_lo; ho <= ho_hi; ++ho)
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      const uint8_t code = (uint8_t)((hp - ho * G.s) * G.k + (wp - wo * G.s));
      const long long o = ((n * G.Ho + ho) * G.Wo + wo) * G.C + c;
      uint8_t am[VEC];
      load_codes<VEC>(G.argmax + o, am);
      bool any = false;
#pragma unroll
      for (int j = 0; j < VEC; ++j) any |= am[j] == code;
      if (!any) continue;
      float g[VEC];
      load_vec<VEC>(G.dout + o, g);
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (am[j] == code) acc[j] += g[j];
    }
  store_vec<VEC>(G.dx + ((n * G.H + h) * G.W + w) * G.C + c, acc);
}

// ---- bias + ReLU (decoder): elementwise over [rows, C] -------------------------------------------------
template <int VEC, bool RELU>
__global__ void __launch_bounds__(kPoolBlock) bias_relu_fwd_kernel(const float* x, const float* bias, float* out,
                                                                   long long n_vec, int Cv) {
  const long long i = (long long)blockIdx.x * kPoolBlock + threadIdx.x;
  if (i >= n_vec) return;
  float v[VEC], b[VEC];
  load_vec<VEC>(x + i * VEC, v);
  load_vec<VEC>(bias + (i % Cv) * VEC, b);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float vb = v[j] + b[j];
    v[j] = (!RELU || vb > 0.f || vb != vb) ? vb : 0.f;
  }
  store_vec<VEC>(out + i * VEC, v);
}

template <int VEC>
__global__ void __launch_bounds__(kPoolBlock) relu_bwd_kernel(const float* dout, const float* out, float* dx,
                                                              long long n_vec) {
  const long long i = (long long)blockIdx.x * kPoolBlock + threadIdx.x;
  if (i >= n_vec) return;
  float g[VEC], o[VEC];
  load_vec<VEC>(dout + i * VEC, g);
  load_vec<VEC>(out + i * VEC, o);
#pragma unroll
  for (int j = 0; j < VEC; ++j)
    if (o[j] <= 0.f) g[j] = 0.f;  // threshold_backward: NaN outputs pass the gradient
  store_vec<VEC>(dx + i * VEC, g);
}

// ---- d bias: masked column sums of dout [rows, C] --------------------------------------------------------
// MASK 0: argmax code != kDead (pooled outputs), 1: out > 0 (bias + ReLU outputs; `out <= 0` drops, NaN passes),
// 2: every entry (plain bias).
// Grid (channel tiles, slabs): a CTA of TX x TY threads owns TX channel vectors and the rows of one slab; thread
// (cx, ry) sums rows ry, ry + TY, ...; the TY partials fold in shared memory in ry order; stage 2 adds the slabs in
// slab order.
struct ColSumDev {
  const float* dout;
  const uint8_t* argmax;
  const float* out;
  float* partial;  // [slabs][C]
  long long rows, rows_per_slab;
  int C, tx;
};

template <int VEC, int MASK>
__global__ void __launch_bounds__(kPoolBlock) colsum_partial_kernel(const ColSumDev G) {
  __shared__ float s_acc[kPoolBlock * 4];
  const int Cv = G.C / VEC;
  const int tx = G.tx, ty = kPoolBlock / tx;
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int cv = blockIdx.x * tx + cx;
  const long long r0 = (long long)blockIdx.y * G.rows_per_slab;
  const long long r1 = r0 + G.rows_per_slab < G.rows ? r0 + G.rows_per_slab : G.rows;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (cv < Cv) {
    for (long long r = r0 + ry; r < r1; r += ty) {
      const long long o = r * G.C + (long long)cv * VEC;
      float g[VEC];
      load_vec<VEC>(G.dout + o, g);
      if constexpr (MASK == 0) {
        uint8_t am[VEC];
        load_codes<VEC>(G.argmax + o, am);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (am[j] != kDead) acc[j] += g[j];
      } else if constexpr (MASK == 1) {
        float ov[VEC];
        load_vec<VEC>(G.out + o, ov);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (!(ov[j] <= 0.f)) acc[j] += g[j];
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += g[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) s_acc[threadIdx.x * 4 + j] = acc[j];
  __syncthreads();
  if (ry == 0 && cv < Cv) {
    for (int q = 1; q < ty; ++q)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += s_acc[(q * tx + cx) * 4 + j];
    store_vec<VEC>(G.partial + (long long)blockIdx.y * G.C + (long long)cv * VEC, acc);
  }
}

// stage 2: a CTA owns 32 channels; thread (cx, sy) adds slabs sy, sy + 8, ... (independent loads), the 8 partials fold in
// sy order -- a fixed order, whatever the slab count
__global__ void __launch_bounds__(kPoolBlock) colsum_final_kernel(const float* partial, float* dbias, int slabs, int C) {
  __shared__ float s_part[kPoolBlock];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (c < C)
    for (int s = sy; s < slabs; s += kPoolBlock / 32) acc += partial[(long long)s * C + c];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (sy == 0 && c < C) {
    for (int q = 1; q < kPoolBlock / 32; ++q) acc += s_part[q * 32 + cx];
    dbias[c] = acc;
  }
}

// lanes over channel vectors (the smallest power of two >= min(Cv, 32)) and the number of row slabs
static void colsum_plan(long long rows, int C, int vec, int* tx_out, int* slabs_out) {
  const int Cv = C / vec;
  int tx = 1;
  while (tx < Cv && tx < 32) tx *= 2;
  const int tiles = (Cv + tx - 1) / tx;
  const int ty = kPoolBlock / tx;
  long long want = (long long)device_info().sm_count * 2 / tiles;  // ~2 CTAs per SM (stage 2 walks the slabs)
  const long long cap = (rows + ty * 8 - 1) / (ty * 8);           // >= 8 rows per thread
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 4096) want = 4096;
  *tx_out = tx, *slabs_out = (int)want;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int launch_colsum(const float* dout, const uint8_t* argmax, const float* out, float* dbias, void* ws, size_t ws_bytes,
                         long long rows, int C, cudaStream_t stream, const char* who) {
  B200RL_REQUIRE(ws != nullptr, "%s: workspace is null", who);
  float* partial = static_cast<float*>(ws);
  partial += ((16 - (reinterpret_cast<uintptr_t>(partial) & 15)) & 15) / sizeof(float);
  const size_t usable = ws_bytes - (size_t)(reinterpret_cast<uint8_t*>(partial) - static_cast<uint8_t*>(ws));
  const bool v4 = C % 4 == 0 && aligned16(dout) && (!out || aligned16(out)) &&
                  (!argmax || (reinterpret_cast<uintptr_t>(argmax) & 3u) == 0);
  const int vec = v4 ? 4 : 1;
  int tx, slabs;
  colsum_plan(rows, C, vec, &tx, &slabs);
  B200RL_REQUIRE(usable >= (size_t)slabs * C * sizeof(float), "%s: workspace too small (%zu < %zu bytes)", who, usable,
                 (size_t)slabs * C * sizeof(float));
  const int Cv = C / vec;
  ColSumDev G{dout, argmax, out, partial, rows, (rows + slabs - 1) / slabs, C, tx};
  const dim3 grid((unsigned)((Cv + tx - 1) / tx), (unsigned)slabs);
  if (argmax) {
    if (v4) colsum_partial_kernel<4, 0><<<grid, kPoolBlock, 0, stream>>>(G);
    else colsum_partial_kernel<1, 0><<<grid, kPoolBlock, 0, stream>>>(G);
  } else if (out) {
    if (v4) colsum_partial_kernel<4, 1><<<grid, kPoolBlock, 0, stream>>>(G);
    else colsum_partial_kernel<1, 1><<<grid, kPoolBlock, 0, stream>>>(G);
  } else {
    if (v4) colsum_partial_kernel<4, 2><<<grid, kPoolBlock, 0, stream>>>(G);
    else colsum_partial_kernel<1, 2><<<grid, kPoolBlock, 0, stream>>>(G);
  }
  int rc = check_launch(who);
  if (rc) return rc;
  colsum_final_kernel<<<(C + 31) / 32, kPoolBlock, 0, stream>>>(partial, dbias, slabs, C);
  return check_launch(who);
}

static int make_pool(PoolDev* G, int64_t N, int64_t H, int64_t W, int64_t C, int k, int s, int p, const char* who) {
  B200RL_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 1, "%s: bad shape", who);
  B200RL_REQUIRE(k >= 1 && s >= 1 && p >= 0 && 2 * p <= k, "%s: bad window (kernel %d, stride %d, padding %d)", who, k, s, p);
  B200RL_UNSUPPORTED(k > 15, "%s: kernel %d exceeds 15", who, k);
  B200RL_UNSUPPORTED(H > (1 << 20) || W > (1 << 20) || C > (1 << 20), "%s: extent too large", who);
  B200RL_REQUIRE(H + 2 * p >= k && W + 2 * p >= k, "%s: window larger than the padded input", who);
  G->N = N, G->H = (int)H, G->W = (int)W, G->C = (int)C, G->k = k, G->s = s, G->p = p;
  G->Ho = (int)((H + 2 * p - k) / s + 1), G->Wo = (int)((W + 2 * p - k) / s + 1);
  return B200RL_OK;
}

}  // namespace b200rl

extern "C" size_t b200rl_nhwc_bias_grad_workspace_bytes(int64_t rows, int64_t C) {
  if (rows < 1 || C < 1) return 32;
  int tx, s4 = 0, s1 = 0;
  if (C % 4 == 0) b200rl::colsum_plan(rows, (int)C, 4, &tx, &s4);
  b200rl::colsum_plan(rows, (int)C, 1, &tx, &s1);  // the scalar form (unaligned pointers) may cut more slabs
  return (size_t)(s4 > s1 ? s4 : s1) * (size_t)C * sizeof(float) + 32;
}

extern "C" int b200rl_nhwc_bias_pool_relu_fwd(const float* x, const float* bias, float* out, uint8_t* argmax, int64_t N,
                                              int64_t H, int64_t W, int64_t C, int kernel, int stride, int padding, int relu,
                                              b200rl_stream_t stream) {
  using namespace b200rl;
  PoolDev G{};
  int rc = make_pool(&G, N, H, W, C, kernel, stride, padding, "nhwc_bias_pool_relu_fwd");
  if (rc) return rc;
  if (N == 0) return B200RL_OK;
  B200RL_REQUIRE(x && out, "nhwc_bias_pool_relu_fwd: null pointer");
  G.x = x, G.bias = bias, G.out = out, G.argmax = argmax, G.relu = relu;
  const bool v4 = C % 4 == 0 && aligned16(x) && aligned16(out) && (!bias || aligned16(bias)) &&
                  (!argmax || (reinterpret_cast<uintptr_t>(argmax) & 3u) == 0);
  const long long total = N * G.Ho * G.Wo * (C / (v4 ? 4 : 1));
  B200RL_UNSUPPORTED((total + kPoolBlock - 1) / kPoolBlock > 0x7fffffffLL, "nhwc_bias_pool_relu_fwd: tensor too large");
  const unsigned grid = (unsigned)((total + kPoolBlock - 1) / kPoolBlock);
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = N * H * W * C < 0x7fffffffLL;  // every element index fits 31 bits
  if (v4) {
    if (kernel == 3 && small) pool_fwd_kernel<4, 3, unsigned><<<grid, kPoolBlock, 0, st>>>(G);
    else if (kernel == 3) pool_fwd_kernel<4, 3, long long><<<grid, kPoolBlock, 0, st>>>(G);
    else pool_fwd_kernel<4, 0, long long><<<grid, kPoolBlock, 0, st>>>(G);
  } else {
    if (kernel == 3) pool_fwd_kernel<1, 3, long long><<<grid, kPoolBlock, 0, st>>>(G);
    else pool_fwd_kernel<1, 0, long long><<<grid, kPoolBlock, 0, st>>>(G);
  }
  return check_launch("nhwc_bias_pool_relu_fwd");
}

extern "C" int b200rl_nhwc_bias_pool_relu_bwd(const float* dout, const uint8_t* argmax, float* dx, float* dbias, void* workspace,
                                              size_t workspace_bytes, int64_t N, int64_t H, int64_t W, int64_t C, int kernel,
                                              int stride, int padding, b200rl_stream_t stream) {
  using namespace b200rl;
  PoolDev G{};
  int rc = make_pool(&G, N, H, W, C, kernel, stride, padding, "nhwc_bias_pool_relu_bwd");
  if (rc) return rc;
  if (N == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
    return check_launch("nhwc_bias_pool_relu_bwd");
  }
  B200RL_REQUIRE(dout && argmax && dx, "nhwc_bias_pool_relu_bwd: null pointer");
  G.dout = dout, G.argmax = const_cast<uint8_t*>(argmax), G.dx = dx;
  const bool v4 = C % 4 == 0 && aligned16(dout) && aligned16(dx) && (reinterpret_cast<uintptr_t>(argmax) & 3u) == 0;
  const long long total = N * H * W * (C / (v4 ? 4 : 1));
  B200RL_UNSUPPORTED((total + kPoolBlock - 1) / kPoolBlock > 0x7fffffffLL, "nhwc_bias_pool_relu_bwd: tensor too large");
  const unsigned grid = (unsigned)((total + kPoolBlock - 1) / kPoolBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (kernel == 3 && stride == 2 && padding == 1) {  // a thread per 2 x 2 input patch
    const long long patches = N * ((H + 1) / 2) * ((W + 1) / 2) * (C / (v4 ? 4 : 1));
    const unsigned pgrid = (unsigned)((patches + kPoolBlock - 1) / kPoolBlock);
    const bool small = N * H * W * C < 0x7fffffffLL;
    if (v4 && small) pool_bwd_k3s2_kernel<4, unsigned><<<pgrid, kPoolBlock, 0, st>>>(G);
    else if (v4) pool_bwd_k3s2_kernel<4, long long><<<pgrid, kPoolBlock, 0, st>>>(G);
    else pool_bwd_k3s2_kernel<1, long long><<<pgrid, kPoolBlock, 0, st>>>(G);
  } else if (v4) {
    pool_bwd_kernel<4><<<grid, kPoolBlock, 0, st>>>(G);
  } else {
    pool_bwd_kernel<1><<<grid, kPoolBlock, 0, st>>>(G);
  }
  rc = check_launch("nhwc_bias_pool_relu_bwd");
  if (rc || !dbias) return rc;
  return launch_colsum(dout, argmax, nullptr, dbias, workspace, workspace_bytes, N * G.Ho * G.Wo, (int)C, st,
                       "nhwc_bias_pool_relu_bwd (bias gradient)");
}

extern "C" int b200rl_nhwc_bias_relu_fwd(const float* x, const float* bias, float* out, int64_t rows, int64_t C, int relu,
                                         b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rows >= 0 && C >= 1 && C <= (1 << 20), "nhwc_bias_relu_fwd: bad shape");
  if (rows == 0) return B200RL_OK;
  B200RL_REQUIRE(x && bias && out, "nhwc_bias_relu_fwd: null pointer");
  const bool v4 = C % 4 == 0 && aligned16(x) && aligned16(out) && aligned16(bias);
  const int vec = v4 ? 4 : 1;
  const long long n_vec = rows * C / vec;
  B200RL_UNSUPPORTED((n_vec + kPoolBlock - 1) / kPoolBlock > 0x7fffffffLL, "nhwc_bias_relu_fwd: tensor too large");
  const unsigned grid = (unsigned)((n_vec + kPoolBlock - 1) / kPoolBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (v4 && relu) bias_relu_fwd_kernel<4, true><<<grid, kPoolBlock, 0, st>>>(x, bias, out, n_vec, (int)(C / 4));
  else if (v4) bias_relu_fwd_kernel<4, false><<<grid, kPoolBlock, 0, st>>>(x, bias, out, n_vec, (int)(C / 4));
  else if (relu) bias_relu_fwd_kernel<1, true><<<grid, kPoolBlock, 0, st>>>(x, bias, out, n_vec, (int)C);
  else bias_relu_fwd_kernel<1, false><<<grid, kPoolBlock, 0, st>>>(x, bias, out, n_vec, (int)C);
  return check_launch("nhwc_bias_relu_fwd");
}

extern "C" int b200rl_nhwc_bias_relu_bwd(const float* dout, const float* out, float* dx, float* dbias, void* workspace,
                                         size_t workspace_bytes, int64_t rows, int64_t C, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rows >= 0 && C >= 1 && C <= (1 << 20), "nhwc_bias_relu_bwd: bad shape");
  B200RL_REQUIRE(rows == 0 || (dout && (out == nullptr || dx)), "nhwc_bias_relu_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (dbias) cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), st);
    return check_launch("nhwc_bias_relu_bwd");
  }
  // the bias gradient reads dout: before dx, which may alias it, is written
  if (dbias) {
    int rc = launch_colsum(dout, nullptr, out, dbias, workspace, workspace_bytes, rows, (int)C, st, "nhwc_bias_relu_bwd (bias gradient)");
    if (rc) return rc;
  }
  if (out == nullptr) return B200RL_OK;  // plain bias: dx IS dout, nothing to mask
  const bool v4 = C % 4 == 0 && aligned16(dout) && aligned16(out) && aligned16(dx);
  const int vec = v4 ? 4 : 1;
  const long long n_vec = rows * C / vec;
  B200RL_UNSUPPORTED((n_vec + kPoolBlock - 1) / kPoolBlock > 0x7fffffffLL, "nhwc_bias_relu_bwd: tensor too large");
  const unsigned grid = (unsigned)((n_vec + kPoolBlock - 1) / kPoolBlock);
  if (v4) relu_bwd_kernel<4><<<grid, kPoolBlock, 0, st>>>(dout, out, dx, n_vec);
  else relu_bwd_kernel<1><<<grid, kPoolBlock, 0, st>>>(dout, out, dx, n_vec);
  return check_launch("nhwc_bias_relu_bwd");
}

// K4c: GridNet per-cell MultiDiscrete heads (+ pick_position) -- forward, backward and the
// fully fused PPO loss, one launch each.
//
// Replaces shared/actor/gridnet.py:38-193 over shared/actor/categorical.py:12-54 and, in the
// fused mode, ppo/ppo.py:326-361 with the autograd backward of the whole chain: ~60-80
// eager launches forward and twice that backward in the reference.
//
// Shape of the kernel (memory-bound; 8*HW*S' + HW*S + ... bytes per sample):
//   * one thread-block cluster per sample; each CTA of the cluster owns HW / cluster cells
//     and stages its [cells, S'] logits tile and [cells, S] mask tile in shared memory with
//     1-D bulk async copies (TMA engine, completion on an mbarrier) -- logits are read from
//     HBM exactly once;
//   * one thread per cell walks the heads of its cell: masked max, one exp per valid entry
//     (sum e and sum e*(x-max) give logsumexp and entropy together), the chosen action's
//     log-prob with the value-dependent gate; (lse, entropy) per head stay in registers;
//   * per-sample sums go through warp shuffles, shared memory and, when the sample spans
//     several CTAs, distributed shared memory across the cluster; the pick_position
//     categorical over all cells of the sample uses the same path for its max / partition sum;
//   * thread 0 turns the sample's log-prob into ratio / clipped surrogate / KL terms
//     (ppo_terms.cuh); one warp handles the value heads;
//   * the backward overwrites the logits tile in place with d loss / d logits and one bulk
//     async copy writes the tile back -- dlogits are written exactly once.
#include <cooperative_groups.h>

#include "categorical.cuh"
#include "philox.cuh"
#include "ppo_terms.cuh"

namespace cg = cooperative_groups;

namespace b200rl {

constexpr int kGridBlock = 256;
constexpr int kRegHeads = 8;   // heads whose (lse, entropy) live in registers
constexpr int kMaxPick = 4;
constexpr int kMaxCpt = 4;     // cells per thread

enum GridMode { kFwd = 0, kBwd = 1, kPpo = 2, kSample = 3 };

struct GridDev {
  const void* logits;
  void* dlogits;
  const uint8_t* mask;
  const uint8_t* pick_mask;
  const void* actions;
  const void* pick_actions;
  long long B, HW;
  int A, S, Sp, n_pick;
  int act_dtype, pick_dtype;
  int nvec[B200RL_MAX_HEADS], off[B200RL_MAX_HEADS], gate_ref[B200RL_MAX_HEADS], gate_val[B200RL_MAX_HEADS];
  float* logp;
  float* entropy;
  const float* dlogp_in;
  const float* dent_in;
  int cluster;        // CTAs per sample
  int cells_per_cta;  // HW / cluster
  // sampling
  uint64_t seed, offset;
  void* actions_out;
  void* pick_actions_out;
};

__device__ __forceinline__ int load_index(const void* base, int dtype, long long i) {
  switch (dtype) {
    case B200RL_U8: return (int)static_cast<const uint8_t*>(base)[i];
    case B200RL_I32: return (int)static_cast<const int32_t*>(base)[i];
    default: return (int)static_cast<const long long*>(base)[i];
  }
}
__device__ __forceinline__ void store_index(void* base, int dtype, long long i, int v) {
  switch (dtype) {
    case B200RL_U8: static_cast<uint8_t*>(base)[i] = (uint8_t)v; break;
    case B200RL_I32: static_cast<int32_t*>(base)[i] = v; break;
    default: static_cast<long long*>(base)[i] = v; break;
  }
}

// ---- tile staging ----------------------------------------------------------------------------
// Global bytes [src, src+bytes) land at smem_base + (src & 15) so that the 16-byte aligned
// middle can go through one bulk async copy; the <16-byte head and tail are plain byte copies.
struct TilePlan {
  uint32_t lead;    // src & 15
  uint32_t head;    // bytes before the aligned middle
  uint32_t middle;  // multiple of 16
  uint32_t tail;
};
__device__ __forceinline__ TilePlan plan_tile(const void* src, uint32_t bytes) {
  TilePlan t;
  t.lead = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
  t.head = (16u - t.lead) & 15u;
  if (t.head > bytes) t.head = bytes;
  t.middle = (bytes - t.head) & ~15u;
  t.tail = bytes - t.head - t.middle;
  return t;
}
__device__ __forceinline__ void copy_edges_in(uint8_t* dst, const uint8_t* src, const TilePlan& t) {
  const int tid = threadIdx.x;
  if (tid < (int)t.head) dst[tid] = src[tid];
  if (tid >= 32 && tid - 32 < (int)t.tail) dst[t.head + t.middle + tid - 32] = src[t.head + t.middle + tid - 32];
}

// ---- per-sample reductions over the CTA and the cluster --------------------------------------
// v[0..K) summed (or maxed) over every thread of every CTA of the cluster; result in all threads.
template <int K, bool IS_MAX>
__device__ __forceinline__ void sample_reduce(float (&v)[K], float* s_warp /*[K*32]*/, float* s_cta /*[K]*/,
                                              int cluster_size) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kGridBlock / 32;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = IS_MAX ? warp_max(v[k]) : warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_warp[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float x = lane < nwarps ? s_warp[k * 32 + lane] : (IS_MAX ? -INFINITY : 0.f);
      x = IS_MAX ? warp_max(x) : warp_sum(x);
      if (lane == 0) s_cta[k] = x;
    }
  }
  if (cluster_size > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();  // every CTA's s_cta is written
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float acc = IS_MAX ? -INFINITY : 0.f;
      for (int r = 0; r < cluster_size; ++r) {
        const float x = cluster.map_shared_rank(s_cta, r)[k];
        acc = IS_MAX ? fmaxf(acc, x) : acc + x;
      }
      v[k] = acc;
    }
    cluster.sync();  // nobody still reads s_cta when it is reused
  } else {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = s_cta[k];
    __syncthreads();
  }
}

// ---- the kernel --------------------------------------------------------------------------------
template <int MODE, typename LT, int CPT>
__global__ void __launch_bounds__(kGridBlock) gridnet_kernel(const GridDev G, const PpoDev P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ float s_warp[8 * 32];
  __shared__ float s_cta[8];
  __shared__ float s_bcast[4];
  __shared__ float s_pick[kMaxPick * 3];  // lse, entropy, any per pick head

  const int tid = threadIdx.x;
  const int cluster_size = G.cluster;
  const int rank = cluster_size > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const long long b = blockIdx.x / cluster_size;
  const int cells = G.cells_per_cta;
  const long long cell0 = (long long)rank * cells;  // first cell of this CTA within the sample

  const uint32_t tile_bytes = (uint32_t)cells * G.Sp * sizeof(LT);
  const uint32_t mask_bytes = (uint32_t)cells * G.S;
  const uint8_t* g_tile = static_cast<const uint8_t*>(G.logits) + ((b * G.HW + cell0) * G.Sp) * sizeof(LT);
  const uint8_t* g_mask = G.mask + (b * G.HW + cell0) * G.S;
  const TilePlan tp = plan_tile(g_tile, tile_bytes);
  const TilePlan mp = plan_tile(g_mask, mask_bytes);
  uint8_t* s_tile_base = smem;  // capacity tile_bytes + 16, 16-byte aligned
  uint8_t* s_mask_base = smem + ((tile_bytes + 16 + 15) & ~15u);
  LT* tile = reinterpret_cast<LT*>(s_tile_base + tp.lead);
  const uint8_t* mtile = s_mask_base + mp.lead;

  // ---- 1. stage logits + masks -------------------------------------------------------------
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, tp.middle + mp.middle);
    if (tp.middle) bulk_g2s(s_tile_base + tp.lead + tp.head, g_tile + tp.head, tp.middle, &s_bar);
    if (mp.middle) bulk_g2s(s_mask_base + mp.lead + mp.head, g_mask + mp.head, mp.middle, &s_bar);
  }
  copy_edges_in(s_tile_base + tp.lead, g_tile, tp);
  copy_edges_in(s_mask_base + mp.lead, g_mask, mp);

  // per-cell actions (registers) while the copies fly
  uint32_t act_lo[CPT], act_hi[CPT];  // 8 packed bytes per cell
  if (MODE != kSample) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + j * kGridBlock;
      uint32_t lo = 0, hi = 0;
      if (c < cells) {
        const long long base = (b * G.HW + cell0 + c) * G.A;
        for (int h = 0; h < G.A; ++h) {
          const uint32_t a = (uint32_t)load_index(G.actions, G.act_dtype, base + h) & 0xffu;
          if (h < 4) lo |= a << (8 * h); else hi |= a << (8 * (h - 4));
        }
      }
      act_lo[j] = lo, act_hi[j] = hi;
    }
  }
  auto action_of = [&](int j, int h) -> int {
    return (int)(((h < 4 ? act_lo[j] >> (8 * h) : act_hi[j] >> (8 * (h - 4)))) & 0xffu);
  };

  __syncthreads();        // edge bytes written by other threads
  mbar_wait(&s_bar, 0);   // bulk bytes landed

  // ---- 2. forward over this thread's cells ---------------------------------------------------
  float lse[CPT][kRegHeads], ent[CPT][kRegHeads];
  uint32_t any_bits[CPT], gate_bits[CPT];
  float logp_acc = 0.f, ent_acc = 0.f;
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int c = tid + j * kGridBlock;
    any_bits[j] = 0, gate_bits[j] = 0;
    if (c >= cells) continue;
    const LT* x = tile + (long long)c * G.Sp;
    const uint8_t* m = mtile + (long long)c * G.S;
#pragma unroll
    for (int h = 0; h < kRegHeads; ++h) {
      lse[j][h] = 0.f, ent[j][h] = 0.f;
      if (h >= G.A) continue;
      const int off = G.off[h], n = G.nvec[h];
      float mx = -INFINITY;
      bool any = false;
      for (int k = 0; k < n; ++k)
        if (m[off + k]) {
          any = true;
          mx = fmaxf(mx, to_f32(x[off + k]));
        }
      if (!any) continue;
      float s = 0.f, q = 0.f;
      for (int k = 0; k < n; ++k)
        if (m[off + k]) {
          const float d = to_f32(x[off + k]) - mx;
          const float e = expf(d);
          s += e;
          q = fmaf(e, d, q);
        }
      const float ls = logf(s);
      lse[j][h] = mx + ls;
      ent[j][h] = ls - q / s;  // -sum p * logp
      any_bits[j] |= 1u << h;
      ent_acc += ent[j][h];
      if (MODE != kSample) {
        const int gr = G.gate_ref[h];
        const bool gated_in = gr < 0 || action_of(j, gr) == G.gate_val[h];
        if (gated_in) {
          gate_bits[j] |= 1u << h;
          const int a = action_of(j, h);
          const float xa = (a < n && m[off + a]) ? to_f32(x[off + a]) : kF32Lowest;
          logp_acc += xa - lse[j][h];
        }
      }
    }
  }

  // ---- 3. pick_position categoricals over all cells of the sample ------------------------------
  for (int kp = 0; kp < G.n_pick; ++kp) {
    const uint8_t* pm = G.pick_mask + (b * G.n_pick + kp) * G.HW + cell0;
    float mx[1] = {-INFINITY};
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = tid + j * kGridBlock;
      if (c < cells && pm[c]) mx[0] = fmaxf(mx[0], to_f32(tile[(long long)c * G.Sp + G.S + kp]));
    }
    sample_reduce<1, true>(mx, s_warp, s_cta, cluster_size);
    const bool any = mx[0] > -INFINITY;
    float sq[2] = {0.f, 0.f};
    if (any) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = tid + j * kGridBlock;
        if (c < cells && pm[c]) {
          const float d = to_f32(tile[(long long)c * G.Sp + G.S + kp]) - mx[0];
          const float e = expf(d);
          sq[0] += e;
          sq[1] = fmaf(e, d, sq[1]);
        }
      }
    }
    sample_reduce<2, false>(sq, s_warp, s_cta, cluster_size);
    float p_lse = 0.f, p_ent = 0.f;
    if (any) {
      const float ls = logf(sq[0]);
      p_lse = mx[0] + ls;
      p_ent = ls - sq[1] / sq[0];
    }
    if (tid == 0) s_pick[kp * 3] = p_lse, s_pick[kp * 3 + 1] = p_ent, s_pick[kp * 3 + 2] = any ? 1.f : 0.f;
    if (MODE != kSample && any) {
      const long long a = load_index(G.pick_actions, G.pick_dtype, b * G.n_pick + kp);
      const long long local = a - cell0;
      if (local >= 0 && local < cells && (int)(local % kGridBlock) == tid) {
        const float xa = pm[local] ? to_f32(tile[local * G.Sp + G.S + kp]) : kF32Lowest;
        logp_acc += xa - p_lse;
      }
    }
    if (rank == 0 && tid == 0) ent_acc += p_ent;
  }

  // ---- 4. per-sample totals -----------------------------------------------------------------
  float tot[2] = {logp_acc, ent_acc};
  float dlogp = 0.f, dent = 0.f;
  if (MODE == kFwd || MODE == kPpo) sample_reduce<2, false>(tot, s_warp, s_cta, cluster_size);

  if (MODE == kFwd) {
    if (rank == 0 && tid == 0) G.logp[b] = tot[0], G.entropy[b] = tot[1];
    if (cluster_size > 1) cg::this_cluster().sync();
    return;
  }
  if (MODE == kBwd) {
    dlogp = G.dlogp_in[b], dent = G.dent_in[b];
    __syncthreads();  // s_pick visible
  }
  if (MODE == kPpo) {
    // ---- 5. PPO scalar stage ----------------------------------------------------------------
    if (tid == 0) {
      PolicyTerms t = ppo_policy_terms(P, b, tot[0]);
      s_bcast[0] = t.dlogp;
      if (rank == 0) {
        double* row = P.partials + b * ppo_nstat(P.V);
        row[0] = t.surrogate, row[1] = tot[1], row[2] = t.kl, row[3] = t.clipped;
        if (G.logp) G.logp[b] = tot[0];
        if (G.entropy) G.entropy[b] = tot[1];
      }
    }
    if (rank == 0 && tid >= 32 && tid < 32 + P.V) {
      const int v = tid - 32;
      float2 r = ppo_value_terms(P, b, v);
      double* row = P.partials + b * ppo_nstat(P.V);
      row[kPolicyStats + v] = r.x, row[kPolicyStats + P.V + v] = r.y;
    }
    __syncthreads();
    dlogp = s_bcast[0];
    dent = ppo_dentropy(P, 1);
  }

  // ---- 6. backward in place -------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const int c = tid + j * kGridBlock;
    if (c >= cells) continue;
    LT* x = tile + (long long)c * G.Sp;
    const uint8_t* m = mtile + (long long)c * G.S;
#pragma unroll
    for (int h = 0; h < kRegHeads; ++h) {
      if (h >= G.A) continue;
      const int off = G.off[h], n = G.nvec[h];
      if (!((any_bits[j] >> h) & 1u)) {
        for (int k = 0; k < n; ++k) x[off + k] = from_f32<LT>(0.f);
        continue;
      }
      const float dl = ((gate_bits[j] >> h) & 1u) ? dlogp : 0.f;
      const int a = action_of(j, h);
      const float l = lse[j][h], e = ent[j][h];
      for (int k = 0; k < n; ++k) {
        float g = 0.f;
        if (m[off + k]) {
          const float lp = to_f32(x[off + k]) - l;
          const float p = expf(lp);
          g = dl * ((k == a ? 1.f : 0.f) - p) - dent * p * (lp + e);
        }
        x[off + k] = from_f32<LT>(g);
      }
    }
    for (int kp = 0; kp < G.n_pick; ++kp) {
      const float p_lse = s_pick[kp * 3], p_ent = s_pick[kp * 3 + 1];
      const bool any = s_pick[kp * 3 + 2] != 0.f;
      float g = 0.f;
      if (any && G.pick_mask[(b * G.n_pick + kp) * G.HW + cell0 + c]) {
        const long long a = load_index(G.pick_actions, G.pick_dtype, b * G.n_pick + kp);
        const float lp = to_f32(x[G.S + kp]) - p_lse;
        const float p = expf(lp);
        g = dlogp * ((a == cell0 + c ? 1.f : 0.f) - p) - dent * p * (lp + p_ent);
      }
      x[G.S + kp] = from_f32<LT>(g);
    }
  }

  // ---- 7. write the gradient tile back ----------------------------------------------------------
  uint8_t* g_out = static_cast<uint8_t*>(G.dlogits) + ((b * G.HW + cell0) * G.Sp) * sizeof(LT);
  const bool same_phase = (reinterpret_cast<uintptr_t>(g_out) & 15u) == tp.lead;
  fence_async_smem();
  __syncthreads();
  const uint8_t* s_src = s_tile_base + tp.lead;
  if (same_phase) {
    if (tid == 0 && tp.middle) {
      bulk_s2g(g_out + tp.head, s_src + tp.head, tp.middle);
      bulk_commit();
    }
    if (tid < (int)tp.head) g_out[tid] = s_src[tid];
    if (tid >= 32 && tid - 32 < (int)tp.tail)
      g_out[tp.head + tp.middle + tid - 32] = s_src[tp.head + tp.middle + tid - 32];
    if (tid == 0 && tp.middle) bulk_wait_read<0>();  // shared memory must outlive the copy's reads
  } else {
    for (uint32_t o = tid; o < tile_bytes; o += kGridBlock) g_out[o] = s_src[o];
  }
  if (cluster_size > 1) cg::this_cluster().sync();  // no CTA exits while a peer may read its smem
}

// ---- host side -----------------------------------------------------------------------------------
struct GridLaunch {
  int cluster;
  int cells_per_cta;
  int cpt;
  size_t smem;
};

static size_t grid_smem(long long cells, int Sp, int S, size_t lt) {
  const size_t tile = (size_t)cells * Sp * lt;
  return ((tile + 16 + 15) & ~(size_t)15) + (size_t)cells * S + 32;
}

static int plan_launch(const GridDev& G, size_t lt, GridLaunch* out) {
  const size_t limit_two = 110 * 1024, limit_one = (size_t)device_info().max_smem_optin - 2048;
  GridLaunch best{0, 0, 0, 0};
  for (int pass = 0; pass < 2 && !best.cluster; ++pass) {
    for (int cs = 1; cs <= 8; cs *= 2) {
      if (G.HW % cs) continue;
      const long long cells = G.HW / cs;
      if (cells > (long long)kGridBlock * kMaxCpt) continue;
      const size_t smem = grid_smem(cells, G.Sp, G.S, lt);
      if (smem <= (pass == 0 ? limit_two : limit_one)) {
        int cpt = (int)((cells + kGridBlock - 1) / kGridBlock);
        cpt = cpt <= 1 ? 1 : (cpt <= 2 ? 2 : 4);
        best = GridLaunch{cs, (int)cells, cpt, smem};
        break;
      }
    }
  }
  if (!best.cluster) {
    set_error("gridnet: a sample of HW=%lld cells x S'=%d logits does not fit 8 CTAs of shared memory", G.HW, G.Sp);
    return B200RL_EUNSUPPORTED;
  }
  *out = best;
  return B200RL_OK;
}

template <int MODE, typename LT, int CPT>
static int launch_one(GridDev& G, const PpoDev& P, const GridLaunch& L, cudaStream_t stream) {
  auto kernel = gridnet_kernel<MODE, LT, CPT>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
  if (e != cudaSuccess) {
    set_error("gridnet: cudaFuncSetAttribute(%zu bytes): %s", L.smem, cudaGetErrorString(e));
    return B200RL_ECUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(G.B * L.cluster));
  cfg.blockDim = dim3(kGridBlock);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)L.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, (const GridDev)G, P);
  if (e != cudaSuccess) {
    set_error("gridnet launch (cluster %d, %zu B smem): %s", L.cluster, L.smem, cudaGetErrorString(e));
    return B200RL_ECUDA;
  }
  return B200RL_OK;
}

template <int MODE, typename LT>
static int launch_cpt(GridDev& G, const PpoDev& P, const GridLaunch& L, cudaStream_t stream) {
  switch (L.cpt) {
    case 1: return launch_one<MODE, LT, 1>(G, P, L, stream);
    case 2: return launch_one<MODE, LT, 2>(G, P, L, stream);
    default: return launch_one<MODE, LT, 4>(G, P, L, stream);
  }
}

template <int MODE>
static int launch_mode(GridDev& G, const PpoDev& P, int logits_dtype, cudaStream_t stream) {
  GridLaunch L;
  int rc = plan_launch(G, logits_dtype == B200RL_BF16 ? 2 : 4, &L);
  if (rc) return rc;
  G.cluster = L.cluster, G.cells_per_cta = L.cells_per_cta;
  if (logits_dtype == B200RL_BF16) return launch_cpt<MODE, __nv_bfloat16>(G, P, L, stream);
  return launch_cpt<MODE, float>(G, P, L, stream);
}

static int make_grid(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask, const uint8_t* pick_mask,
                     const void* actions, const void* pick_actions, bool need_actions, GridDev* out, const char* who) {
  B200RL_REQUIRE(d && logits && mask, "%s: null pointer", who);
  B200RL_REQUIRE(d->B >= 0 && d->HW >= 1 && d->A >= 1 && d->n_pick >= 0, "%s: bad shape", who);
  B200RL_REQUIRE(d->nvec_host != nullptr, "%s: nvec is null", who);
  B200RL_UNSUPPORTED(d->A > kRegHeads, "%s: A=%d action planes (this build keeps at most %d in registers)", who, d->A,
                     kRegHeads);
  B200RL_UNSUPPORTED(d->n_pick > kMaxPick, "%s: n_pick=%d exceeds %d", who, d->n_pick, kMaxPick);
  B200RL_UNSUPPORTED(d->logits_dtype != B200RL_F32 && d->logits_dtype != B200RL_BF16, "%s: logits dtype %d", who,
                     d->logits_dtype);
  B200RL_REQUIRE(d->n_pick == 0 || pick_mask, "%s: pick_mask is null", who);
  if (need_actions) {
    B200RL_REQUIRE(actions != nullptr, "%s: actions is null", who);
    B200RL_REQUIRE(d->n_pick == 0 || pick_actions, "%s: pick_actions is null", who);
    B200RL_UNSUPPORTED(d->act_dtype != B200RL_U8 && d->act_dtype != B200RL_I32 && d->act_dtype != B200RL_I64,
                       "%s: action dtype %d", who, d->act_dtype);
    B200RL_UNSUPPORTED(d->n_pick > 0 && d->pick_dtype != B200RL_I32 && d->pick_dtype != B200RL_I64,
                       "%s: pick action dtype %d", who, d->pick_dtype);
  }
  GridDev G{};
  G.logits = logits, G.mask = mask, G.pick_mask = pick_mask, G.actions = actions, G.pick_actions = pick_actions;
  G.B = d->B, G.HW = d->HW, G.A = d->A, G.n_pick = d->n_pick;
  G.act_dtype = d->act_dtype, G.pick_dtype = d->pick_dtype;
  int S = 0;
  for (int h = 0; h < d->A; ++h) {
    B200RL_REQUIRE(d->nvec_host[h] >= 1, "%s: nvec[%d]=%d", who, h, d->nvec_host[h]);
    B200RL_UNSUPPORTED(d->nvec_host[h] > 256, "%s: nvec[%d]=%d exceeds 256", who, h, d->nvec_host[h]);
    G.nvec[h] = d->nvec_host[h], G.off[h] = S;
    S += d->nvec_host[h];
    const int gr = d->gate_ref_host ? d->gate_ref_host[h] : -1;
    B200RL_REQUIRE(gr < d->A, "%s: gate_ref[%d]=%d out of range", who, h, gr);
    G.gate_ref[h] = gr;
    G.gate_val[h] = (gr >= 0 && d->gate_val_host) ? d->gate_val_host[h] : 0;
  }
  G.S = S, G.Sp = S + d->n_pick;
  *out = G;
  return B200RL_OK;
}

}  // namespace b200rl

extern "C" int b200rl_gridnet_fwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                  const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                  float* logp, float* entropy, b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, true, &G, "gridnet_fwd");
  if (rc) return rc;
  B200RL_REQUIRE(logp && entropy, "gridnet_fwd: null output");
  if (G.B == 0) return B200RL_OK;
  G.logp = logp, G.entropy = entropy;
  PpoDev P{};
  return launch_mode<kFwd>(G, P, d->logits_dtype, (cudaStream_t)stream);
}

extern "C" int b200rl_gridnet_bwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                  const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                  const float* dlogp, const float* dentropy, void* dlogits,
                                  b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, true, &G, "gridnet_bwd");
  if (rc) return rc;
  B200RL_REQUIRE(dlogp && dentropy && dlogits, "gridnet_bwd: null pointer");
  if (G.B == 0) return B200RL_OK;
  G.dlogp_in = dlogp, G.dent_in = dentropy, G.dlogits = dlogits;
  PpoDev P{};
  return launch_mode<kBwd>(G, P, d->logits_dtype, (cudaStream_t)stream);
}

extern "C" int b200rl_ppo_gridnet_loss(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                       const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                       const b200rl_ppo_args* args, void* dlogits, float* logp_out,
                                       float* entropy_out, void* workspace, size_t workspace_bytes,
                                       b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, true, &G, "ppo_gridnet_loss");
  if (rc) return rc;
  B200RL_REQUIRE(dlogits != nullptr, "ppo_gridnet_loss: dlogits is null");
  PpoDev P;
  rc = ppo_make_dev(args, G.B, workspace, workspace_bytes, &P);
  if (rc) return rc;
  G.dlogits = dlogits, G.logp = logp_out, G.entropy = entropy_out;
  cudaStream_t s = (cudaStream_t)stream;
  rc = launch_mode<kPpo>(G, P, d->logits_dtype, s);
  if (rc) return rc;
  return ppo_launch_finalize(P, G.B, 1, s);
}

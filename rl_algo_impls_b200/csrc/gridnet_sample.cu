// K5: rollout-time GridNet sampling -- one action per head per cell (+ pick_position) and the
// joint log-prob of the draw, in one launch.
//
// Replaces GridnetDistribution.sample (shared/actor/gridnet.py:195-207: one torch.multinomial
// per head) followed by log_prob (gridnet.py:104-176) at shared/policy/actor_critic.py:311-314.
// Gumbel-max: argmax_k (x_k + g_k) over the valid entries is an exact draw from the masked
// softmax; g_k comes from Philox4x32-10 keyed by `seed` with counter
// (sample*HW + cell, offset, head, k/4), so the stream is reproducible and independent of the
// launch geometry.  A row with no valid entry draws uniformly over all entries and contributes
// log-prob 0, as the reference (probs = 1/n, normalised logits = 0).
// One CTA per sample: rollouts have few samples (n_envs) and the pick categorical needs a
// sample-wide argmax.  Like the fused loss, the kernel is mask driven: it zero-fills the sample's
// action row, scans the mask bytes with 128-bit loads into a list of non-empty cells and only
// those cells draw random numbers (a head with no valid entry returns action 0, log-prob 0).
#include "categorical.cuh"
#include "mask_scan.cuh"
#include "philox.cuh"

namespace b200rl {

constexpr int kSampleBlock = 256;
constexpr int kStashUnits = 1024;  // non-empty cells whose per-head log-probs are parked in shared memory

struct SampleDev {
  const void* logits;
  int logits_dtype;
  const uint8_t* mask;
  const uint8_t* pick_mask;
  long long B, HW;
  int A, S, Sp, n_pick;
  int ld;  // elements between consecutive cells' logit rows (>= Sp)
  int nvec[B200RL_MAX_HEADS], off[B200RL_MAX_HEADS], gate_ref[B200RL_MAX_HEADS], gate_val[B200RL_MAX_HEADS];
  uint64_t seed, offset;
  const long long* offset_dev;
  void* actions_out;
  int act_dtype;
  void* pick_out;
  int pick_dtype;
  float* logp;
  // wide heads are cut into parts of kPartEntries entries that adjacent lanes of one warp draw in parallel:
  // part slot p of a cell covers entries [part_lo[p], part_lo[p] + part_len[p]) of head part_head[p];
  // n_parts == 0: the plan does not fit a warp (one thread per head instead)
  int n_parts;
  int8_t part_head[32];
  int16_t part_lo[32], part_len[32];
  int8_t part_count[32];  // parts of this slot's head (valid on the head's first slot), 0 on the others
  int max_part_count;
};
constexpr int kPartEntries = 16;  // a multiple of the 4 entries one Philox block serves

__device__ __forceinline__ float logit_at(const SampleDev& G, long long i) {
  return G.logits_dtype == B200RL_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(G.logits)[i])
                                       : static_cast<const float*>(G.logits)[i];
}
__device__ __forceinline__ void put_index(void* base, int dtype, long long i, long long v) {
  switch (dtype) {
    case B200RL_U8: static_cast<uint8_t*>(base)[i] = (uint8_t)v; break;
    case B200RL_I32: static_cast<int32_t*>(base)[i] = (int32_t)v; break;
    default: static_cast<long long*>(base)[i] = v; break;
  }
}
// -log(-log u): sampling noise only (never enters a reported log-prob), so the fast logarithms do
__device__ __forceinline__ float gumbel(uint32_t bits) { return -__logf(-__logf(u01(bits))); }
__device__ __forceinline__ uint64_t stream_id(uint64_t offset, int head, int kblock) {
  return (offset << 24) ^ ((uint64_t)head << 16) ^ (uint64_t)kblock;
}

__global__ void __launch_bounds__(kSampleBlock) gridnet_sample_kernel(const SampleDev G) {
  __shared__ float s_red[32];
  __shared__ float s_best[32];
  __shared__ long long s_arg[32];
  __shared__ float s_scalar[2];
  const long long b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t offset = G.offset + (G.offset_dev ? (uint64_t)*G.offset_dev : 0ull);
  float logp_acc = 0.f;

  // ---- non-empty cells of this sample -----------------------------------------------------------------
  extern __shared__ uint16_t s_list[];  // [HW]
  __shared__ uint32_t s_bitmap[kChunkCells / 32];
  __shared__ int s_n, s_chunk_n;
  const long long act_bytes = G.HW * G.A * (G.act_dtype == B200RL_U8 ? 1 : (G.act_dtype == B200RL_I32 ? 4 : 8));
  zero_fill<kSampleBlock>(static_cast<uint8_t*>(G.actions_out) + b * act_bytes, (uint32_t)act_bytes);
  if (tid == 0) s_n = 0;
  for (long long c0 = 0; c0 < G.HW; c0 += kChunkCells) {
    const int cells = (int)min((long long)kChunkCells, G.HW - c0);
    if (tid < kChunkCells / 32) s_bitmap[tid] = 0u;
    __syncthreads();
    scan_mask<kSampleBlock>(G.mask + (b * G.HW + c0) * G.S, (uint32_t)cells * (uint32_t)G.S, (uint32_t)G.S, s_bitmap,
                            RowPrefetch{nullptr, 0u, 0u});
    __syncthreads();
    compact_cells(s_bitmap, (cells + 31) >> 5, s_list + s_n, (int)c0, &s_chunk_n);
    __syncthreads();
    if (tid == 0) s_n += s_chunk_n;
    __syncthreads();
  }
  const int n_unit = s_n;

  // ---- one thread per (non-empty cell, head): Gumbel-max draw + log-prob in a single pass -----------------
  // The pass keeps an online softmax (running max m, sum s of e^(x-m)) next to the running arg-max of
  // x + gumbel, so the chosen entry's log-prob is x_best - (m + log s) without a second sweep.
  __shared__ float s_lp[kStashUnits * B200RL_MAX_HEADS > 8192 ? 8192 : kStashUnits * B200RL_MAX_HEADS];
  const int stash_units = 8192 / (G.A > 0 ? G.A : 1) < kStashUnits ? 8192 / G.A : kStashUnits;
  // running statistics of a draw over a range of entries: Gumbel arg-max and online softmax
  struct Draw {
    float best_score, x_best, mx, sum;
    int best;
  };
  auto draw_range = [&](long long cell, int h, int k_lo, int k_hi) -> Draw {  // k_lo a multiple of 4
    const int off = G.off[h];
    const long long xbase = cell * G.ld + off;
    const uint8_t* m = G.mask + cell * G.S + off;
    Draw d{-INFINITY, 0.f, -INFINITY, 0.f, 0};
    for (int k0 = k_lo; k0 < k_hi; k0 += 4) {
      uint32_t valid = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k0 + j < k_hi && m[k0 + j]) valid |= 1u << j;
      if (!valid) continue;  // no random numbers spent on masked entries
      const Philox4 r = philox4x32_10(G.seed, (uint64_t)cell, stream_id(offset, h, k0 >> 2));
      const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!((valid >> j) & 1u)) continue;
        const float x = logit_at(G, xbase + k0 + j);
        const float score = x + gumbel(bits[j]);
        if (score > d.best_score) d.best_score = score, d.best = k0 + j, d.x_best = x;
        const float nm = fmaxf(d.mx, x);
        d.sum = d.sum * __expf(d.mx - nm) + __expf(x - nm);  // exp(-inf) == 0 on the first valid entry
        d.mx = nm;
      }
    }
    return d;
  };
  // `b` covers the entries after `a`'s: the earlier entry wins a tie, as one sequential scan would decide
  auto merge = [](Draw a, const Draw& b) -> Draw {
    if (b.best_score > a.best_score) a.best_score = b.best_score, a.best = b.best, a.x_best = b.x_best;
    const float nm = fmaxf(a.mx, b.mx);
    if (nm > -INFINITY) a.sum = a.sum * __expf(a.mx - nm) + b.sum * __expf(b.mx - nm);
    a.mx = nm;
    return a;
  };
  auto finish = [](const Draw& d, float* lp_out) -> int {
    *lp_out = d.sum > 0.f ? d.x_best - (d.mx + logf(d.sum)) : 0.f;  // a head with no valid entry: action 0, log-prob 0
    return d.best;
  };
  auto sample_head = [&](long long cell, int h, float* lp_out) -> int {
    return finish(draw_range(cell, h, 0, G.nvec[h]), lp_out);
  };

  const int n_stashed = n_unit < stash_units ? n_unit : stash_units;
  const int n_items = n_stashed * G.A;
  if (G.n_parts > 0) {
    // a warp takes 32 / n_parts cells at a time; lane = (cell slot, part slot); the first lane of a head folds the
    // partial draws of its other parts in, in entry order, through shuffles (every lane takes part in them)
    const int per_warp = 32 / G.n_parts;
    const int slot = lane / G.n_parts, p = lane - slot * G.n_parts;
    const bool lane_used = slot < per_warp;
    const int h = lane_used ? G.part_head[p] : 0;
    for (int u0 = warp * per_warp; u0 < n_stashed; u0 += (kSampleBlock / 32) * per_warp) {  // warp-uniform trips
      const int u = u0 + slot;
      const bool live = lane_used && u < n_stashed;
      const long long cell = live ? b * G.HW + s_list[u] : 0;
      Draw d{-INFINITY, 0.f, -INFINITY, 0.f, 0};
      if (live) d = draw_range(cell, h, G.part_lo[p], G.part_lo[p] + G.part_len[p]);
      const int my_parts = live ? G.part_count[p] : 0;
      for (int q = 1; q < G.max_part_count; ++q) {
        Draw o;
        o.best_score = __shfl_down_sync(0xffffffffu, d.best_score, q);
        o.x_best = __shfl_down_sync(0xffffffffu, d.x_best, q);
        o.mx = __shfl_down_sync(0xffffffffu, d.mx, q);
        o.sum = __shfl_down_sync(0xffffffffu, d.sum, q);
        o.best = __shfl_down_sync(0xffffffffu, d.best, q);
        if (q < my_parts) d = merge(d, o);  // only a head's first lane accumulates; it never reads its own merges back
      }
      if (my_parts > 0) {
        float lp;
        const int a = finish(d, &lp);
        put_index(G.actions_out, G.act_dtype, cell * G.A + h, a);
        s_lp[u * G.A + h] = lp;
      }
    }
  } else {
    for (int item = tid; item < n_items; item += kSampleBlock) {
      const int u = item / G.A, h = item - u * G.A;
      const long long cell = b * G.HW + s_list[u];
      float lp;
      const int a = sample_head(cell, h, &lp);
      put_index(G.actions_out, G.act_dtype, cell * G.A + h, a);
      s_lp[item] = lp;
    }
  }
  __syncthreads();  // actions of the reference heads are written
  for (int item = tid; item < n_items; item += kSampleBlock) {
    const int u = item / G.A, h = item - u * G.A;
    const int gr = G.gate_ref[h];
    if (gr >= 0) {
      const long long cell = b * G.HW + s_list[u];
      long long a_ref;
      switch (G.act_dtype) {
        case B200RL_U8: a_ref = static_cast<const uint8_t*>(G.actions_out)[cell * G.A + gr]; break;
        case B200RL_I32: a_ref = static_cast<const int32_t*>(G.actions_out)[cell * G.A + gr]; break;
        default: a_ref = static_cast<const long long*>(G.actions_out)[cell * G.A + gr]; break;
      }
      if (a_ref != G.gate_val[h]) continue;
    }
    logp_acc += s_lp[item];
  }
  // cells beyond the shared-memory stash (dense masks on big maps): one thread walks all heads of a cell
  for (int u = stash_units + tid; u < n_unit; u += kSampleBlock) {
    const long long cell = b * G.HW + s_list[u];
    int chosen[B200RL_MAX_HEADS];
    float lp[B200RL_MAX_HEADS];
    for (int h = 0; h < G.A; ++h) {
      chosen[h] = sample_head(cell, h, &lp[h]);
      put_index(G.actions_out, G.act_dtype, cell * G.A + h, chosen[h]);
    }
    for (int h = 0; h < G.A; ++h) {
      const int gr = G.gate_ref[h];
      if (gr < 0 || chosen[gr] == G.gate_val[h]) logp_acc += lp[h];
    }
  }

  // pick_position: Gumbel arg-max over the cells of the sample, then its log-prob
  for (int kp = 0; kp < G.n_pick; ++kp) {
    const uint8_t* pm = G.pick_mask + (b * G.n_pick + kp) * G.HW;
    // does any cell qualify?
    float anyf = 0.f;
    for (long long c = tid; c < G.HW; c += kSampleBlock) anyf = fmaxf(anyf, pm[c] ? 1.f : 0.f);
    anyf = warp_max(anyf);
    if (lane == 0) s_red[warp] = anyf;
    __syncthreads();
    if (warp == 0) {
      float x = lane < kSampleBlock / 32 ? s_red[lane] : 0.f;
      x = warp_max(x);
      if (lane == 0) s_scalar[0] = x;
    }
    __syncthreads();
    const bool any = s_scalar[0] > 0.f;
    float best_score = -INFINITY, mx = -INFINITY;
    long long best = 0x7fffffffffffLL;
    for (long long c = tid; c < G.HW; c += kSampleBlock) {
      if (any && !pm[c]) continue;
      const Philox4 r = philox4x32_10(G.seed, (uint64_t)(b * G.HW + c), stream_id(offset, G.A + kp, 0));
      const float x = any ? logit_at(G, (b * G.HW + c) * G.ld + G.S + kp) : 0.f;
      const float score = x + gumbel(r.x);
      if (score > best_score) best_score = score, best = c;
      mx = fmaxf(mx, x);
    }
    // block arg-max (ties -> lowest cell index, so the result does not depend on warp order)
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best_score, o);
      const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      if (os > best_score || (os == best_score && ob < best)) best_score = os, best = ob;
    }
    mx = warp_max(mx);
    __syncthreads();
    if (lane == 0) s_best[warp] = best_score, s_arg[warp] = best, s_red[warp] = mx;
    __syncthreads();
    if (warp == 0) {
      float bs = lane < kSampleBlock / 32 ? s_best[lane] : -INFINITY;
      long long ba = lane < kSampleBlock / 32 ? s_arg[lane] : 0x7fffffffffffLL;
      float m2 = lane < kSampleBlock / 32 ? s_red[lane] : -INFINITY;
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long ob = __shfl_xor_sync(0xffffffffu, ba, o);
        if (os > bs || (os == bs && ob < ba)) bs = os, ba = ob;
      }
      m2 = warp_max(m2);
      if (lane == 0) s_arg[0] = ba, s_scalar[1] = m2;
    }
    __syncthreads();
    const long long pick = s_arg[0];
    mx = s_scalar[1];
    float se = 0.f;
    if (any)
      for (long long c = tid; c < G.HW; c += kSampleBlock)
        if (pm[c]) se += expf(logit_at(G, (b * G.HW + c) * G.ld + G.S + kp) - mx);
    se = warp_sum(se);
    __syncthreads();
    if (lane == 0) s_red[warp] = se;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < kSampleBlock / 32; ++w) tot += s_red[w];
      put_index(G.pick_out, G.pick_dtype, b * G.n_pick + kp, pick);
      if (any) logp_acc += logit_at(G, (b * G.HW + pick) * G.ld + G.S + kp) - (mx + logf(tot));
    }
    __syncthreads();
  }

  logp_acc = warp_sum(logp_acc);
  __syncthreads();
  if (lane == 0) s_red[warp] = logp_acc;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < kSampleBlock / 32; ++w) tot += s_red[w];
    G.logp[b] = tot;
  }
}

}  // namespace b200rl

extern "C" int b200rl_gridnet_sample(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                     const uint8_t* pick_mask, uint64_t seed, uint64_t offset,
                                     const int64_t* offset_dev, void* actions_out, void* pick_actions_out,
                                     float* logp, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(d && logits && mask && actions_out && logp, "gridnet_sample: null pointer");
  B200RL_REQUIRE(d->B >= 0 && d->HW >= 1 && d->A >= 1 && d->A <= B200RL_MAX_HEADS && d->n_pick >= 0,
                 "gridnet_sample: bad shape");
  B200RL_REQUIRE(d->n_pick == 0 || (pick_mask && pick_actions_out), "gridnet_sample: pick tensors are null");
  B200RL_UNSUPPORTED(d->logits_dtype != B200RL_F32 && d->logits_dtype != B200RL_BF16, "gridnet_sample: logits dtype");
  if (d->B == 0) return B200RL_OK;
  SampleDev G{};
  G.logits = logits, G.logits_dtype = d->logits_dtype, G.mask = mask, G.pick_mask = pick_mask;
  G.B = d->B, G.HW = d->HW, G.A = d->A, G.n_pick = d->n_pick;
  int S = 0;
  for (int h = 0; h < d->A; ++h) {
    G.nvec[h] = d->nvec_host[h], G.off[h] = S;
    S += d->nvec_host[h];
    const int gr = d->gate_ref_host ? d->gate_ref_host[h] : -1;
    B200RL_REQUIRE(gr < d->A, "gridnet_sample: gate_ref[%d] out of range", h);
    G.gate_ref[h] = gr;
    G.gate_val[h] = (gr >= 0 && d->gate_val_host) ? d->gate_val_host[h] : 0;
  }
  G.S = S, G.Sp = S + d->n_pick;
  B200RL_REQUIRE(d->logits_ld == 0 || d->logits_ld >= G.Sp, "gridnet_sample: logits_ld=%lld is narrower than a row (%d)",
                 (long long)d->logits_ld, G.Sp);
  G.ld = d->logits_ld ? (int)d->logits_ld : G.Sp;
  {  // part plan: heads in order, each cut into parts of kPartEntries entries on adjacent lanes
    int np = 0, max_count = 1;
    bool fits = true;
    for (int h = 0; h < d->A && fits; ++h) {
      const int count = (G.nvec[h] + kPartEntries - 1) / kPartEntries;
      if (np + count > 32 || count > 127) {
        fits = false;
        break;
      }
      for (int q = 0; q < count; ++q, ++np) {
        G.part_head[np] = (int8_t)h;
        G.part_lo[np] = (int16_t)(q * kPartEntries);
        const int len = G.nvec[h] - q * kPartEntries;
        G.part_len[np] = (int16_t)(len < kPartEntries ? len : kPartEntries);
        G.part_count[np] = (int8_t)(q == 0 ? count : 0);
      }
      if (count > max_count) max_count = count;
    }
    G.n_parts = fits ? np : 0;
    G.max_part_count = max_count;
  }
  G.seed = seed, G.offset = offset, G.offset_dev = reinterpret_cast<const long long*>(offset_dev);
  G.actions_out = actions_out, G.act_dtype = d->act_dtype, G.pick_out = pick_actions_out, G.pick_dtype = d->pick_dtype;
  G.logp = logp;
  B200RL_UNSUPPORTED(d->HW > 16384, "gridnet_sample: HW=%lld cells exceeds 16384", (long long)d->HW);
  gridnet_sample_kernel<<<(unsigned)d->B, kSampleBlock, (size_t)d->HW * sizeof(uint16_t), (cudaStream_t)stream>>>(G);
  return check_launch("gridnet_sample");
}

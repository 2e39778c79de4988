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
  long long* actions_wide;  // nullable: the same actions once more as int64 (what a host env is handed)
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
  int8_t head_first_part[B200RL_MAX_HEADS];  // part slot of each head's first part
  int max_part_count;
};
constexpr int kPartEntries = 16;  // a multiple of the 4 entries one Philox block serves

__device__ __forceinline__ float logit_at(const SampleDev& G, long long i) {
  return G.logits_dtype == B200RL_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(G.logits)[i])
                                       : static_cast<const float*>(G.logits)[i];
}
__device__ __forceinline__ void put_index(void* base, int dtype, long long i, long long v);
// a per-cell action: into the compact rollout-buffer dtype and, when asked for, the int64 copy for the host
__device__ __forceinline__ void put_action(const struct SampleDev& G, long long i, long long v);
__device__ __forceinline__ void put_index(void* base, int dtype, long long i, long long v) {
  switch (dtype) {
    case B200RL_U8: static_cast<uint8_t*>(base)[i] = (uint8_t)v; break;
    case B200RL_I32: static_cast<int32_t*>(base)[i] = (int32_t)v; break;
    default: static_cast<long long*>(base)[i] = v; break;
  }
}
__device__ __forceinline__ void put_action(const SampleDev& G, long long i, long long v) {
  put_index(G.actions_out, G.act_dtype, i, v);
  if (G.actions_wide) G.actions_wide[i] = v;
}
// -log(-log u): sampling noise only (never enters a reported log-prob), so the fast logarithms do
__device__ __forceinline__ float gumbel(uint32_t bits) { return -__logf(-__logf(u01(bits))); }
__device__ __forceinline__ uint64_t stream_id(uint64_t offset, int head, int kblock) {
  return (offset << 24) ^ ((uint64_t)head << 16) ^ (uint64_t)kblock;
}

__global__ void __launch_bounds__(kSampleBlock) gridnet_sample_kernel(const SampleDev G) {
  __shared__ float s_red[32];
  __shared__ float s_best[32], s_xbest[32], s_mx[32], s_sum[32];
  __shared__ long long s_arg[32];
  const long long b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t offset = G.offset + (G.offset_dev ? (uint64_t)*G.offset_dev : 0ull);
  float logp_acc = 0.f;

  // ---- non-empty cells of this sample -----------------------------------------------------------------
  // ONE scan of the whole sample's mask bytes (independent 128-bit loads, flags into a shared bitmap of HW bits),
  // one compaction by warp 0: two barriers whatever the map size (was four per chunk of 256 cells)
  extern __shared__ __align__(16) uint8_t s_dyn[];
  uint32_t* s_bitmap = reinterpret_cast<uint32_t*>(s_dyn);                                 // [ceil(HW / 32)]
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_dyn + (((G.HW + 31) >> 5) << 2));       // [HW]
  __shared__ int s_n;
  const long long act_bytes = G.HW * G.A * (G.act_dtype == B200RL_U8 ? 1 : (G.act_dtype == B200RL_I32 ? 4 : 8));
  const int words = (int)((G.HW + 31) >> 5);
  for (int w = tid; w < words; w += kSampleBlock) s_bitmap[w] = 0u;
  zero_fill<kSampleBlock>(static_cast<uint8_t*>(G.actions_out) + b * act_bytes, (uint32_t)act_bytes);
  if (G.actions_wide)
    zero_fill<kSampleBlock>(reinterpret_cast<uint8_t*>(G.actions_wide + b * G.HW * G.A), (uint32_t)(G.HW * G.A * 8));
  __syncthreads();
  scan_mask<kSampleBlock>(G.mask + b * G.HW * G.S, (uint32_t)G.HW * (uint32_t)G.S, (uint32_t)G.S, s_bitmap,
                          RowPrefetch{nullptr, 0u, 0u});
  __syncthreads();
  compact_cells(s_bitmap, words, s_list, 0, &s_n);
  __syncthreads();
  const int n_unit = s_n;

  // ---- one thread per (non-empty cell, head or part of a head): Gumbel-max draw + log-prob in a single pass ----
  // The pass keeps an online softmax (running max m, sum s of e^(x-m)) next to the running arg-max of
  // x + gumbel, so the chosen entry's log-prob is x_best - (m + log s) without a second sweep.
  __shared__ float s_lp[kStashUnits * B200RL_MAX_HEADS > 8192 ? 8192 : kStashUnits * B200RL_MAX_HEADS];
  const int stash_units = 8192 / (G.A > 0 ? G.A : 1) < kStashUnits ? 8192 / G.A : kStashUnits;
  // running statistics of a draw over a range of entries: Gumbel arg-max and online softmax
  struct Draw {
    float best_score, x_best, mx, sum;
    int best;
  };
  // Entries [k_lo, k_hi) of head h of `cell`, kPartEntries at a time: the mask bytes and the logits of a stretch are
  // loaded first, all independent (a load per valid entry behind its mask test serialised 16 round trips), then
  // consumed four entries per Philox block.  k_lo is a multiple of 4.
  auto draw_range = [&](long long cell, int h, int k_lo, int k_hi) -> Draw {
    const int off = G.off[h];
    const long long xbase = cell * G.ld + off;
    const uint8_t* m = G.mask + cell * G.S + off;
    Draw d{-INFINITY, 0.f, -INFINITY, 0.f, 0};
#pragma unroll 1
    for (int c0 = k_lo; c0 < k_hi; c0 += kPartEntries) {
      const int n = k_hi - c0 < kPartEntries ? k_hi - c0 : kPartEntries;
      float x[kPartEntries];
      uint32_t valid = 0u;
      if (G.logits_dtype == B200RL_BF16) {
        const __nv_bfloat16* row = static_cast<const __nv_bfloat16*>(G.logits) + xbase + c0;
#pragma unroll
        for (int j = 0; j < kPartEntries; ++j) x[j] = j < n ? __bfloat162float(row[j]) : 0.f;
      } else {
        const float* row = static_cast<const float*>(G.logits) + xbase + c0;
#pragma unroll
        for (int j = 0; j < kPartEntries; ++j) x[j] = j < n ? row[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kPartEntries; ++j)
        if (j < n && m[c0 + j]) valid |= 1u << j;
#pragma unroll
      for (int q = 0; q < kPartEntries / 4; ++q) {
        if (!((valid >> (4 * q)) & 15u)) continue;  // no random numbers spent on masked entries
        const Philox4 r = philox4x32_10(G.seed, (uint64_t)cell, stream_id(offset, h, (c0 >> 2) + q));
        const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!((valid >> (4 * q + j)) & 1u)) continue;
          const float xv = x[4 * q + j];
          const float score = xv + gumbel(bits[j]);
          if (score > d.best_score) d.best_score = score, d.best = c0 + 4 * q + j, d.x_best = xv;
          const float nm = fmaxf(d.mx, xv);
          d.sum = d.sum * __expf(d.mx - nm) + __expf(xv - nm);  // exp(-inf) == 0 on the first valid entry
          d.mx = nm;
        }
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

  if (G.n_parts > 0) {
    // a warp takes 32 / n_parts cells at a time; lane = (cell slot, part slot); the first lane of a head folds the
    // partial draws of its other parts in, in entry order, through shuffles (every lane takes part in them).  All the
    // heads of a cell sit in one warp, so a gated head reads its reference head's action through a shuffle too: no
    // log-prob stash, no second pass.
    const int per_warp = 32 / G.n_parts;
    const int slot = lane / G.n_parts, p = lane - slot * G.n_parts;
    const bool lane_used = slot < per_warp;
    const int h = lane_used ? G.part_head[p] : 0;
    const int gr = lane_used ? G.gate_ref[h] : -1;
    const int gate_val = G.gate_val[h];
    const int ref_lane = gr >= 0 ? slot * G.n_parts + G.head_first_part[gr] : lane;
    for (int u0 = warp * per_warp; u0 < n_unit; u0 += (kSampleBlock / 32) * per_warp) {  // warp-uniform trips
      const int u = u0 + slot;
      const bool live = lane_used && u < n_unit;
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
      int a = 0;
      float lp = 0.f;
      if (my_parts > 0) {
        a = finish(d, &lp);
        put_action(G, cell * G.A + h, a);
      }
      const int a_ref = __shfl_sync(0xffffffffu, a, ref_lane);
      if (my_parts > 0 && (gr < 0 || a_ref == gate_val)) logp_acc += lp;
    }
  } else {
    const int n_stashed = n_unit < stash_units ? n_unit : stash_units;
    const int n_items = n_stashed * G.A;
    for (int item = tid; item < n_items; item += kSampleBlock) {
      const int u = item / G.A, h = item - u * G.A;
      const long long cell = b * G.HW + s_list[u];
      float lp;
      const int a = sample_head(cell, h, &lp);
      put_action(G, cell * G.A + h, a);
      s_lp[item] = lp;
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
        put_action(G, cell * G.A + h, chosen[h]);
      }
      for (int h = 0; h < G.A; ++h) {
        const int gr = G.gate_ref[h];
        if (gr < 0 || chosen[gr] == G.gate_val[h]) logp_acc += lp[h];
      }
    }
  }

  // pick_position: Gumbel arg-max over the valid cells of the sample and the online softmax of their logits in ONE pass
  // (was: an any-valid pass, the arg-max pass, a sum-exp pass, each with its own block reductions).  No valid cell:
  // a uniform draw over all cells, log-prob 0 (the reference's all-masked row) -- a second pass, rare.
  for (int kp = 0; kp < G.n_pick; ++kp) {
    const uint8_t* pm = G.pick_mask + (b * G.n_pick + kp) * G.HW;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0: valid cells; pass 1 (only if there is none): every cell, x = 0
      float best_score = -INFINITY, x_best = 0.f, mx = -INFINITY, sum = 0.f;
      long long best = 0x7fffffffffffLL;
      for (long long c = tid; c < G.HW; c += kSampleBlock) {
        if (pass == 0 && !pm[c]) continue;
        const Philox4 r = philox4x32_10(G.seed, (uint64_t)(b * G.HW + c), stream_id(offset, G.A + kp, 0));
        const float x = pass == 0 ? logit_at(G, (b * G.HW + c) * G.ld + G.S + kp) : 0.f;
        const float score = x + gumbel(r.x);
        if (score > best_score) best_score = score, best = c, x_best = x;
        const float nm = fmaxf(mx, x);
        sum = sum * __expf(mx - nm) + __expf(x - nm);
        mx = nm;
      }
      // block arg-max (ties -> lowest cell index, so the result does not depend on warp order) + softmax merge
      auto fold = [&](float os, long long ob, float ox, float om, float osum) {
        if (os > best_score || (os == best_score && ob < best)) best_score = os, best = ob, x_best = ox;
        const float nm = fmaxf(mx, om);
        if (nm > -INFINITY) sum = sum * __expf(mx - nm) + osum * __expf(om - nm);
        mx = nm;
      };
      for (int o = 16; o > 0; o >>= 1)
        fold(__shfl_xor_sync(0xffffffffu, best_score, o), __shfl_xor_sync(0xffffffffu, best, o),
             __shfl_xor_sync(0xffffffffu, x_best, o), __shfl_xor_sync(0xffffffffu, mx, o),
             __shfl_xor_sync(0xffffffffu, sum, o));
      __syncthreads();  // the previous round's readers of the partial arrays are done
      if (lane == 0) s_best[warp] = best_score, s_arg[warp] = best, s_xbest[warp] = x_best, s_mx[warp] = mx, s_sum[warp] = sum;
      __syncthreads();
      constexpr int NW = kSampleBlock / 32;
      const int wl = lane & (NW - 1);  // every group of NW lanes of every warp folds the same partials: no broadcast needed
      best_score = s_best[wl], best = s_arg[wl], x_best = s_xbest[wl], mx = s_mx[wl], sum = s_sum[wl];
      for (int o = NW / 2; o > 0; o >>= 1)
        fold(__shfl_xor_sync(0xffffffffu, best_score, o), __shfl_xor_sync(0xffffffffu, best, o),
             __shfl_xor_sync(0xffffffffu, x_best, o), __shfl_xor_sync(0xffffffffu, mx, o),
             __shfl_xor_sync(0xffffffffu, sum, o));
      const bool any = sum > 0.f;  // block-uniform
      if (pass == 0 && !any) continue;
      if (tid == 0) {
        put_index(G.pick_out, G.pick_dtype, b * G.n_pick + kp, best);
        if (pass == 0) logp_acc += x_best - (mx + logf(sum));
      }
      break;
    }
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
                                     float* logp, int64_t* actions_wide_out, b200rl_stream_t stream) {
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
      G.head_first_part[h] = (int8_t)np;
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
  G.actions_wide = reinterpret_cast<long long*>(actions_wide_out);
  G.actions_out = actions_out, G.act_dtype = d->act_dtype, G.pick_out = pick_actions_out, G.pick_dtype = d->pick_dtype;
  G.logp = logp;
  B200RL_UNSUPPORTED(d->HW > 16384, "gridnet_sample: HW=%lld cells exceeds 16384", (long long)d->HW);
  const size_t smem = (size_t)((d->HW + 31) / 32) * sizeof(uint32_t) + (size_t)d->HW * sizeof(uint16_t);
  gridnet_sample_kernel<<<(unsigned)d->B, kSampleBlock, smem, (cudaStream_t)stream>>>(G);
  return check_launch("gridnet_sample");
}

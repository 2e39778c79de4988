// K4c: GridNet per-cell MultiDiscrete heads (+ pick_position) -- forward, backward and the
// fully fused PPO loss, one launch each.
//
// Replaces shared/actor/gridnet.py:38-193 over shared/actor/categorical.py:12-54 and, in the
// fused mode, ppo/ppo.py:326-361 with the autograd backward of the whole chain: ~60-80
// eager launches forward and twice that backward in the reference.
//
// The kernel is driven by the action MASK.  A cell whose S mask bytes are all zero (no unit
// on it: ~94-98 % of the cells of a MicroRTS / Lux map) contributes exactly 0 to the
// log-prob, to the entropy and to the gradient (reference semantics: a fully masked row
// normalises to uniform, log_prob 0, entropy -0, gradient blocked by torch.where), so its
// logits are never read; its dlogits row is plain zero fill.
//
//   1. every thread zero-fills its share of the sample's dlogits (128-bit stores) and scans
//      the mask bytes with 128-bit loads; a non-zero word flags its cell(s) in a shared bitmap;
//   2. warp 0 compacts the bitmap into a list of "unit" cells;
//   3. unit cells are processed by groups of G lanes (G = 8 or 16 for the reference's action
//      planes): every lane owns one piece of <= PMAX adjacent logits of one head -- small
//      heads are one lane, a wide head (MicroRTS' 49-way attack target) is spread over an
//      aligned power-of-two block of lanes and reduced with xor shuffles.  Adjacent lanes read
//      adjacent addresses, so a cell's row is one coalesced ~300-byte access;
//   4. per-sample sums run in float64 through shuffles, shared memory and -- when a sample
//      spans several CTAs -- distributed shared memory across the thread-block cluster; the
//      pick_position categorical over all cells of the sample uses the same path;
//   5. thread 0 turns the sample's log-prob into ratio / clipped surrogate / KL terms
//      (ppo_terms.cuh), one warp handles the value heads;
//   6. the same lanes revisit the unit cells (L1/L2 hits) and overwrite their zero-filled
//      dlogits rows with d loss / d logits.
//
// HBM traffic per sample: dlogits written once, masks read once, logits / actions read only
// for unit cells.  No shared-memory tile => ~20 KB of shared memory per CTA and full occupancy.
#include <cooperative_groups.h>

#include "ppo_terms.cuh"

namespace cg = cooperative_groups;

namespace b200rl {

constexpr int kGridBlock = 128;
constexpr int kMaxPick = 4;
constexpr int kMaxCellsPerCta = 4096;

enum GridMode { kFwd = 0, kBwd = 1, kPpo = 2 };

// One lane slot of a G-lane group: a piece of `len` adjacent logits of head `head`.
struct LaneSlot {
  uint16_t off;       // first logit of the piece within the cell's row
  uint16_t head_off;  // first logit of the piece's head
  uint8_t len;        // 0: idle lane
  uint8_t head;
  uint8_t width;  // lanes of this head's block (power of two, block is width-aligned)
  uint8_t first;  // first lane of its block
};

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
  int gate_ref[B200RL_MAX_HEADS], gate_val[B200RL_MAX_HEADS];
  float* logp;
  float* entropy;
  const float* dlogp_in;
  const float* dent_in;
  int cluster;        // CTAs per sample
  int cells_per_cta;  // HW / cluster
  int G;              // lanes per cell group (power of two <= 32)
  int max_width;      // widest head block
  LaneSlot slot[32];
};

__device__ __forceinline__ int load_index(const void* base, int dtype, long long i) {
  switch (dtype) {
    case B200RL_U8: return (int)__ldg(static_cast<const uint8_t*>(base) + i);
    case B200RL_I32: return (int)__ldg(static_cast<const int32_t*>(base) + i);
    default: return (int)__ldg(static_cast<const long long*>(base) + i);
  }
}

// ---- per-sample reductions over the CTA and the cluster --------------------------------------
__device__ __forceinline__ double shfl_xor_f64(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}

// e^x for x <= 0 through the SFU: one multiply + ex2.approx.ftz (2 ulp; inputs here are x - max <= 0,
// results below 2^-126 flush to zero, far below one ulp of the partition sum they are added to).
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// Online softmax statistics of a set of logits: m = max, s = sum e^(x-m), q = sum e^(x-m)(x-m).
// lse = m + log s, entropy = log s - q / s.  Partials merge associatively in a fixed order.
struct Soft {
  float m, s, q;
};
__device__ __noinline__ Soft soft_merge(const Soft& a, const Soft& b) {
  if (b.s == 0.f) return a;
  if (a.s == 0.f) return b;
  const float M = fmaxf(a.m, b.m);
  const float da = a.m - M, db = b.m - M;
  const float fa = fast_exp(da), fb = fast_exp(db);
  return Soft{M, a.s * fa + b.s * fb, fa * fmaf(a.s, da, a.q) + fb * fmaf(b.s, db, b.q)};
}
__device__ __forceinline__ Soft soft_push(const Soft& a, float x) { return soft_merge(a, Soft{x, 1.f, 0.f}); }

// What one warp / one CTA contributes to its sample.
struct SamplePart {
  double logp, ent;
  Soft pick[kMaxPick];
  float xa[kMaxPick];  // logit of the chosen pick cell (from the CTA that holds it)
};

// ---- zero fill / mask scan -----------------------------------------------------------------------
__device__ __forceinline__ void zero_fill(uint8_t* dst, uint32_t bytes) {
  const uint32_t tid = threadIdx.x;
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  if (tid < head) dst[tid] = 0;
  uint4* d4 = reinterpret_cast<uint4*>(dst + head);
  const uint32_t n4 = (bytes - head) >> 4;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint32_t i = tid;
  for (; i + 3u * kGridBlock < n4; i += 4u * kGridBlock) {
    d4[i] = z, d4[i + kGridBlock] = z, d4[i + 2u * kGridBlock] = z, d4[i + 3u * kGridBlock] = z;
  }
  for (; i < n4; i += kGridBlock) d4[i] = z;
  const uint32_t done = head + (n4 << 4);
  if (tid < bytes - done) dst[done + tid] = 0;
}

// Row prefetch context: the logits row of a flagged cell is wanted a few microseconds from now.
struct RowPrefetch {
  const uint8_t* logits;  // first logit of the CTA's first cell
  uint32_t row_bytes;     // Sp * sizeof(LT)
};

__device__ __forceinline__ void flag_cell(uint32_t* bitmap, uint32_t cell, const RowPrefetch& pf) {
  const uint32_t bit = 1u << (cell & 31);
  const uint32_t old = atomicOr(&bitmap[cell >> 5], bit);
  if (!(old & bit)) {  // first flag of this cell: pull its logits row towards L2
    const uint8_t* row = pf.logits + (size_t)cell * pf.row_bytes;
    for (uint32_t o = 0; o < pf.row_bytes + 127u; o += 128u) prefetch_l2(row + min(o, pf.row_bytes - 1u));
  }
}

// A 16-byte word of the mask chunk with at least one non-zero byte: flag the cell(s) it covers.
// Rows are S bytes, so for S >= 16 a word touches at most two cells; two 32-bit divisions and a
// byte-boundary split decide which.  (Chunk sizes are < 2^31 bytes: cells <= 4096, S <= 65535.)
__device__ __noinline__ void flag_word(uint32_t* bitmap, uint32_t base, const uint4& w, uint32_t S, const RowPrefetch& pf) {
  const uint32_t c0 = base / S, c1 = (base + 15u) / S;
  if (c0 == c1) {
    flag_cell(bitmap, c0, pf);
    return;
  }
  if (S >= 16u) {
    const uint32_t k = c1 * S - base;  // bytes [0, k) belong to c0, [k, 16) to c1; 1 <= k <= 15
    const unsigned long long lo = (unsigned long long)w.x | ((unsigned long long)w.y << 32);
    const unsigned long long hi = (unsigned long long)w.z | ((unsigned long long)w.w << 32);
    unsigned long long first, second;
    if (k < 8u) {
      first = lo & ((1ull << (8u * k)) - 1ull);
      second = (lo >> (8u * k)) | hi;
    } else {
      first = lo | (k == 8u ? 0ull : (hi & ((1ull << (8u * (k - 8u))) - 1ull)));
      second = hi >> (8u * (k - 8u));
    }
    if (first) flag_cell(bitmap, c0, pf);
    if (second) flag_cell(bitmap, c1, pf);
    return;
  }
  const uint32_t word[4] = {w.x, w.y, w.z, w.w};  // narrow rows: several cells per word
#pragma unroll 1
  for (int q = 0; q < 4; ++q)
#pragma unroll 1
    for (int r = 0; r < 4; ++r)
      if ((word[q] >> (8 * r)) & 0xffu) flag_cell(bitmap, (base + 4u * q + r) / S, pf);
}

// flags every cell of [mask, mask + bytes) that has a non-zero byte
__device__ __forceinline__ void scan_mask(const uint8_t* mask, uint32_t bytes, uint32_t S, uint32_t* bitmap,
                                          const RowPrefetch& pf) {
  const uint32_t tid = threadIdx.x;
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(mask) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  if (tid < head && mask[tid]) flag_cell(bitmap, tid / S, pf);
  const uint4* m4 = reinterpret_cast<const uint4*>(mask + head);
  const uint32_t n4 = (bytes - head) >> 4;
  constexpr uint32_t kUnroll = 4;
  for (uint32_t i0 = tid; i0 < n4; i0 += kGridBlock * kUnroll) {
    uint4 w[kUnroll];
#pragma unroll
    for (uint32_t u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * kGridBlock;
      // plain read-only loads (allocate in L1): the unit cells re-read their own mask bytes
      w[u] = i < n4 ? __ldg(m4 + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (uint32_t u = 0; u < kUnroll; ++u)
      if ((w[u].x | w[u].y | w[u].z | w[u].w) != 0u) flag_word(bitmap, head + ((i0 + u * kGridBlock) << 4), w[u], S, pf);
  }
  const uint32_t done = head + (n4 << 4);
  if (tid < bytes - done && mask[done + tid]) flag_cell(bitmap, (done + tid) / S, pf);
}

// warp 0: bitmap -> ascending list of flagged cells; returns the count through *s_count
__device__ __forceinline__ void compact_cells(const uint32_t* bitmap, int words, uint16_t* list, int* s_count) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int base = 0;
  for (int w0 = 0; w0 < words; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = w < words ? bitmap[w] : 0u;
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = base + incl - cnt;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      list[pos++] = (uint16_t)(w * 32 + b);
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) *s_count = base;
}

// ---- one piece of one head ------------------------------------------------------------------------
template <typename LT, int PMAX>
struct Piece {
  float x[PMAX];
  uint32_t valid;  // bit j: entry j of the piece is unmasked
};

template <typename LT, int PMAX>
__device__ __forceinline__ Piece<LT, PMAX> load_piece(const LT* row, const uint8_t* mrow, const LaneSlot& s) {
  Piece<LT, PMAX> p;
  p.valid = 0u;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) {
    p.x[j] = 0.f;
    if (j < s.len) {
      p.x[j] = to_f32(row[s.off + j]);
      if (mrow[s.off + j]) p.valid |= 1u << j;
    }
  }
  return p;
}

// xor-butterfly over the lanes of this head's block; lanes outside the block never mix in
// because blocks are width-aligned.  `steps` = log2(widest block in the plan), warp-uniform.
template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, int width, int max_width) {
  for (int o = 1; o < max_width; o <<= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, v, o);
    if (o < width) v = IS_MAX ? fmaxf(v, other) : v + other;
  }
  return v;
}

// Forward of one head over the lanes of its block.  Leaves, per lane, d_j = x_j - max and
// e_j = exp(d_j) of its own valid entries (0 for masked ones) so that the backward can form
// p_j = e_j / sum and log p_j = d_j - log(sum) without touching memory or the SFU again.
struct HeadStat {
  float mx, ls, inv_sum, ent;  // max, log(sum e), 1 / sum e, entropy
  bool any;
};

template <typename LT, int PMAX>
__device__ __forceinline__ HeadStat head_forward(const Piece<LT, PMAX>& p, const LaneSlot& s, int max_width,
                                                 float (&d)[PMAX], float (&e)[PMAX]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) mx = fmaxf(mx, ((p.valid >> j) & 1u) ? p.x[j] : -INFINITY);
  mx = block_reduce<true>(mx, s.width, max_width);
  HeadStat h;
  h.any = mx > -INFINITY;
  h.mx = h.any ? mx : 0.f;
  float sum = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < PMAX; ++j) {
    const bool v = (p.valid >> j) & 1u;
    d[j] = v ? p.x[j] - h.mx : 0.f;
    e[j] = v ? fast_exp(d[j]) : 0.f;
    sum += e[j];
    q = fmaf(e[j], d[j], q);
  }
  sum = block_reduce<false>(sum, s.width, max_width);
  q = block_reduce<false>(q, s.width, max_width);
  h.ls = 0.f, h.inv_sum = 0.f, h.ent = 0.f;
  if (h.any) {
    h.ls = logf(sum);
    h.inv_sum = 1.f / sum;
    h.ent = h.ls - q * h.inv_sum;  // -sum p * logp over the valid entries
  }
  return h;
}

// d loss / d logits of this lane's piece, written over the zero-filled row.
template <typename LT, int PMAX>
__device__ __forceinline__ void piece_backward(LT* out, const LaneSlot& s, uint32_t valid, const float (&d)[PMAX],
                                               const float (&e)[PMAX], float ls, float inv_sum, float ent, int local,
                                               float dl, float dent) {
#pragma unroll
  for (int j = 0; j < PMAX; ++j) {
    if (j < s.len && ((valid >> j) & 1u)) {
      const float lp = d[j] - ls;
      const float pr = e[j] * inv_sum;
      out[j] = from_f32<LT>(dl * ((j == local ? 1.f : 0.f) - pr) - dent * pr * (lp + ent));
    }
  }
}

// ---- the kernel --------------------------------------------------------------------------------
template <int MODE, typename LT, int PMAX, bool PICK>
__global__ void __launch_bounds__(kGridBlock, PMAX <= 8 ? 8 : 4) gridnet_kernel(const __grid_constant__ GridDev G,
                                                             const __grid_constant__ PpoDev P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ double s_wsum[2][kGridBlock / 32];        // per-warp (logp, entropy)
  __shared__ Soft s_wpick[kMaxPick][kGridBlock / 32];  // per-warp pick statistics
  __shared__ SamplePart s_cta;                         // this CTA's record, read by its cluster peers
  __shared__ float s_bcast[2];
  __shared__ float s_pick[kMaxPick * 3];  // lse, entropy, any per pick head
  __shared__ int s_count;

  constexpr int NP = PICK ? kMaxPick : 1;  // pick heads compiled in (PICK == false: none)
  const int n_pick = PICK ? G.n_pick : 0;
  const int tid = threadIdx.x;
  const int cluster_size = G.cluster;
  const int rank = cluster_size > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int cells = G.cells_per_cta;
  const long long cell0 = (long long)rank * cells;  // first cell of this CTA within a sample
  const long long n_clusters = gridDim.x / cluster_size;
  const uint32_t out_bytes = (uint32_t)cells * (uint32_t)G.Sp * (uint32_t)sizeof(LT);
  const uint32_t mask_bytes = (uint32_t)cells * (uint32_t)G.S;

  // dynamic shared memory: bitmap | unit-cell list | per (unit cell, head) lse and entropy
  const int words = (cells + 31) >> 5;
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem);
  uint16_t* list = reinterpret_cast<uint16_t*>(bitmap + words);
  float* s_lse = reinterpret_cast<float*>(smem + (((size_t)words * 4 + (size_t)cells * 2 + 15) & ~(size_t)15));
  float* s_ent = s_lse + (size_t)cells * G.A;

  // A CTA (cluster) handles sample b, then b + n_clusters, ... (normally exactly one: the grid is
  // one cluster per sample).  The streaming half of a sample (zero fill of its dlogits, its mask
  // bytes on their way to L2) is issued before the latency-bound half of the previous one.
  auto stream_ahead = [&](long long bn) {
    const long long r0 = bn * G.HW + cell0;
    const uint8_t* m = G.mask + r0 * G.S;
    for (uint32_t o = (uint32_t)tid * 128u; o < mask_bytes; o += kGridBlock * 128u) prefetch_l2(m + o);
    if (MODE != kFwd) zero_fill(reinterpret_cast<uint8_t*>(static_cast<LT*>(G.dlogits) + r0 * G.Sp), out_bytes);
  };
  const long long b_first = blockIdx.x / cluster_size;
  if (b_first < G.B) stream_ahead(b_first);

  for (long long b = b_first; b < G.B; b += n_clusters) {
  const long long row0 = b * G.HW + cell0;  // global index of this CTA's first cell
  const LT* g_logits = static_cast<const LT*>(G.logits) + row0 * G.Sp;
  LT* g_out = static_cast<LT*>(G.dlogits) + row0 * G.Sp;
  const uint8_t* g_mask = G.mask + row0 * G.S;

  // the per-sample PPO scalars are consumed by one thread after the reductions: start pulling them now
  if (MODE == kPpo && rank == 0) {
    if (tid == 0) {
      prefetch_l1(P.old_logp + b);
      prefetch_l1(P.adv + b * P.adv_v);
      if (P.adv_mode) prefetch_l1(P.norm);
    } else if (tid >= 32 && tid < 32 + P.V) {
      const long long o = b * P.V + (tid - 32);
      prefetch_l1(P.new_values + o), prefetch_l1(P.old_values + o), prefetch_l1(P.returns + o);
    }
  }

  // ---- 1. mask scan (this sample's zero fill was issued one iteration ago) ------------------------
  for (int w = tid; w < words; w += kGridBlock) bitmap[w] = 0u;
  __syncthreads();
  scan_mask(g_mask, mask_bytes, (uint32_t)G.S, bitmap,
            RowPrefetch{reinterpret_cast<const uint8_t*>(g_logits), (uint32_t)G.Sp * (uint32_t)sizeof(LT)});
  __syncthreads();
  if (b + n_clusters < G.B) stream_ahead(b + n_clusters);

  // ---- 2. compaction -------------------------------------------------------------------------------
  compact_cells(bitmap, words, list, &s_count);
  __syncthreads();
  const int n_unit = s_count;

  // ---- 3. forward over the unit cells -----------------------------------------------------------------
  const int group = tid / G.G, n_groups = kGridBlock / G.G;
  const LaneSlot slot = G.slot[tid & (G.G - 1)];
  const int gate_ref = slot.len ? G.gate_ref[slot.head] : -1;
  const int gate_val = slot.len ? G.gate_val[slot.head] : 0;
  double logp_acc = 0.0, ent_acc = 0.0;
  // state of the first pass stays in registers for the backward (a CTA rarely has more unit cells
  // than lane groups); later passes park (log-sum-exp, entropy) in shared memory and reload.
  float k_d[PMAX], k_e[PMAX];
  uint32_t k_valid = 0u;
  float k_ls = 0.f, k_inv = 0.f, k_ent = 0.f;
  int k_cell = 0, k_local = -1;
  bool k_gated = false;
#pragma unroll 1
  for (int i0 = 0; i0 < n_unit; i0 += n_groups) {  // uniform trip count: shuffles stay converged
    const int i = i0 + group;
    const bool live = i < n_unit && slot.len > 0;
    const int cell = live ? (int)list[i] : 0;
    Piece<LT, PMAX> p;
    p.valid = 0u;
#pragma unroll
    for (int j = 0; j < PMAX; ++j) p.x[j] = 0.f;
    int a_head = 0, a_ref = gate_val;
    if (live) {
      p = load_piece<LT, PMAX>(g_logits + (long long)cell * G.Sp, g_mask + (long long)cell * G.S, slot);
      const long long abase = (row0 + cell) * G.A;
      a_head = load_index(G.actions, G.act_dtype, abase + slot.head);
      if (gate_ref >= 0) a_ref = load_index(G.actions, G.act_dtype, abase + gate_ref);
    }
    float d[PMAX], e[PMAX];
    const HeadStat h = head_forward<LT, PMAX>(p, slot, G.max_width, d, e);
    if (!live) continue;
    const bool gated_in = a_ref == gate_val;
    const int local = a_head - (int)(slot.off - slot.head_off);
    if (slot.first && h.any) ent_acc += (double)h.ent;
    if (h.any && gated_in && local >= 0 && local < slot.len) {
      float da = kF32Lowest;  // a masked action keeps the reference's finfo.min logit
#pragma unroll
      for (int j = 0; j < PMAX; ++j)
        if (j == local && ((p.valid >> j) & 1u)) da = d[j];
      logp_acc += (double)(da == kF32Lowest ? kF32Lowest - (h.mx + h.ls) : da - h.ls);
    }
    if (i0 == 0) {
#pragma unroll
      for (int j = 0; j < PMAX; ++j) k_d[j] = d[j], k_e[j] = e[j];
      k_valid = h.any ? p.valid : 0u, k_ls = h.ls, k_inv = h.inv_sum, k_ent = h.ent;
      k_cell = cell, k_local = local, k_gated = gated_in;
    } else if (slot.first) {
      s_lse[i * G.A + slot.head] = h.any ? h.mx + h.ls : INFINITY;  // +inf marks a head with no valid entry
      s_ent[i * G.A + slot.head] = h.ent;
    }
  }

  // ---- 4. pick_position: one online-softmax pass over this CTA's cells ---------------------------------
  // (m, s, q) = running max, sum e^(x-m), sum e^(x-m)(x-m) over the valid cells; partials merge
  // associatively, so the whole sample needs a single exchange (section 5).
  Soft pick[NP];
  float pick_xa[NP];
#pragma unroll
  for (int kp = 0; kp < NP; ++kp) {
    pick[kp] = Soft{-INFINITY, 0.f, 0.f};
    pick_xa[kp] = 0.f;
    if (kp >= n_pick) continue;
    const uint8_t* pm = G.pick_mask + (b * n_pick + kp) * G.HW + cell0;
#pragma unroll 1
    for (int c = tid; c < cells; c += kGridBlock)
      if (pm[c]) pick[kp] = soft_push(pick[kp], to_f32(g_logits[(long long)c * G.Sp + G.S + kp]));
    if (MODE != kBwd && tid == 0) {  // the CTA that holds the chosen cell contributes its logit
      const long long local = (long long)load_index(G.pick_actions, G.pick_dtype, b * n_pick + kp) - cell0;
      if (local >= 0 && local < cells) pick_xa[kp] = pm[local] ? to_f32(g_logits[local * G.Sp + G.S + kp]) : kF32Lowest;
    }
  }

  // ---- 5. one exchange: warp partials -> CTA record -> (cluster) -> sample totals --------------------------
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      logp_acc += shfl_xor_f64(logp_acc, o);
      ent_acc += shfl_xor_f64(ent_acc, o);
#pragma unroll
      for (int kp = 0; kp < NP; ++kp)
        if (kp < n_pick) {
          Soft other;
          other.m = __shfl_xor_sync(0xffffffffu, pick[kp].m, o);
          other.s = __shfl_xor_sync(0xffffffffu, pick[kp].s, o);
          other.q = __shfl_xor_sync(0xffffffffu, pick[kp].q, o);
          pick[kp] = soft_merge(pick[kp], other);
        }
    }
    if (lane == 0) {
      s_wsum[0][warp] = logp_acc, s_wsum[1][warp] = ent_acc;
#pragma unroll
      for (int kp = 0; kp < NP; ++kp)
        if (kp < n_pick) s_wpick[kp][warp] = pick[kp];
    }
  }
  __syncthreads();
  float dlogp = 0.f, dent = 0.f;
  double tot_logp = 0.0, tot_ent = 0.0;
  if (tid == 0) {  // CTA totals = the eight warp partials, in warp order
#pragma unroll
    for (int w = 0; w < kGridBlock / 32; ++w) tot_logp += s_wsum[0][w], tot_ent += s_wsum[1][w];
#pragma unroll
    for (int kp = 0; kp < NP; ++kp) {
      if (kp >= n_pick) continue;
      Soft t = s_wpick[kp][0];
#pragma unroll 1
      for (int w = 1; w < kGridBlock / 32; ++w) t = soft_merge(t, s_wpick[kp][w]);
      s_cta.pick[kp] = t, s_cta.xa[kp] = pick_xa[kp];
    }
    if (cluster_size > 1) s_cta.logp = tot_logp, s_cta.ent = tot_ent;
  }
  if (cluster_size > 1) cg::this_cluster().sync();  // every CTA's record is published (all threads arrive)
  if (tid == 0) {
    if (cluster_size > 1) {
      cg::cluster_group cluster = cg::this_cluster();
      tot_logp = 0.0, tot_ent = 0.0;
      for (int r = 0; r < cluster_size; ++r) {
        const SamplePart* o = cluster.map_shared_rank(&s_cta, r);
        tot_logp += o->logp, tot_ent += o->ent;
      }
    }
    for (int kp = 0; kp < n_pick; ++kp) {
      Soft t = s_cta.pick[kp];
      float xa = s_cta.xa[kp];
      if (cluster_size > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        t = cluster.map_shared_rank(&s_cta, 0)->pick[kp], xa = cluster.map_shared_rank(&s_cta, 0)->xa[kp];
#pragma unroll 1
        for (int r = 1; r < cluster_size; ++r) {
          const SamplePart* o = cluster.map_shared_rank(&s_cta, r);
          t = soft_merge(t, o->pick[kp]);
          xa += o->xa[kp];  // exactly one CTA holds the chosen cell
        }
      }
      const bool any = t.s > 0.f;
      float p_lse = 0.f, p_ent = 0.f;
      if (any) {
        const float ls = logf(t.s);
        p_lse = t.m + ls;
        p_ent = ls - t.q / t.s;
        tot_logp += (double)(xa - p_lse);
        tot_ent += (double)p_ent;
      }
      s_pick[kp * 3] = p_lse, s_pick[kp * 3 + 1] = p_ent, s_pick[kp * 3 + 2] = any ? 1.f : 0.f;
    }
    if (MODE == kFwd) {
      if (rank == 0) G.logp[b] = (float)tot_logp, G.entropy[b] = (float)tot_ent;
    } else if (MODE == kBwd) {
      s_bcast[0] = G.dlogp_in[b], s_bcast[1] = G.dent_in[b];
    } else {
      // ---- 6. PPO scalar stage (every CTA of the cluster derives the same dlogp; rank 0 records) ----------
      PolicyTerms t = ppo_policy_terms(P, b, tot_logp);
      s_bcast[0] = t.dlogp, s_bcast[1] = ppo_dentropy(P, 1);
      if (rank == 0) {
        double* row = P.partials + b * ppo_nstat(P.V);
        row[0] = t.surrogate, row[1] = tot_ent, row[2] = t.kl, row[3] = t.clipped;
        if (G.logp) G.logp[b] = (float)tot_logp;
        if (G.entropy) G.entropy[b] = (float)tot_ent;
      }
    }
  }
  if (MODE == kPpo && rank == 0 && tid >= 32 && tid < 32 + P.V) {
    const int v = tid - 32;
    float2 r = ppo_value_terms(P, b, v);
    double* row = P.partials + b * ppo_nstat(P.V);
    row[kPolicyStats + v] = r.x, row[kPolicyStats + P.V + v] = r.y;
  }
  __syncthreads();
  if (MODE == kFwd) {
    if (cluster_size > 1) cg::this_cluster().sync();  // peers may still be reading this CTA's record
    continue;
  }
  dlogp = s_bcast[0], dent = s_bcast[1];

  // ---- 7. backward over the unit cells: overwrite their zero-filled rows ------------------------------------
  if (k_valid)
    piece_backward<LT, PMAX>(g_out + (long long)k_cell * G.Sp + slot.off, slot, k_valid, k_d, k_e, k_ls, k_inv, k_ent,
                             k_local, k_gated ? dlogp : 0.f, dent);
#pragma unroll 1
  for (int i = n_groups + group; i < n_unit; i += n_groups) {  // cells beyond the first pass
    if (slot.len == 0) continue;
    const float lse = s_lse[i * G.A + slot.head];
    if (lse == INFINITY) continue;  // no valid entry in this head: gradient stays zero
    const float ent = s_ent[i * G.A + slot.head];
    const int cell = (int)list[i];
    const Piece<LT, PMAX> p = load_piece<LT, PMAX>(g_logits + (long long)cell * G.Sp, g_mask + (long long)cell * G.S, slot);
    const long long abase = (row0 + cell) * G.A;
    const bool gated_in = gate_ref < 0 || load_index(G.actions, G.act_dtype, abase + gate_ref) == gate_val;
    const int local = load_index(G.actions, G.act_dtype, abase + slot.head) - (int)(slot.off - slot.head_off);
    float d[PMAX], e[PMAX];
#pragma unroll
    for (int j = 0; j < PMAX; ++j) {
      const bool v = (p.valid >> j) & 1u;
      d[j] = v ? p.x[j] - lse : 0.f;  // log p_j directly
      e[j] = v ? fast_exp(d[j]) : 0.f;
    }
    piece_backward<LT, PMAX>(g_out + (long long)cell * G.Sp + slot.off, slot, p.valid, d, e, 0.f, 1.f, ent, local,
                             gated_in ? dlogp : 0.f, dent);
  }
  for (int kp = 0; kp < n_pick; ++kp) {
    if (s_pick[kp * 3 + 2] == 0.f) continue;
    const float p_lse = s_pick[kp * 3], p_ent = s_pick[kp * 3 + 1];
    const uint8_t* pm = G.pick_mask + (b * n_pick + kp) * G.HW + cell0;
    const long long a = (long long)load_index(G.pick_actions, G.pick_dtype, b * n_pick + kp) - cell0;
    for (int c = tid; c < cells; c += kGridBlock)
      if (pm[c]) {
        const float lp = to_f32(g_logits[(long long)c * G.Sp + G.S + kp]) - p_lse;
        const float pr = expf(lp);
        g_out[(long long)c * G.Sp + G.S + kp] = from_f32<LT>(dlogp * ((a == c ? 1.f : 0.f) - pr) - dent * pr * (lp + p_ent));
      }
  }
  // the next sample reuses the bitmap / list / stash, and peers may still be reading this CTA's record
  if (cluster_size > 1) cg::this_cluster().sync();
  else __syncthreads();
  }  // samples
}

// ---- host side -----------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Lay the heads out over the lanes of a group: head h is cut into pieces of <= pmax logits that
// occupy a width-aligned power-of-two block of lanes; wide blocks first so that alignment is free.
static bool plan_lanes(GridDev* G, const int* nvec, int pmax) {
  struct Blk { int head, pieces, width, piece_len; };
  Blk blk[B200RL_MAX_HEADS];
  int off[B200RL_MAX_HEADS], S = 0;
  for (int h = 0; h < G->A; ++h) {
    off[h] = S, S += nvec[h];
    const int pieces = (nvec[h] + pmax - 1) / pmax;
    blk[h] = Blk{h, pieces, pow2_ceil(pieces), (nvec[h] + pieces - 1) / pieces};
  }
  for (int i = 0; i < G->A; ++i)  // stable insertion sort by width, descending
    for (int j = i; j > 0 && blk[j].width > blk[j - 1].width; --j) {
      Blk t = blk[j]; blk[j] = blk[j - 1]; blk[j - 1] = t;
    }
  int lanes = 0, max_width = 1;
  for (int i = 0; i < G->A; ++i) lanes += blk[i].width, max_width = blk[i].width > max_width ? blk[i].width : max_width;
  if (lanes > 32) return false;
  G->G = pow2_ceil(lanes), G->max_width = max_width;
  for (int l = 0; l < 32; ++l) G->slot[l] = LaneSlot{0, 0, 0, 0, 1, 0};
  int lane = 0;
  for (int i = 0; i < G->A; ++i) {
    const Blk& k = blk[i];
    for (int q = 0; q < k.width; ++q, ++lane) {
      LaneSlot s{0, (uint16_t)off[k.head], 0, (uint8_t)k.head, (uint8_t)k.width, (uint8_t)(q == 0)};
      const int begin = q * k.piece_len;
      if (q < k.pieces && begin < nvec[k.head]) {
        const int len = nvec[k.head] - begin < k.piece_len ? nvec[k.head] - begin : k.piece_len;
        s.off = (uint16_t)(off[k.head] + begin), s.len = (uint8_t)len;
      } else {
        s.off = (uint16_t)off[k.head];  // idle lane of the block: takes part in the shuffles only
      }
      G->slot[lane] = s;
    }
  }
  return true;
}

struct GridLaunch {
  int cluster;
  int cells_per_cta;
  size_t smem;
};

static size_t grid_smem(int cells, int A) {
  const size_t words = (size_t)(cells + 31) / 32;
  return ((words * 4 + (size_t)cells * 2 + 15) & ~(size_t)15) + (size_t)cells * A * 2 * sizeof(float);
}

static int plan_launch(const GridDev& G, GridLaunch* out) {
  // spread a sample over a cluster until a CTA holds <= 512 cells (one or two cells per thread of
  // scan work, ~25 KB of shared memory); larger maps fall back to the biggest portable cluster
  int cs = 1;
  while (cs < 8 && (G.HW % (cs * 2) == 0) && G.HW / cs > 512) cs *= 2;
  const long long cells = G.HW / cs;
  if (cells > kMaxCellsPerCta) {
    set_error("gridnet: HW=%lld cells needs %lld cells per CTA (max %d)", G.HW, cells, kMaxCellsPerCta);
    return B200RL_EUNSUPPORTED;
  }
  *out = GridLaunch{cs, (int)cells, grid_smem((int)cells, G.A)};
  return B200RL_OK;
}

template <int MODE, typename LT, int PMAX, bool PICK>
static int launch_one(GridDev& G, const PpoDev& P, const GridLaunch& L, cudaStream_t stream) {
  auto kernel = gridnet_kernel<MODE, LT, PMAX, PICK>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
  if (e != cudaSuccess) {
    set_error("gridnet: cudaFuncSetAttribute(%zu bytes): %s", L.smem, cudaGetErrorString(e));
    return B200RL_ECUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.cluster);
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
  // One cluster per sample while the grid fits (the hardware scheduler overlaps CTAs at different
  // phases better than a persistent loop does: measured 116 vs 136 us on the C4 minibatch); beyond
  // 2^31 - 1 CTAs the kernel's sample loop takes over.
  long long ctas = G.B * L.cluster;
  const long long max_ctas = 0x7fffffffLL - (0x7fffffffLL % L.cluster);
  if (ctas > max_ctas) ctas = max_ctas;
  cfg.gridDim = dim3((unsigned)ctas);
  e = cudaLaunchKernelEx(&cfg, kernel, (const GridDev)G, P);
  if (e != cudaSuccess) {
    set_error("gridnet launch (cluster %d, %zu B smem): %s", L.cluster, L.smem, cudaGetErrorString(e));
    return B200RL_ECUDA;
  }
  return B200RL_OK;
}

template <int MODE>
static int launch_mode(GridDev& G, const PpoDev& P, const int* nvec, int logits_dtype, cudaStream_t stream) {
  GridLaunch L;
  int rc = plan_launch(G, &L);
  if (rc) return rc;
  G.cluster = L.cluster, G.cells_per_cta = L.cells_per_cta;
  const bool bf16 = logits_dtype == B200RL_BF16;
  if (plan_lanes(&G, nvec, 8)) {
    if (G.n_pick == 0)
      return bf16 ? launch_one<MODE, __nv_bfloat16, 8, false>(G, P, L, stream)
                  : launch_one<MODE, float, 8, false>(G, P, L, stream);
    return bf16 ? launch_one<MODE, __nv_bfloat16, 8, true>(G, P, L, stream)
                : launch_one<MODE, float, 8, true>(G, P, L, stream);
  }
  if (plan_lanes(&G, nvec, 32)) {
    return bf16 ? launch_one<MODE, __nv_bfloat16, 32, true>(G, P, L, stream)
                : launch_one<MODE, float, 32, true>(G, P, L, stream);
  }
  set_error("gridnet: the action planes do not fit one warp (32 lanes x 32 logits)");
  return B200RL_EUNSUPPORTED;
}

static int make_grid(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask, const uint8_t* pick_mask,
                     const void* actions, const void* pick_actions, GridDev* out, const char* who) {
  B200RL_REQUIRE(d && logits && mask, "%s: null pointer", who);
  B200RL_REQUIRE(d->B >= 0 && d->HW >= 1 && d->A >= 1 && d->n_pick >= 0, "%s: bad shape", who);
  B200RL_REQUIRE(d->nvec_host != nullptr, "%s: nvec is null", who);
  B200RL_UNSUPPORTED(d->A > B200RL_MAX_HEADS, "%s: A=%d action planes (max %d)", who, d->A, B200RL_MAX_HEADS);
  B200RL_UNSUPPORTED(d->n_pick > kMaxPick, "%s: n_pick=%d exceeds %d", who, d->n_pick, kMaxPick);
  B200RL_UNSUPPORTED(d->logits_dtype != B200RL_F32 && d->logits_dtype != B200RL_BF16, "%s: logits dtype %d", who,
                     d->logits_dtype);
  B200RL_REQUIRE(d->n_pick == 0 || pick_mask, "%s: pick_mask is null", who);
  B200RL_REQUIRE(actions != nullptr, "%s: actions is null", who);
  B200RL_REQUIRE(d->n_pick == 0 || pick_actions, "%s: pick_actions is null", who);
  B200RL_UNSUPPORTED(d->act_dtype != B200RL_U8 && d->act_dtype != B200RL_I32 && d->act_dtype != B200RL_I64,
                     "%s: action dtype %d", who, d->act_dtype);
  B200RL_UNSUPPORTED(d->n_pick > 0 && d->pick_dtype != B200RL_I32 && d->pick_dtype != B200RL_I64,
                     "%s: pick action dtype %d", who, d->pick_dtype);
  GridDev G{};
  G.logits = logits, G.mask = mask, G.pick_mask = pick_mask, G.actions = actions, G.pick_actions = pick_actions;
  G.B = d->B, G.HW = d->HW, G.A = d->A, G.n_pick = d->n_pick;
  G.act_dtype = d->act_dtype, G.pick_dtype = d->pick_dtype;
  int S = 0;
  for (int h = 0; h < d->A; ++h) {
    B200RL_REQUIRE(d->nvec_host[h] >= 1, "%s: nvec[%d]=%d", who, h, d->nvec_host[h]);
    B200RL_UNSUPPORTED(d->nvec_host[h] > 1024, "%s: nvec[%d]=%d exceeds 1024", who, h, d->nvec_host[h]);
    S += d->nvec_host[h];
    const int gr = d->gate_ref_host ? d->gate_ref_host[h] : -1;
    B200RL_REQUIRE(gr < d->A, "%s: gate_ref[%d]=%d out of range", who, h, gr);
    G.gate_ref[h] = gr;
    G.gate_val[h] = (gr >= 0 && d->gate_val_host) ? d->gate_val_host[h] : 0;
  }
  B200RL_UNSUPPORTED(S > 65535, "%s: sum(nvec)=%d exceeds 65535", who, S);
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
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, "gridnet_fwd");
  if (rc) return rc;
  B200RL_REQUIRE(logp && entropy, "gridnet_fwd: null output");
  if (G.B == 0) return B200RL_OK;
  G.logp = logp, G.entropy = entropy;
  PpoDev P{};
  return launch_mode<kFwd>(G, P, d->nvec_host, d->logits_dtype, (cudaStream_t)stream);
}

extern "C" int b200rl_gridnet_bwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                  const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                  const float* dlogp, const float* dentropy, void* dlogits,
                                  b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, "gridnet_bwd");
  if (rc) return rc;
  B200RL_REQUIRE(dlogp && dentropy && dlogits, "gridnet_bwd: null pointer");
  if (G.B == 0) return B200RL_OK;
  G.dlogp_in = dlogp, G.dent_in = dentropy, G.dlogits = dlogits;
  PpoDev P{};
  return launch_mode<kBwd>(G, P, d->nvec_host, d->logits_dtype, (cudaStream_t)stream);
}

extern "C" int b200rl_ppo_gridnet_loss(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                       const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                       const b200rl_ppo_args* args, void* dlogits, float* logp_out,
                                       float* entropy_out, void* workspace, size_t workspace_bytes,
                                       b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, "ppo_gridnet_loss");
  if (rc) return rc;
  B200RL_REQUIRE(dlogits != nullptr, "ppo_gridnet_loss: dlogits is null");
  PpoDev P;
  rc = ppo_make_dev(args, G.B, workspace, workspace_bytes, &P);
  if (rc) return rc;
  G.dlogits = dlogits, G.logp = logp_out, G.entropy = entropy_out;
  cudaStream_t s = (cudaStream_t)stream;
  rc = ppo_launch_prepare(P, s);
  if (rc) return rc;
  rc = launch_mode<kPpo>(G, P, d->nvec_host, d->logits_dtype, s);
  if (rc) return rc;
  return ppo_launch_finalize(P, G.B, 1, s);
}

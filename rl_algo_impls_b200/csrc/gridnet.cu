// K4c: GridNet per-cell MultiDiscrete heads (+ pick_position) -- forward, backward and the
// fully fused PPO loss.
//
// Replaces shared/actor/gridnet.py:38-193 over shared/actor/categorical.py:12-54 and, in the
// fused mode, ppo/ppo.py:326-361 with the autograd backward of the whole chain: ~60-80
// eager launches forward and twice that backward in the reference.
//
// The work is driven by the action MASK.  A cell whose S mask bytes are all zero (no unit on
// it: ~94-98 % of the cells of a MicroRTS / Lux map) contributes exactly 0 to the log-prob,
// to the entropy and to the gradient (reference semantics: a fully masked row normalises to
// uniform, log_prob 0, entropy -0, gradient blocked by torch.where), so its logits are never
// read and its dlogits row is plain zero fill.  Two phases:
//
//  S  streaming, one chunk of 256 cells at a time: zero-fills the chunk's dlogits with 128-bit
//     stores, scans its mask bytes (and the pick_position mask) with 128-bit loads, flags non-empty
//     cells in a shared bitmap (the first flag prefetches that cell's logits row to L2) and
//     compacts the bitmap into the chunk's ascending list of unit cells (+ count).  HBM-bound.
//  C  fused compute, one CTA per sample -- masked logsumexp / log-prob / entropy, ratio / clip /
//     value-clip / KL, and the backward into dlogits and dvalues in ONE launch: groups of G lanes
//     (G = 16 for MicroRTS' planes, 8 for Lux') take one listed cell each; a lane owns <= 8
//     adjacent logits of one head, a wide head (the 49-way attack target)
//     spans an aligned power-of-two block of lanes reduced with xor shuffles; per-sample sums
//     run in float64; thread 0 forms the PPO terms (ppo_terms.cuh), one warp the value heads;
//     the same lanes then overwrite the listed cells' zero-filled rows with d loss / d logits
//     straight from registers.  The pick_position categorical over the (compacted) valid cells
//     is an online-softmax reduction in the same pass.  Latency-bound, but on ~5 % of the cells.
//
// Maps of up to 256 cells (MicroRTS 16x16: one chunk per sample) run S and C in the SAME 128-thread
// CTA of one launch: the lists stay in shared memory, the zero-fill stores of a sample drain while
// its compute phase runs and CTAs at different phases share an SM.  Larger maps (Lux 64x64: 16
// chunks per sample) run S as its own launch (one CTA per chunk, lists to the workspace) and C as a
// second launch: measured 130 us against 176-200 us for single-kernel variants (an 8-CTA cluster
// exchanging through DSMEM; "the CTA that finishes a sample's last chunk computes", whose
// __threadfence before the arrival counter serialises the zero-fill stores), B=512.
//
// HBM traffic per sample: dlogits written once, masks read once, logits / actions read only
// for unit cells.  Deterministic: lists are ascending per chunk, every reduction has a fixed order.
#include "mask_scan.cuh"
#include "ppo_terms.cuh"

namespace b200rl {

constexpr int kStreamBlock = 256;   // threads of the streaming kernel
constexpr int kMaxPick = 4;
constexpr int kStashCells = 512;    // unit cells whose (lse, entropy) are parked in shared memory
constexpr int kStashCellsSelf = 64;  // ... in a self-streaming CTA (maps of <= 256 cells)
constexpr size_t kMaxImageBytes = 40 * 1024;  // widest chunk mask (256 cells x S bytes) staged through shared memory

enum GridMode { kFwd = 0, kBwd = 1, kPpo = 2 };
enum RowsMode { kRowsOff = 0, kRowsRecord = 1, kRowsClear = 2 };

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
  int ld;            // elements between consecutive cells' rows of logits / dlogits (>= Sp; padded NHWC heads)
  int rows_mode;     // kRowsOff: dlogits is zero-filled; kRowsRecord: ... and the written rows are recorded in the
                     // lists; kRowsClear: the lists hold the rows of the previous call -- only those are cleared
  int act_dtype, pick_dtype;
  int gate_ref[B200RL_MAX_HEADS], gate_val[B200RL_MAX_HEADS];
  float* logp;
  float* entropy;
  const float* dlogp_in;
  const float* dent_in;
  int chunks;        // streaming CTAs per sample = ceil(HW / kChunkCells)
  int stash;         // unit cells with a shared-memory (lse, entropy) slot = min(HW, kStashCells)
  int G;             // lanes per cell group (power of two <= 32)
  int max_width;     // widest head block
  int image_bytes;   // shared-memory mask image of a streaming CTA (0: masks too wide, plain loads)
  // workspace (written by the streaming kernel, read by the compute kernel)
  uint16_t* unit_list;   // [B][chunks * kChunkCells]  ascending cell ids per chunk
  int* unit_count;       // [B][chunks]
  uint16_t* pick_list;   // [B][n_pick][chunks * kChunkCells]
  int* pick_count;       // [B][n_pick][chunks]
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

// ---- S: streaming phase ---------------------------------------------------------------------------
// Where a chunk's lists go: the workspace (split launches) or the CTA's own shared memory.
struct ChunkLists {
  uint16_t* unit;      // this chunk's unit-cell list
  int* unit_count;
  uint16_t* pick;      // pick head 0's list; head kp is at pick + kp * pick_stride
  int* pick_count;     // head kp's count is at pick_count + kp * pick_count_stride
  long long pick_stride, pick_count_stride;
};

// Shared-memory staging of a streaming CTA.  `image` == nullptr selects the plain load / store path (mask
// rows too wide for the staging buffer).
struct StreamStage {
  uint8_t* image;   // 16-byte aligned, >= chunk mask bytes + 32
  uint8_t* zeros;   // kZeroBuf bytes, 16-byte aligned
  uint64_t* bar;    // mbarrier the mask image completes on (phase 0; one chunk per CTA)
  uint32_t* pick_bitmap;  // [kMaxPick][kChunkCells / 32] words: valid pick cells of the chunk, per pick head
};

// The rows a previous call wrote into a persistent dlogits buffer (b200rl_ppo_gridnet_loss_inplace): the unit cells of
// this chunk's previous list get their whole row cleared, the previous pick lists their pick column.  Generic stores
// by the whole CTA; the caller separates them from this call's gradient stores with a barrier / a grid dependency.
template <typename LT, int BLOCK>
__device__ __forceinline__ void clear_previous_rows(const GridDev& G, long long b, const ChunkLists& prev, int n,
                                                    const int (&np)[kMaxPick]) {
  LT* out = static_cast<LT*>(G.dlogits) + b * G.HW * G.ld;
  const uint32_t row_bytes = (uint32_t)G.ld * (uint32_t)sizeof(LT);
  if (((reinterpret_cast<uintptr_t>(out) | row_bytes) & 15u) == 0) {
    // padded rows (ld = 80 / 32: 16-byte multiples, 16-byte aligned): whole rows with 128-bit stores
    const uint32_t q = row_bytes >> 4;  // 16-byte words per row
    for (uint32_t i = threadIdx.x; i < (uint32_t)n * q; i += BLOCK) {
      const uint32_t k = i / q;
      reinterpret_cast<uint4*>(out + (long long)prev.unit[k] * G.ld)[i - k * q] = make_uint4(0u, 0u, 0u, 0u);
    }
  } else {
    const uint32_t Sp = (uint32_t)G.Sp;
    for (uint32_t i = threadIdx.x; i < (uint32_t)n * Sp; i += BLOCK) {
      const uint32_t k = i / Sp;
      out[(long long)prev.unit[k] * G.ld + (i - k * Sp)] = from_f32<LT>(0.f);
    }
  }
  for (int kp = 0; kp < G.n_pick; ++kp) {
    const uint16_t* pl = prev.pick + kp * prev.pick_stride;
    for (int i = threadIdx.x; i < np[kp]; i += BLOCK) out[(long long)pl[i] * G.ld + G.S + kp] = from_f32<LT>(0.f);
  }
}

// Returns the skew of the mask image (byte k of the chunk's mask is image[skew + k]).
// `ZERO`: the mode writes dlogits.  With G.rows_mode == kRowsClear the rows listed in `out` (the previous call's)
// are cleared instead of the whole chunk being filled; `out` is then overwritten with this call's lists.
template <typename LT, bool ZERO, int BLOCK>
__device__ __forceinline__ uint32_t stream_chunk(const GridDev& G, long long b, int chunk, uint32_t* bitmap,
                                                 const ChunkLists& out, const StreamStage& st, const ChunkLists* prev = nullptr) {
  const int tid = threadIdx.x;
  const int cell0 = chunk * kChunkCells;
  const int cells = (int)min((long long)kChunkCells, G.HW - cell0);
  const long long row0 = b * G.HW + cell0;
  const uint8_t* g_mask = G.mask + row0 * G.S;
  const uint32_t mask_bytes = (uint32_t)cells * (uint32_t)G.S;
  uint8_t* g_zero = reinterpret_cast<uint8_t*>(static_cast<LT*>(G.dlogits) + row0 * G.ld);
  const uint32_t zero_bytes = (uint32_t)cells * (uint32_t)G.ld * (uint32_t)sizeof(LT);
  const RowPrefetch pf{reinterpret_cast<const uint8_t*>(static_cast<const LT*>(G.logits) + row0 * G.ld),
                       (uint32_t)G.Sp * (uint32_t)sizeof(LT), (uint32_t)G.ld * (uint32_t)sizeof(LT)};
  const bool fill = ZERO && G.rows_mode != kRowsClear;
  // Independent global loads first, so that they are all in flight together: the counts of the previous call's rows
  // (the lists they index come next) and this chunk's pick_position mask bytes (one byte per cell: a ballot over them
  // IS a bitmap word).
  int n_prev = 0, np_prev[kMaxPick] = {0, 0, 0, 0};
  if (ZERO && !fill) {
    n_prev = *prev->unit_count;
    for (int kp = 0; kp < G.n_pick; ++kp) np_prev[kp] = prev->pick_count[kp * prev->pick_count_stride];
  }
  constexpr int kCellIters = kChunkCells / BLOCK;
  uint8_t pmv[kMaxPick][kCellIters];
#pragma unroll
  for (int kp = 0; kp < kMaxPick; ++kp)
#pragma unroll
    for (int it = 0; it < kCellIters; ++it) {
      const int c = tid + it * BLOCK;
      pmv[kp][it] = (kp < G.n_pick && c < cells) ? G.pick_mask[(b * G.n_pick + kp) * G.HW + cell0 + c] : (uint8_t)0;
    }
  auto scan_pick_masks = [&]() {
#pragma unroll
    for (int kp = 0; kp < kMaxPick; ++kp)
#pragma unroll
      for (int it = 0; it < kCellIters; ++it) {
        if (kp >= G.n_pick) continue;  // warp-uniform
        const uint32_t bits = __ballot_sync(0xffffffffu, pmv[kp][it] != 0);
        if ((tid & 31) == 0) st.pick_bitmap[kp * (kChunkCells / 32) + ((tid + it * BLOCK) >> 5)] = bits;
      }
  };
  uint32_t skew = 0;
  if (st.image != nullptr) {
    // TMA path: ONE bulk copy lands the chunk's mask bytes in shared memory while bulk copies of a zeroed
    // shared buffer fill the chunk's dlogits; the threads only scan shared memory, one cell each.
    skew = (uint32_t)(reinterpret_cast<uintptr_t>(g_mask) & 15u);
    uint32_t head = skew ? 16u - skew : 0u;
    if (head > mask_bytes) head = mask_bytes;
    const uint32_t body = (mask_bytes - head) & ~15u;
    const uint32_t tail = mask_bytes - head - body;
    if (tid == 0) {
      mbar_init(st.bar, 1);
      fence_async_smem();
      mbar_expect_tx(st.bar, body);
      if (body) bulk_load(st.image + skew + head, g_mask + head, body, st.bar);
    } else if (tid >= 32 && (uint32_t)tid < 32u + head) {
      st.image[skew + tid - 32] = g_mask[tid - 32];
    } else if (tid >= 64 && (uint32_t)tid < 64u + tail) {
      st.image[skew + head + body + tid - 64] = g_mask[head + body + tid - 64];
    }
    if (fill) {
      for (uint32_t i = tid; i < kZeroBuf / 16u; i += BLOCK) reinterpret_cast<uint4*>(st.zeros)[i] = make_uint4(0u, 0u, 0u, 0u);
      fence_async_smem();
    }
    __syncthreads();  // barrier initialised, zero buffer and head / tail bytes in place
    if (fill) zero_fill_bulk(g_zero, zero_bytes, st.zeros);
    if (ZERO && !fill) clear_previous_rows<LT, BLOCK>(G, b, *prev, n_prev, np_prev);  // while the mask image is in flight
    scan_pick_masks();
    mbar_wait(st.bar, 0);
    scan_cells_image<BLOCK>(st.image + skew, cells, (uint32_t)G.S, bitmap, pf);
  } else {
    // the mask bytes are consumed right after the zero fill: start them towards L2 first
    for (uint32_t o = (uint32_t)tid * 128u; o < mask_bytes; o += BLOCK * 128u) prefetch_l2(g_mask + o);
    if (tid < kChunkCells / 32) bitmap[tid] = 0u;
    if (fill) zero_fill<BLOCK>(g_zero, zero_bytes);
    if (ZERO && !fill) clear_previous_rows<LT, BLOCK>(G, b, *prev, n_prev, np_prev);
    scan_pick_masks();
    __syncthreads();
    scan_mask<BLOCK>(g_mask, mask_bytes, (uint32_t)G.S, bitmap, pf);
  }
  __syncthreads();  // (also: every read of the previous lists above is done before `out` may alias them below)
  // one warp per list: the unit cells, then each pick head's valid cells
  for (int job = tid >> 5; job < 1 + G.n_pick; job += BLOCK / 32) {
    if (job == 0)
      compact_cells_warp(bitmap, (cells + 31) >> 5, out.unit, cell0, out.unit_count);
    else
      compact_cells_warp(st.pick_bitmap + (job - 1) * (kChunkCells / 32), (cells + 31) >> 5,
                         out.pick + (job - 1) * out.pick_stride, cell0, out.pick_count + (job - 1) * out.pick_count_stride);
  }
  return skew;
}

template <typename LT, bool ZERO>
__global__ void __launch_bounds__(kStreamBlock) gridnet_stream_kernel(const __grid_constant__ GridDev G) {
  extern __shared__ __align__(16) uint8_t stream_smem[];  // mask image (G.image_bytes; 0: plain path)
  __shared__ __align__(16) uint8_t zeros[kZeroBuf];
  __shared__ uint64_t bar;
  __shared__ uint32_t bitmap[kChunkCells / 32];
  __shared__ uint32_t pick_bitmap[kMaxPick * (kChunkCells / 32)];
  const long long b = blockIdx.x / G.chunks;
  const int chunk = (int)(blockIdx.x - b * G.chunks);
  const ChunkLists out{G.unit_list + (b * G.chunks + chunk) * kChunkCells, G.unit_count + b * G.chunks + chunk,
                       G.pick_list + ((b * G.n_pick) * G.chunks + chunk) * kChunkCells,
                       G.pick_count + (b * G.n_pick) * G.chunks + chunk, (long long)G.chunks * kChunkCells, G.chunks};
  // in the rows modes the list workspace persists across calls: what it holds on entry are the previous call's rows
  stream_chunk<LT, ZERO, kStreamBlock>(G, b, chunk, bitmap, out, StreamStage{G.image_bytes ? stream_smem : nullptr, zeros, &bar, pick_bitmap}, &out);
  pdl_trigger();  // the compute launch may be scheduled: it waits for this grid's completion before reading the lists
  if (ZERO && G.rows_mode != kRowsClear && G.image_bytes && threadIdx.x == 0) bulk_wait_read();  // the zero buffer must outlive the copies that read it
}

// ---- C: compute kernel ---------------------------------------------------------------------------
// Flat position i of a sample's unit list -> cell id, through the per-chunk sub-lists.
struct ListView {
  const uint16_t* list;  // [chunks * kChunkCells]
  const int* prefix;     // shared: exclusive prefix of the chunk counts, [chunks + 1]
  int chunks;
  __device__ __forceinline__ int at(int i) const {
    int c = 0;
    while (c + 1 < chunks && i >= prefix[c + 1]) ++c;
    return (int)list[c * kChunkCells + (i - prefix[c])];
  }
};

template <int MODE, typename LT, int PMAX, bool PICK, bool SELF_STREAM, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (PMAX <= 8 ? 1024 : 512) / BLOCK)
    gridnet_kernel(const __grid_constant__ GridDev G, const __grid_constant__ PpoDev P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_bitmap[kChunkCells / 32];
  __shared__ uint32_t s_pick_bitmap[SELF_STREAM && PICK ? kMaxPick * (kChunkCells / 32) : 1];
  __shared__ double s_wsum[2][BLOCK / 32];        // per-warp (logp, entropy)
  __shared__ Soft s_wpick[kMaxPick][BLOCK / 32];  // per-warp pick statistics
  __shared__ float s_bcast[2];
  __shared__ float s_pick[kMaxPick * 3];  // lse, entropy, any per pick head

  constexpr int NP = PICK ? kMaxPick : 1;  // pick heads compiled in (PICK == false: none)
  const int n_pick = PICK ? G.n_pick : 0;
  const int tid = threadIdx.x;
  const long long b = blockIdx.x;
  const long long row0 = b * G.HW;

  // ---- S. single-chunk samples stream their own chunk here; larger maps ran the streaming launch -------------
  __shared__ uint16_t s_list[SELF_STREAM ? kChunkCells : 1];
  __shared__ uint16_t s_plist[SELF_STREAM && PICK ? kMaxPick * kChunkCells : 1];
  __shared__ int s_counts[1 + kMaxPick];
  // dynamic shared memory: chunk-count prefixes | (log-sum-exp, entropy) stash of later passes | mask image
  int* prefix = reinterpret_cast<int*>(smem);                      // [chunks + 1]
  int* pick_prefix = prefix + (G.chunks + 1);                      // [n_pick][chunks + 1]
  float* s_lse = reinterpret_cast<float*>(smem + (((size_t)(G.chunks + 1) * (1 + kMaxPick) * 4 + 15) & ~(size_t)15));
  float* s_ent = s_lse + (size_t)G.stash * G.A;
  uint8_t* s_image = reinterpret_cast<uint8_t*>(s_lse) + (((size_t)G.stash * G.A * 2 * sizeof(float) + 15) & ~(size_t)15);

  const LT* g_logits = static_cast<const LT*>(G.logits) + row0 * G.ld;
  LT* g_out = static_cast<LT*>(G.dlogits) + row0 * G.ld;
  const uint8_t* g_mask = G.mask + row0 * G.S;
  __shared__ __align__(16) uint8_t s_zeros[SELF_STREAM && MODE != kFwd ? kZeroBuf : 16];
  __shared__ uint64_t s_bar;
  if constexpr (SELF_STREAM) {  // the lists never leave the SM
    const ChunkLists out{s_list, s_counts, s_plist, s_counts + 1, kChunkCells, 1};
    // the rows the previous call wrote into this (persistent) dlogits buffer, recorded in the list workspace
    const ChunkLists prev{G.unit_list + b * kChunkCells, G.unit_count + b, G.pick_list + (b * G.n_pick) * kChunkCells,
                          G.pick_count + b * G.n_pick, kChunkCells, 1};
    const uint32_t skew = stream_chunk<LT, MODE != kFwd, BLOCK>(G, b, 0, s_bitmap, out,
                                                               StreamStage{G.image_bytes ? s_image : nullptr, s_zeros, &s_bar, s_pick_bitmap}, &prev);
    if (G.image_bytes) g_mask = s_image + skew;  // the unit cells re-read their mask bytes from the image
    __syncthreads();
    if (MODE == kPpo && G.rows_mode != kRowsOff) {  // record this call's rows for the next one (the previous were read above)
      const int n = s_counts[0];
      for (int i = tid; i < n; i += BLOCK) prev.unit[i] = s_list[i];
      if (tid == 0) *prev.unit_count = n;
      for (int kp = 0; kp < n_pick; ++kp) {
        const int np = s_counts[1 + kp];
        for (int i = tid; i < np; i += BLOCK) prev.pick[kp * kChunkCells + i] = s_plist[kp * kChunkCells + i];
        if (tid == 0) prev.pick_count[kp] = np;
      }
    }
  }

  // the per-sample PPO scalars are consumed by one thread after the reductions: start pulling them now; the
  // advantage normaliser (mean, std + 1e-8) is derived from the float64 moments here, off the critical path
  __shared__ float s_norm[2 * B200RL_MAX_VALUE_HEADS];
  __shared__ float s_ahead[2];           // normalised advantage and behaviour log-prob of this sample
  __shared__ float s_pick_xa[kMaxPick];  // the chosen pick cell's logit (finfo.min when that cell is masked)
  if (MODE == kPpo) {
    const int Vm = P.adv_mode == 3 ? 1 : P.adv_v;
    if (P.adv_mode != 0 && tid >= 64 && tid < 64 + Vm) ppo_norm_pair(P, Vm, tid - 64, s_norm);
    if (tid >= 32 && tid < 32 + P.V) {
      const long long o = b * P.V + (tid - 32);
      prefetch_l1(P.new_values + o), prefetch_l1(P.old_values + o), prefetch_l1(P.returns + o);
    }
  }
  // Three dependent global loads (pick action -> its mask byte -> its logit) that do not depend on the lists: done
  // here, by a thread of the last warp, instead of on thread 0's path between the reductions and the PPO terms
  if (PICK && MODE != kBwd && tid >= BLOCK - kMaxPick && tid - (BLOCK - kMaxPick) < n_pick) {
    const int kp = tid - (BLOCK - kMaxPick);
    const long long a = (long long)load_index(G.pick_actions, G.pick_dtype, b * n_pick + kp);
    float xa = 0.f;
    if (a >= 0 && a < G.HW)
      xa = G.pick_mask[(b * n_pick + kp) * G.HW + a]
               ? to_f32((static_cast<const LT*>(G.logits) + row0 * G.ld)[a * G.ld + G.S + kp]) : kF32Lowest;
    s_pick_xa[kp] = xa;
  }

  // ---- 1. this sample's lists ------------------------------------------------------------------------
  // the lane layout: a thread's slot depends on its lane, and a divergent read of the kernel parameters (constant
  // bank) serialises, so one warp copies the table to shared memory (13 % of the stall samples before)
  __shared__ LaneSlot s_slot[32];
  if (tid >= BLOCK - 32) s_slot[tid - (BLOCK - 32)] = G.slot[tid - (BLOCK - 32)];
  if (!SELF_STREAM) pdl_wait();  // the streaming launch ahead of this one wrote them (programmatic dependent launch)
  // exclusive prefixes of the per-chunk counts: one warp per list, counts loaded side by side (a serial walk of 16
  // dependent global loads by one thread cost ~10 us at 64x64) and scanned with shuffles
  for (int job = tid >> 5; job < 1 + n_pick; job += BLOCK / 32) {
    const int lane = tid & 31;
    int* pp = job == 0 ? prefix : pick_prefix + (job - 1) * (G.chunks + 1);
    int acc = 0;
    for (int c0 = 0; c0 < G.chunks; c0 += 32) {
      const int c = c0 + lane;
      int v = 0;
      if (c < G.chunks)
        v = SELF_STREAM ? s_counts[job]
                        : (job == 0 ? G.unit_count[b * G.chunks + c] : G.pick_count[(b * n_pick + job - 1) * G.chunks + c]);
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (c < G.chunks) pp[c] = acc + incl - v;
      acc += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) pp[G.chunks] = acc;
  }
  __syncthreads();
  if (MODE == kPpo && tid == BLOCK - 33) {  // s_norm is in place: the sample's advantage and behaviour log-prob, ahead of time
    s_ahead[0] = ppo_sample_advantage(P, b, s_norm);
    s_ahead[1] = P.old_logp[b];
  }
  const ListView units{SELF_STREAM ? s_list : G.unit_list + (b * G.chunks) * kChunkCells, prefix, G.chunks};
  const int n_unit = prefix[G.chunks];

  // the pick_position logit of this thread's first valid cell, per pick head: the loads are issued here, unconsumed, so
  // that they travel together with the forward pass's loads instead of after its reductions
  LT pick_raw[NP];
#pragma unroll
  for (int kp = 0; kp < NP; ++kp) {
    pick_raw[kp] = from_f32<LT>(0.f);
    if (kp < n_pick) {
      const ListView pl{SELF_STREAM ? s_plist + kp * kChunkCells
                                    : G.pick_list + ((b * n_pick + kp) * G.chunks) * kChunkCells,
                        pick_prefix + kp * (G.chunks + 1), G.chunks};
      if (tid < pl.prefix[G.chunks]) pick_raw[kp] = g_logits[(long long)pl.at(tid) * G.ld + G.S + kp];
    }
  }

  // ---- 2. forward over the unit cells -----------------------------------------------------------------
  const int group = tid / G.G, n_groups = BLOCK / G.G;
  const LaneSlot slot = s_slot[tid & (G.G - 1)];
  const int gate_ref = slot.len ? G.gate_ref[slot.head] : -1;
  const int gate_val = slot.len ? G.gate_val[slot.head] : 0;
  double logp_acc = 0.0, ent_acc = 0.0;
  // state of the first pass stays in registers for the backward (a sample rarely has more unit
  // cells than lane groups); later passes park (log-sum-exp, entropy) in shared memory, and cells
  // beyond the stash recompute their forward in the backward pass.
  float k_d[PMAX], k_e[PMAX];
  uint32_t k_valid = 0u;
  float k_ls = 0.f, k_inv = 0.f, k_ent = 0.f;
  int k_cell = 0, k_local = -1;
  bool k_gated = false;
#pragma unroll 1
  for (int i0 = 0; i0 < n_unit; i0 += n_groups) {  // uniform trip count: shuffles stay converged
    const int i = i0 + group;
    const bool live = i < n_unit && slot.len > 0;
    const int cell = live ? units.at(i) : 0;
    Piece<LT, PMAX> p;
    p.valid = 0u;
#pragma unroll
    for (int j = 0; j < PMAX; ++j) p.x[j] = 0.f;
    int a_head = 0, a_ref = gate_val;
    if (live) {
      p = load_piece<LT, PMAX>(g_logits + (long long)cell * G.ld, g_mask + (long long)cell * G.S, slot);
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
    } else if (slot.first && i < G.stash) {
      s_lse[i * G.A + slot.head] = h.any ? h.mx + h.ls : INFINITY;  // +inf marks a head with no valid entry
      s_ent[i * G.A + slot.head] = h.ent;
    }
  }

  // ---- 3. pick_position: one online-softmax pass over the valid cells of the sample ----------------------
  Soft pick[NP];
#pragma unroll
  for (int kp = 0; kp < NP; ++kp) {
    pick[kp] = Soft{-INFINITY, 0.f, 0.f};
    if (kp >= n_pick) continue;
    const ListView pl{SELF_STREAM ? s_plist + kp * kChunkCells
                                  : G.pick_list + ((b * n_pick + kp) * G.chunks) * kChunkCells,
                      pick_prefix + kp * (G.chunks + 1), G.chunks};
    const int n_valid = pl.prefix[G.chunks];
    if (tid < n_valid) pick[kp] = Soft{to_f32(pick_raw[kp]), 1.f, 0.f};
#pragma unroll 1
    for (int i = tid + BLOCK; i < n_valid; i += BLOCK)
      pick[kp] = soft_push(pick[kp], to_f32(g_logits[(long long)pl.at(i) * G.ld + G.S + kp]));
  }

  // ---- 4. per-sample totals: warp partials -> thread 0 ---------------------------------------------------
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
  // warp 0 folds the per-warp partials with shuffles (fixed order), thread 0 finishes the sample
  double tot_logp = 0.0, tot_ent = 0.0;
  Soft tot_pick[NP];
  if (tid < 32) {
    constexpr int NW = BLOCK / 32;
    tot_logp = tid < NW ? s_wsum[0][tid] : 0.0, tot_ent = tid < NW ? s_wsum[1][tid] : 0.0;
#pragma unroll
    for (int kp = 0; kp < NP; ++kp) tot_pick[kp] = (kp < n_pick && tid < NW) ? s_wpick[kp][tid] : Soft{-INFINITY, 0.f, 0.f};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (o < NW) {
        tot_logp += shfl_xor_f64(tot_logp, o);
        tot_ent += shfl_xor_f64(tot_ent, o);
#pragma unroll
        for (int kp = 0; kp < NP; ++kp)
          if (kp < n_pick) {
            Soft other;
            other.m = __shfl_xor_sync(0xffffffffu, tot_pick[kp].m, o);
            other.s = __shfl_xor_sync(0xffffffffu, tot_pick[kp].s, o);
            other.q = __shfl_xor_sync(0xffffffffu, tot_pick[kp].q, o);
            tot_pick[kp] = soft_merge(tot_pick[kp], other);
          }
      }
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int kp = 0; kp < NP; ++kp) {
      if (kp >= n_pick) continue;
      const Soft t = tot_pick[kp];
      const bool any = t.s > 0.f;
      float p_lse = 0.f, p_ent = 0.f;
      if (any) {
        const float ls = logf(t.s);
        p_lse = t.m + ls;
        p_ent = ls - t.q / t.s;
        tot_logp += (double)(s_pick_xa[kp] - p_lse);
        tot_ent += (double)p_ent;
      }
      s_pick[kp * 3] = p_lse, s_pick[kp * 3 + 1] = p_ent, s_pick[kp * 3 + 2] = any ? 1.f : 0.f;
    }
    if (MODE == kFwd) {
      G.logp[b] = (float)tot_logp, G.entropy[b] = (float)tot_ent;
    } else if (MODE == kBwd) {
      s_bcast[0] = G.dlogp_in[b], s_bcast[1] = G.dent_in[b];
    } else {
      // ---- 5. PPO scalar stage ------------------------------------------------------------------------------
      PolicyTerms t = ppo_policy_terms_at(P, b, tot_logp, s_ahead[0], s_ahead[1]);
      s_bcast[0] = t.dlogp, s_bcast[1] = ppo_dentropy(P, 1);
      double* row = P.partials + b * ppo_nstat(P.V);
      row[0] = t.surrogate, row[1] = tot_ent, row[2] = t.kl, row[3] = t.clipped, row[4] = t.teacher;
      if (G.logp) G.logp[b] = (float)tot_logp;
      if (G.entropy) G.entropy[b] = (float)tot_ent;
    }
  }
  if (MODE == kPpo && tid >= 32 && tid < 32 + P.V) {
    const int v = tid - 32;
    float2 r = ppo_value_terms(P, b, v);
    double* row = P.partials + b * ppo_nstat(P.V);
    row[kPolicyStats + v] = r.x, row[kPolicyStats + P.V + v] = r.y;
  }
  if (MODE == kFwd) return;
  if (MODE == kPpo) pdl_trigger();  // the stats finaliser may be scheduled; it waits for this grid to complete
  if (SELF_STREAM && G.image_bytes && G.rows_mode != kRowsClear && tid == 0) {  // the zero fill of this sample's rows must have landed
    bulk_wait_all();
    fence_async_all();
  }
  __syncthreads();
  dlogp = s_bcast[0], dent = s_bcast[1];

  // ---- 6. backward over the unit cells: overwrite their zero-filled rows -----------------------------------
  if (k_valid)
    piece_backward<LT, PMAX>(g_out + (long long)k_cell * G.ld + slot.off, slot, k_valid, k_d, k_e, k_ls, k_inv, k_ent,
                             k_local, k_gated ? dlogp : 0.f, dent);
#pragma unroll 1
  for (int i0 = n_groups; i0 < n_unit; i0 += n_groups) {  // cells beyond the first pass (uniform trip count)
    const int i = i0 + group;
    const bool live = i < n_unit && slot.len > 0;
    const int cell = live ? units.at(i) : 0;
    Piece<LT, PMAX> p;
    p.valid = 0u;
#pragma unroll
    for (int j = 0; j < PMAX; ++j) p.x[j] = 0.f;
    if (live) p = load_piece<LT, PMAX>(g_logits + (long long)cell * G.ld, g_mask + (long long)cell * G.S, slot);
    float lse, ent;
    if (i0 + n_groups - 1 < G.stash) {  // the whole pass is in the stash (warp-uniform test)
      lse = live ? s_lse[i * G.A + slot.head] : INFINITY;
      ent = live ? s_ent[i * G.A + slot.head] : 0.f;
    } else {  // beyond the stash: redo the forward reductions
      float d0[PMAX], e0[PMAX];
      const HeadStat h = head_forward<LT, PMAX>(p, slot, G.max_width, d0, e0);
      lse = h.any ? h.mx + h.ls : INFINITY, ent = h.ent;
    }
    if (!live || lse == INFINITY) continue;  // no valid entry in this head: gradient stays zero
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
    piece_backward<LT, PMAX>(g_out + (long long)cell * G.ld + slot.off, slot, p.valid, d, e, 0.f, 1.f, ent, local,
                             gated_in ? dlogp : 0.f, dent);
  }
#pragma unroll
  for (int kp = 0; kp < NP; ++kp) {
    if (kp >= n_pick || s_pick[kp * 3 + 2] == 0.f) continue;
    const float p_lse = s_pick[kp * 3], p_ent = s_pick[kp * 3 + 1];
    const ListView pl{SELF_STREAM ? s_plist + kp * kChunkCells
                                  : G.pick_list + ((b * n_pick + kp) * G.chunks) * kChunkCells,
                      pick_prefix + kp * (G.chunks + 1), G.chunks};
    const int n_valid = pl.prefix[G.chunks];
    const long long a = (long long)load_index(G.pick_actions, G.pick_dtype, b * n_pick + kp);
#pragma unroll 1
    for (int i = tid; i < n_valid; i += BLOCK) {
      const int c = pl.at(i);
      const float lp = to_f32(g_logits[(long long)c * G.ld + G.S + kp]) - p_lse;
      const float pr = fast_exp(lp);
      g_out[(long long)c * G.ld + G.S + kp] = from_f32<LT>(dlogp * ((a == c ? 1.f : 0.f) - pr) - dent * pr * (lp + p_ent));
    }
  }
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

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// workspace of the two-launch scheme: unit lists + counts (+ pick lists + counts)
static size_t grid_workspace_bytes(long long B, long long HW, int n_pick) {
  const long long chunks = (HW + kChunkCells - 1) / kChunkCells;
  size_t n = align16((size_t)B * chunks * kChunkCells * sizeof(uint16_t)) + align16((size_t)B * chunks * sizeof(int));
  n += align16((size_t)B * n_pick * chunks * kChunkCells * sizeof(uint16_t)) + align16((size_t)B * n_pick * chunks * sizeof(int));
  return n + 64;
}

static int bind_workspace(GridDev* G, void* workspace, size_t workspace_bytes, const char* who) {
  G->chunks = (int)((G->HW + kChunkCells - 1) / kChunkCells);
  G->stash = (int)(G->HW < kStashCells ? G->HW : kStashCells);
  if (G->chunks == 1 && G->stash > kStashCellsSelf) G->stash = kStashCellsSelf;  // self-streaming CTAs: shared memory buys CTAs / SM
  const size_t image = align16((size_t)(G->HW < kChunkCells ? G->HW : kChunkCells) * G->S) + 32;
  G->image_bytes = image <= kMaxImageBytes ? (int)image : 0;
  B200RL_REQUIRE(workspace != nullptr && workspace_bytes >= grid_workspace_bytes(G->B, G->HW, G->n_pick),
                 "%s: workspace too small (%zu < %zu bytes)", who, workspace_bytes,
                 grid_workspace_bytes(G->B, G->HW, G->n_pick));
  uint8_t* w = static_cast<uint8_t*>(workspace);
  w += (16 - (reinterpret_cast<uintptr_t>(w) & 15)) & 15;
  G->unit_list = reinterpret_cast<uint16_t*>(w), w += align16((size_t)G->B * G->chunks * kChunkCells * sizeof(uint16_t));
  G->unit_count = reinterpret_cast<int*>(w), w += align16((size_t)G->B * G->chunks * sizeof(int));
  G->pick_list = reinterpret_cast<uint16_t*>(w), w += align16((size_t)G->B * G->n_pick * G->chunks * kChunkCells * sizeof(uint16_t));
  G->pick_count = reinterpret_cast<int*>(w);
  return B200RL_OK;
}

static size_t compute_smem(const GridDev& G, bool self_stream) {
  return align16((size_t)(G.chunks + 1) * (1 + kMaxPick) * 4) + align16((size_t)G.stash * G.A * 2 * sizeof(float)) +
         (self_stream ? (size_t)G.image_bytes : 0);
}

template <int MODE, typename LT, int PMAX, bool PICK, bool SELF_STREAM, int BLOCK>
static int launch_compute_block(GridDev& G, const PpoDev& P, cudaStream_t stream) {
  auto kernel = gridnet_kernel<MODE, LT, PMAX, PICK, SELF_STREAM, BLOCK>;
  const size_t smem = compute_smem(G, SELF_STREAM);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("gridnet: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return B200RL_ECUDA;
    }
  }
  if (SELF_STREAM) {
    kernel<<<(unsigned)G.B, BLOCK, smem, stream>>>(G, P);
  } else {  // dependent of the streaming launch: prologue (normaliser, prefetches) overlaps its tail
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)G.B), cfg.blockDim = dim3(BLOCK), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, G, P);
  }
  return check_launch("gridnet");
}

template <int MODE, typename LT, int PMAX, bool PICK, bool SELF_STREAM>
static int launch_compute(GridDev& G, const PpoDev& P, cudaStream_t stream) {
  // a self-streaming CTA is mostly a streamer: 128 threads keep 8 of them per SM; the compute-only
  // launch of the split path gets 256 threads (more lane groups per sample) -- and 1024 when the minibatch has fewer
  // samples than the GPU has SMs (one C5 per-GPU minibatch: B = 128 on 148 SMs): one CTA per SM either way, so
  // the sample's unit cells are taken in ONE pass of 128 lane groups instead of 3-4 passes of 32
  if (SELF_STREAM) return launch_compute_block<MODE, LT, PMAX, PICK, true, 128>(G, P, stream);
  if (MODE == kPpo && PMAX <= 8 && G.B <= device_info().sm_count)
    return launch_compute_block<MODE, LT, PMAX, PICK, false, (MODE == kPpo && PMAX <= 8) ? 1024 : 256>(G, P, stream);
  return launch_compute_block<MODE, LT, PMAX, PICK, false, 256>(G, P, stream);
}

template <int MODE, typename LT, int PMAX, bool PICK>
static int launch_pair(GridDev& G, const PpoDev& P, cudaStream_t stream) {
  const long long stream_ctas = G.B * G.chunks;
  if (stream_ctas > 0x7fffffffLL) {
    set_error("gridnet: batch too large (%lld samples x %d chunks)", G.B, G.chunks);
    return B200RL_EUNSUPPORTED;
  }
  if (G.chunks == 1) return launch_compute<MODE, LT, PMAX, PICK, true>(G, P, stream);
  if (MODE == kFwd) gridnet_stream_kernel<LT, false><<<(unsigned)stream_ctas, kStreamBlock, G.image_bytes, stream>>>(G);
  else gridnet_stream_kernel<LT, true><<<(unsigned)stream_ctas, kStreamBlock, G.image_bytes, stream>>>(G);
  int rc = check_launch("gridnet_stream");
  if (rc) return rc;
  return launch_compute<MODE, LT, PMAX, PICK, false>(G, P, stream);
}

template <int MODE>
static int launch_mode(GridDev& G, const PpoDev& P, const int* nvec, int logits_dtype, cudaStream_t stream) {
  B200RL_UNSUPPORTED(G.chunks * kChunkCells > 65536, "gridnet: HW=%lld cells exceeds 65536", G.HW);
  const bool bf16 = logits_dtype == B200RL_BF16;
  if (plan_lanes(&G, nvec, 8)) {
    if (G.n_pick == 0)
      return bf16 ? launch_pair<MODE, __nv_bfloat16, 8, false>(G, P, stream)
                  : launch_pair<MODE, float, 8, false>(G, P, stream);
    return bf16 ? launch_pair<MODE, __nv_bfloat16, 8, true>(G, P, stream) : launch_pair<MODE, float, 8, true>(G, P, stream);
  }
  if (plan_lanes(&G, nvec, 32)) {
    return bf16 ? launch_pair<MODE, __nv_bfloat16, 32, true>(G, P, stream)
                : launch_pair<MODE, float, 32, true>(G, P, stream);
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
  B200RL_REQUIRE(d->logits_ld == 0 || d->logits_ld >= G.Sp, "%s: logits_ld=%lld is narrower than a row (%d)", who,
                 (long long)d->logits_ld, G.Sp);
  B200RL_UNSUPPORTED(d->logits_ld > 65535, "%s: logits_ld=%lld exceeds 65535", who, (long long)d->logits_ld);
  G.ld = d->logits_ld ? (int)d->logits_ld : G.Sp;
  G.rows_mode = kRowsOff;
  *out = G;
  return B200RL_OK;
}

}  // namespace b200rl

extern "C" size_t b200rl_gridnet_workspace_bytes(int64_t B, int64_t HW, int n_pick) {
  return b200rl::grid_workspace_bytes(B < 1 ? 1 : B, HW < 1 ? 1 : HW, n_pick < 0 ? 0 : n_pick);
}

extern "C" size_t b200rl_ppo_gridnet_workspace_bytes(int64_t B, int64_t HW, int n_pick, int64_t V) {
  return b200rl_ppo_workspace_bytes(B, V) + b200rl_gridnet_workspace_bytes(B, HW, n_pick);
}

extern "C" int b200rl_gridnet_fwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                  const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                  float* logp, float* entropy, void* workspace, size_t workspace_bytes,
                                  b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, "gridnet_fwd");
  if (rc) return rc;
  B200RL_REQUIRE(logp && entropy, "gridnet_fwd: null output");
  if (G.B == 0) return B200RL_OK;
  rc = bind_workspace(&G, workspace, workspace_bytes, "gridnet_fwd");
  if (rc) return rc;
  G.logp = logp, G.entropy = entropy;
  PpoDev P{};
  return launch_mode<kFwd>(G, P, d->nvec_host, d->logits_dtype, (cudaStream_t)stream);
}

extern "C" int b200rl_gridnet_bwd(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                  const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                  const float* dlogp, const float* dentropy, void* dlogits, void* workspace,
                                  size_t workspace_bytes, b200rl_stream_t stream) {
  using namespace b200rl;
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, "gridnet_bwd");
  if (rc) return rc;
  B200RL_REQUIRE(dlogp && dentropy && dlogits, "gridnet_bwd: null pointer");
  if (G.B == 0) return B200RL_OK;
  rc = bind_workspace(&G, workspace, workspace_bytes, "gridnet_bwd");
  if (rc) return rc;
  G.dlogp_in = dlogp, G.dent_in = dentropy, G.dlogits = dlogits;
  PpoDev P{};
  return launch_mode<kBwd>(G, P, d->nvec_host, d->logits_dtype, (cudaStream_t)stream);
}

namespace b200rl {
// rows == nullptr: the lists live in the call's workspace and dlogits is zero-filled; otherwise they live in the
// caller's persistent `rows` buffer and say which rows of the (persistent) dlogits the previous call wrote.
static int ppo_gridnet_loss_impl(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                 const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                 const b200rl_ppo_args* args, void* dlogits, float* logp_out, float* entropy_out,
                                 void* workspace, size_t workspace_bytes, void* rows, size_t rows_bytes, int rows_valid,
                                 cudaStream_t s, const char* who) {
  GridDev G;
  int rc = make_grid(d, logits, mask, pick_mask, actions, pick_actions, &G, who);
  if (rc) return rc;
  B200RL_REQUIRE(dlogits != nullptr, "%s: dlogits is null", who);
  // workspace = [PPO partials | GridNet lists]; b200rl_ppo_gridnet_workspace_bytes() sizes both
  const size_t ppo_bytes = b200rl_ppo_workspace_bytes(G.B, args ? args->V : 1);
  B200RL_REQUIRE(workspace && workspace_bytes >= ppo_bytes, "%s: workspace too small", who);
  PpoDev P;
  rc = ppo_make_dev(args, G.B, workspace, ppo_bytes, &P);
  if (rc) return rc;
  if (rows != nullptr) {
    rc = bind_workspace(&G, rows, rows_bytes, who);
    G.rows_mode = rows_valid ? kRowsClear : kRowsRecord;
  } else {
    rc = bind_workspace(&G, static_cast<uint8_t*>(workspace) + ppo_bytes, workspace_bytes - ppo_bytes, who);
  }
  if (rc) return rc;
  G.dlogits = dlogits, G.logp = logp_out, G.entropy = entropy_out;
  rc = launch_mode<kPpo>(G, P, d->nvec_host, d->logits_dtype, s);  // derives the advantage normaliser itself
  if (rc) return rc;
  return ppo_launch_finalize(P, G.B, 1, s);
}
}  // namespace b200rl

extern "C" int b200rl_ppo_gridnet_loss(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                       const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                       const b200rl_ppo_args* args, void* dlogits, float* logp_out,
                                       float* entropy_out, void* workspace, size_t workspace_bytes,
                                       b200rl_stream_t stream) {
  return b200rl::ppo_gridnet_loss_impl(d, logits, mask, pick_mask, actions, pick_actions, args, dlogits, logp_out,
                                       entropy_out, workspace, workspace_bytes, nullptr, 0, 0, (cudaStream_t)stream,
                                       "ppo_gridnet_loss");
}

extern "C" size_t b200rl_gridnet_rows_bytes(int64_t B, int64_t HW, int n_pick) {
  return b200rl_gridnet_workspace_bytes(B, HW, n_pick);
}

extern "C" int b200rl_ppo_gridnet_loss_inplace(const b200rl_gridnet_desc* d, const void* logits, const uint8_t* mask,
                                               const uint8_t* pick_mask, const void* actions, const void* pick_actions,
                                               const b200rl_ppo_args* args, void* dlogits, float* logp_out,
                                               float* entropy_out, void* workspace, size_t workspace_bytes, void* rows,
                                               size_t rows_bytes, int rows_valid, b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(rows != nullptr, "ppo_gridnet_loss_inplace: rows is null");
  return ppo_gridnet_loss_impl(d, logits, mask, pick_mask, actions, pick_actions, args, dlogits, logp_out, entropy_out,
                               workspace, workspace_bytes, rows, rows_bytes, rows_valid, (cudaStream_t)stream,
                               "ppo_gridnet_loss_inplace");
}

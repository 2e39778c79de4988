// Mask-scan building blocks shared by the fused GridNet loss (gridnet.cu) and the rollout-time
// sampler (gridnet_sample.cu): 128-bit zero fill, 128-bit scan of the action-mask bytes into a
// shared bitmap of non-empty cells, and the bitmap -> ascending cell list compaction.
#pragma once
#include "bulk.cuh"
#include "common.cuh"

namespace b200rl {

constexpr int kChunkCells = 256;  // cells per scan chunk (one bitmap of 8 words)

// ---- zero fill / mask scan -----------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ void zero_fill(uint8_t* dst, uint32_t bytes) {
  const uint32_t tid = threadIdx.x;
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  if (tid < head) dst[tid] = 0;
  uint4* d4 = reinterpret_cast<uint4*>(dst + head);
  const uint32_t n4 = (bytes - head) >> 4;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint32_t i = tid;
  for (; i + 3u * BLOCK < n4; i += 4u * BLOCK) {
    d4[i] = z, d4[i + BLOCK] = z, d4[i + 2u * BLOCK] = z, d4[i + 3u * BLOCK] = z;
  }
  for (; i < n4; i += BLOCK) d4[i] = z;
  const uint32_t done = head + (n4 << 4);
  if (tid < bytes - done) dst[done + tid] = 0;
}

// Same fill through the TMA: the unaligned head / tail bytes with plain stores, the 16-byte aligned body as
// bulk copies of a zeroed shared buffer (`zeros`, kZeroBuf bytes, already fenced towards the async proxy
// and followed by a CTA barrier).  Thread 0 issues and commits; the caller waits (bulk_wait_*).
constexpr uint32_t kZeroBuf = 2048;
__device__ __forceinline__ void zero_fill_bulk(uint8_t* dst, uint32_t bytes, const uint8_t* zeros) {
  const uint32_t tid = threadIdx.x;
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  const uint32_t body = (bytes - head) & ~15u;
  const uint32_t tail = bytes - head - body;
  if (tid == 0) {
    for (uint32_t o = 0; o < body; o += kZeroBuf) bulk_store(dst + head + o, zeros, min(kZeroBuf, body - o));
    bulk_commit();
  } else if (tid >= 32u && tid < 32u + head) {
    dst[tid - 32u] = 0;
  } else if (tid >= 64u && tid < 64u + tail) {
    dst[head + body + tid - 64u] = 0;
  }
}

// Row prefetch context: the logits row of a flagged cell is wanted a few microseconds from now.
struct RowPrefetch {
  const uint8_t* logits;  // first logit of the CTA's first cell (NULL: no prefetch)
  uint32_t row_bytes;     // Sp * sizeof(LT): the bytes of a row that will be read
  uint32_t stride;        // ld * sizeof(LT): bytes between consecutive cells' rows
};

// byte offset -> cell without an integer division: floor(n / S) == umulhi(n, ceil(2^32 / S)) for
// n * S < 2^32 (checked by the caller); inv == 0 selects the plain division (S == 1 or huge chunks).
__device__ __forceinline__ uint32_t cell_of(uint32_t n, uint32_t S, uint32_t inv) {
  return inv ? __umulhi(n, inv) : n / S;
}

__device__ __forceinline__ void flag_cell(uint32_t* bitmap, uint32_t cell, const RowPrefetch& pf) {
  const uint32_t bit = 1u << (cell & 31);
  const uint32_t old = atomicOr(&bitmap[cell >> 5], bit);
  if (!(old & bit) && pf.logits != nullptr) {  // first flag of this cell: pull its logits row towards L2
    const uint8_t* row = pf.logits + (size_t)cell * pf.stride;
    for (uint32_t o = 0; o < pf.row_bytes + 127u; o += 128u) prefetch_l2(row + min(o, pf.row_bytes - 1u));
  }
}

// A 16-byte word of the mask chunk with at least one non-zero byte: flag the cell(s) it covers.
// Rows are S bytes, so for S >= 16 a word touches at most two cells; two 32-bit divisions and a
// byte-boundary split decide which.  (Chunk sizes are < 2^31 bytes: cells <= 4096, S <= 65535.)
static __device__ __noinline__ void flag_word(uint32_t* bitmap, uint32_t base, const uint4& w, uint32_t S, uint32_t inv,
                                              const RowPrefetch& pf) {
  const uint32_t c0 = cell_of(base, S, inv), c1 = cell_of(base + 15u, S, inv);
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
      if ((word[q] >> (8 * r)) & 0xffu) flag_cell(bitmap, cell_of(base + 4u * q + r, S, inv), pf);
}

// flags every cell of [mask, mask + bytes) that has a non-zero byte
template <int BLOCK>
__device__ __forceinline__ void scan_mask(const uint8_t* mask, uint32_t bytes, uint32_t S, uint32_t* bitmap,
                                          const RowPrefetch& pf) {
  const uint32_t tid = threadIdx.x;
  const uint32_t inv = (S > 1u && (unsigned long long)(bytes + 16u) * S < (1ull << 32)) ? 0xFFFFFFFFu / S + 1u : 0u;
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(mask) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  if (tid < head && mask[tid]) flag_cell(bitmap, tid / S, pf);
  const uint4* m4 = reinterpret_cast<const uint4*>(mask + head);
  const uint32_t n4 = (bytes - head) >> 4;
  constexpr uint32_t kUnroll = 4;
  for (uint32_t i0 = tid; i0 < n4; i0 += BLOCK * kUnroll) {
    uint4 w[kUnroll];
#pragma unroll
    for (uint32_t u = 0; u < kUnroll; ++u) {
      const uint32_t i = i0 + u * BLOCK;
      // plain read-only loads (allocate in L1): the unit cells re-read their own mask bytes
      w[u] = i < n4 ? __ldg(m4 + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (uint32_t u = 0; u < kUnroll; ++u)
      if ((w[u].x | w[u].y | w[u].z | w[u].w) != 0u) flag_word(bitmap, head + ((i0 + u * BLOCK) << 4), w[u], S, inv, pf);
  }
  const uint32_t done = head + (n4 << 4);
  if (tid < bytes - done && mask[done + tid]) flag_cell(bitmap, (done + tid) / S, pf);
}

// The scan over a shared-memory image of the chunk's mask (landed there by a bulk copy): thread t ORs the S
// bytes of cells t, t + BLOCK, ... with aligned 32-bit loads (edge words masked), and one ballot per warp IS a
// word of the bitmap -- no atomics, no per-word cell arithmetic.  `m` = byte 0 of the chunk's first cell.
template <int BLOCK>
__device__ __forceinline__ void scan_cells_image(const uint8_t* m, int cells, uint32_t S, uint32_t* bitmap,
                                                 const RowPrefetch& pf) {
  static_assert(kChunkCells % BLOCK == 0 && BLOCK % 32 == 0, "one ballot per bitmap word");
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll 1
  for (int c = threadIdx.x; c < kChunkCells; c += BLOCK) {  // warp-uniform trip count
    uint32_t any = 0u;
    if (c < cells) {
      const uintptr_t lo = reinterpret_cast<uintptr_t>(m) + (uintptr_t)c * S, hi = lo + S;  // [lo, hi)
      const uint32_t* w = reinterpret_cast<const uint32_t*>(lo & ~(uintptr_t)3);
      const uint32_t* w_end = reinterpret_cast<const uint32_t*>((hi + 3) & ~(uintptr_t)3);
      const uint32_t lead = (uint32_t)(lo & 3u), trail = (uint32_t)((0u - (uint32_t)hi) & 3u);  // bytes outside the row
      const int n = (int)(w_end - w);
      uint32_t first = w[0] & (0xFFFFFFFFu << (8u * lead));  // little-endian: low bytes come first
      if (n == 1) {
        any = first & (0xFFFFFFFFu >> (8u * trail));
      } else {
        any = first | (w[n - 1] & (0xFFFFFFFFu >> (8u * trail)));
#pragma unroll 4
        for (int k = 1; k < n - 1; ++k) any |= w[k];
      }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, any != 0u);
    if (lane == 0) bitmap[c >> 5] = bits;
    if (any != 0u && pf.logits != nullptr) {  // pull the unit cell's logits row towards L2
      const uint8_t* row = pf.logits + (size_t)c * pf.stride;
      for (uint32_t o = 0; o < pf.row_bytes + 127u; o += 128u) prefetch_l2(row + min(o, pf.row_bytes - 1u));
    }
  }
}

// ONE warp (every lane of it calls): bitmap -> ascending list of flagged cells (ids offset by `base_id`); count
// through *out_count
__device__ __forceinline__ void compact_cells_warp(const uint32_t* bitmap, int words, uint16_t* list, int base_id,
                                                   int* out_count) {
  const int lane = threadIdx.x & 31;
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
      list[pos++] = (uint16_t)(base_id + w * 32 + b);
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) *out_count = base;
}

// warp 0 of the CTA does the compaction
__device__ __forceinline__ void compact_cells(const uint32_t* bitmap, int words, uint16_t* list, int base_id,
                                              int* out_count) {
  if (threadIdx.x >= 32) return;
  compact_cells_warp(bitmap, words, list, base_id, out_count);
}

}  // namespace b200rl

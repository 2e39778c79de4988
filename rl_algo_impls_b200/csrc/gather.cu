// K3: minibatch row gather -- dst[t][b, :] = src[t][idx[b], :] for every Batch field in one call.
// Replaces the 8-12 separate fancy-index kernels of rollout/rollout.py:56-69
// (Batch.__getitem__) / shared/tensor_utils.py:66-72 and the .to(device) of
// rollout/vec_rollout.py:175 (the rollout is already device resident here).
//
// Wide rows (observations, masks, per-cell actions: KBs to MBs per row) are cut into 16 KB
// chunks, one CTA per chunk, moved with 128-bit streaming loads/stores -- every load of a
// chunk is issued before its first store.  Narrow rows (log-probs, values, advantages,
// returns: 4..128 bytes) are handled one row per thread in the same launch family so a
// minibatch costs two launches instead of one per field.
#include "common.cuh"

namespace b200rl {

constexpr int kGatherBlock = 256;
constexpr int kChunkBytes = 16 * 1024;
constexpr int kNarrowRow = 256;  // rows up to this many bytes take the row-per-thread path

struct GatherParams {
  const uint8_t* src[B200RL_MAX_GATHER];
  uint8_t* dst[B200RL_MAX_GATHER];
  long long row_bytes[B200RL_MAX_GATHER];
  long long first_item[B200RL_MAX_GATHER + 1];  // prefix sum of work items per tensor
  int chunks_per_row[B200RL_MAX_GATHER];
  int n;
  const long long* idx;
  long long B, n_src_rows;
};

__global__ void __launch_bounds__(kGatherBlock) gather_wide_kernel(const GatherParams p) {
  pdl_trigger();  // the narrow-row launch behind this one is independent of it and may start right away
  const long long item = blockIdx.x;
  int t = 0;
  while (t + 1 < p.n && item >= p.first_item[t + 1]) ++t;
  const long long local = item - p.first_item[t];
  const long long b = local / p.chunks_per_row[t];
  const int chunk = (int)(local - b * p.chunks_per_row[t]);
  const long long row = p.idx[b];
  if (row < 0 || row >= p.n_src_rows) return;
  const long long rb = p.row_bytes[t];
  const long long begin = (long long)chunk * kChunkBytes;
  const long long bytes = (rb - begin < kChunkBytes) ? rb - begin : kChunkBytes;
  const uint8_t* s = p.src[t] + row * rb + begin;
  uint8_t* d = p.dst[t] + b * rb + begin;
  const int tid = threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(s);
    uint4* d4 = reinterpret_cast<uint4*>(d);
    const int n4 = (int)(bytes >> 4);
    constexpr int kIter = kChunkBytes / 16 / kGatherBlock;  // 4
    uint4 v[kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int o = tid + i * kGatherBlock;
      if (o < n4) v[i] = ldg_stream_u4(s4 + o);
    }
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int o = tid + i * kGatherBlock;
      if (o < n4) stg_stream_u4(d4 + o, v[i]);
    }
    for (long long o = ((long long)n4 << 4) + tid; o < bytes; o += kGatherBlock) d[o] = s[o];
  } else if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 3u) == 0) {
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(s);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(d);
    const int n1 = (int)(bytes >> 2);
    for (int o = tid; o < n1; o += kGatherBlock) d1[o] = __ldg(s1 + o);
    for (long long o = ((long long)n1 << 2) + tid; o < bytes; o += kGatherBlock) d[o] = s[o];
  } else {
    for (long long o = tid; o < bytes; o += kGatherBlock) d[o] = s[o];
  }
}

__global__ void __launch_bounds__(kGatherBlock) gather_narrow_kernel(const GatherParams p) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = (int)(e / p.B);  // tensor-major: threads of a warp share the tensor, walk rows
  const long long b = e - (long long)t * p.B;
  const long long row = e < p.B * p.n ? p.idx[b] : -1;
  if (row >= 0 && row < p.n_src_rows) {
    const long long rb = p.row_bytes[t];
    const uint8_t* s = p.src[t] + row * rb;
    uint8_t* d = p.dst[t] + b * rb;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)rb) & 3u) == 0) {
      for (long long o = 0; o < rb; o += 4)
        *reinterpret_cast<uint32_t*>(d + o) = __ldg(reinterpret_cast<const uint32_t*>(s + o));
    } else {
      for (long long o = 0; o < rb; ++o) d[o] = s[o];
    }
  }
  // Launched as a programmatic dependent of the wide-row grid (no data dependence, so the copies above run beside
  // it), this grid must not COMPLETE before the wide grid has: the next kernel on the stream (the trunk reading the
  // gathered observations) is ordered after this grid only.  Every thread waits here, which makes stream order
  // transitive; without the launch attribute the wait returns at once.
  pdl_wait();
}

// K0: one env step's [N, ...] slices -> row (*step % T) of the [T, N, ...] rollout buffers.
struct StoreParams {
  const uint8_t* src[B200RL_MAX_GATHER];
  uint8_t* dst[B200RL_MAX_GATHER];
  long long step_bytes[B200RL_MAX_GATHER];
  long long first_chunk[B200RL_MAX_GATHER + 1];
  const uint8_t* carry[B200RL_MAX_GATHER];  // non-null: after the row write, src[t] <- carry[t] (the next step's slice)
  const uint8_t* carry_or[B200RL_MAX_GATHER];  // non-null (with carry[t]): src[t] <- carry[t] | carry_or[t], bytewise
  int n;
  const long long* step_dev;
  long long T;
  // packed-observation field (pack_field >= 0): src[t] is [N, HW, Cp] float32, carry[t] the env's raw [N, C, HW] float32
  // observation; a work item is a tile of kPackCells cells of one sample instead of a 16 KB chunk
  int pack_field, pack_C, pack_HW, pack_Cp, pack_tiles;
  // advance != 0: the last CTA to finish writes *step_dev + 1 back (ticket: a zeroed int the caller keeps)
  long long* step_mut;
  int* ticket;
};
constexpr int kPackCells = 32;
constexpr int kPackMaxCp = 128;

// one chunk src -> dst by the whole CTA; every load is issued before the first store
__device__ __forceinline__ void copy_chunk(const uint8_t* s, uint8_t* d, long long bytes) {
  const int tid = threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(s);
    uint4* d4 = reinterpret_cast<uint4*>(d);
    const int n4 = (int)(bytes >> 4);
    constexpr int kIter = kChunkBytes / 16 / kGatherBlock;
    uint4 v[kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int o = tid + i * kGatherBlock;
      if (o < n4) v[i] = ldg_stream_u4(s4 + o);
    }
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int o = tid + i * kGatherBlock;
      if (o < n4) stg_stream_u4(d4 + o, v[i]);
    }
    for (long long o = ((long long)n4 << 4) + tid; o < bytes; o += kGatherBlock) d[o] = s[o];
  } else {
    for (long long o = tid; o < bytes; o += kGatherBlock) d[o] = s[o];
  }
}

__global__ void advance_step_kernel(long long* step) { *step += 1; }

// src <- a | b (bytewise), whole CTA
__device__ __forceinline__ void or_chunk(const uint8_t* a, const uint8_t* b, uint8_t* d, long long bytes) {
  for (long long o = threadIdx.x; o < bytes; o += kGatherBlock) d[o] = a[o] | b[o];
}

__global__ void __launch_bounds__(kGatherBlock) store_step_kernel(const StoreParams p) {
  __shared__ float s_tile[kPackCells][kPackMaxCp + 1];
  const long long item = blockIdx.x;
  int t = 0;
  while (t + 1 < p.n && item >= p.first_chunk[t + 1]) ++t;
  const long long sb = p.step_bytes[t];
  const long long step_raw = *p.step_dev;
  const long long step = step_raw % p.T;
  if (t == p.pack_field) {
    // this tile's cells of one sample: their packed rows go to the buffer row, then the env's raw [C, HW] planes of
    // the same cells are transposed through shared memory into [cells, Cp] rows (planes C.. written as zeros)
    const long long local = item - p.first_chunk[t];
    const long long n = local / p.pack_tiles;
    const int cell0 = (int)(local - n * p.pack_tiles) * kPackCells;
    const int cells = p.pack_HW - cell0 < kPackCells ? p.pack_HW - cell0 : kPackCells;
    const long long begin = (n * p.pack_HW + cell0) * p.pack_Cp * 4;
    const long long bytes = (long long)cells * p.pack_Cp * 4;
    uint8_t* region = const_cast<uint8_t*>(p.src[t]) + begin;
    copy_chunk(region, p.dst[t] + step * sb + begin, bytes);
    const float* raw = reinterpret_cast<const float*>(p.carry[t]) + n * p.pack_C * p.pack_HW + cell0;
    for (int i = threadIdx.x; i < p.pack_C * kPackCells; i += kGatherBlock) {
      const int c = i / kPackCells, j = i - c * kPackCells;  // adjacent threads read adjacent cells of one plane
      if (j < cells) s_tile[j][c] = raw[(long long)c * p.pack_HW + j];
    }
    __syncthreads();  // (also: every load of the old packed rows above has returned)
    float* out = reinterpret_cast<float*>(region);
    for (int i = threadIdx.x; i < cells * p.pack_Cp; i += kGatherBlock) {
      const int j = i / p.pack_Cp, c = i - j * p.pack_Cp;
      out[i] = c < p.pack_C ? s_tile[j][c] : 0.f;
    }
  } else {
    const long long begin = (item - p.first_chunk[t]) * kChunkBytes;
    const long long bytes = sb - begin < kChunkBytes ? sb - begin : kChunkBytes;
    const uint8_t* s = p.src[t] + begin;
    copy_chunk(s, p.dst[t] + step * sb + begin, bytes);
    // carry-over: this CTA, which alone touches this chunk of the step slice, overwrites it with the next step's once
    // every thread's loads of the old content have returned (the barrier: the two copies may split the chunk
    // differently across threads when their alignments differ)
    if (p.carry[t] != nullptr) {
      __syncthreads();
      if (p.carry_or[t] != nullptr) or_chunk(p.carry[t] + begin, p.carry_or[t] + begin, const_cast<uint8_t*>(s), bytes);
      else copy_chunk(p.carry[t] + begin, const_cast<uint8_t*>(s), bytes);
    }
  }
  if (p.step_mut != nullptr) {
    // every thread of this CTA has read the step index; the CTA that arrives last (all have read it) advances it
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(p.ticket, 1) == (int)gridDim.x - 1) {
        *p.step_mut = step_raw + 1;
        *p.ticket = 0;
      }
    }
  }
}

}  // namespace b200rl

extern "C" int b200rl_rollout_store_step(const void* const* src_host, void* const* dst_host,
                                         const int64_t* step_bytes_host, int n_tensors, const int64_t* step_dev,
                                         int64_t T, b200rl_stream_t stream) {
  return b200rl_rollout_store_step_carry(src_host, dst_host, step_bytes_host, nullptr, n_tensors, step_dev, T, stream);
}

namespace b200rl {
static int store_step_impl(const void* const* src_host, void* const* dst_host, const int64_t* step_bytes_host,
                           const void* const* carry_host, const void* const* carry_or_host, int n_tensors,
                           const b200rl_store_pack* pack, int64_t* step_dev, int64_t T, int advance, int32_t* ticket,
                           cudaStream_t stream) {
  B200RL_REQUIRE(src_host && dst_host && step_bytes_host && step_dev, "rollout_store_step: null pointer");
  B200RL_REQUIRE(n_tensors >= 0 && n_tensors <= B200RL_MAX_GATHER, "rollout_store_step: n_tensors=%d (max %d)",
                 n_tensors, B200RL_MAX_GATHER);
  B200RL_REQUIRE(T >= 1, "rollout_store_step: T=%lld", (long long)T);
  B200RL_REQUIRE(!advance || ticket, "rollout_store_step: advance needs a ticket counter");
  StoreParams p{};
  p.step_dev = reinterpret_cast<const long long*>(step_dev), p.T = T;
  p.pack_field = -1;
  long long chunks = 0;
  for (int t = 0; t < n_tensors; ++t) {
    B200RL_REQUIRE(src_host[t] && dst_host[t] && step_bytes_host[t] >= 0, "rollout_store_step: tensor %d is null", t);
    B200RL_REQUIRE(!(carry_or_host && carry_or_host[t]) || (carry_host && carry_host[t]),
                   "rollout_store_step: carry_or[%d] without carry[%d]", t, t);
    const bool packed = pack != nullptr && pack->field == t;
    if (step_bytes_host[t] == 0) {
      B200RL_REQUIRE(!packed, "rollout_store_step: the packed field is empty");
      continue;
    }
    const int k = p.n++;
    p.src[k] = static_cast<const uint8_t*>(src_host[t]), p.dst[k] = static_cast<uint8_t*>(dst_host[t]);
    p.carry[k] = carry_host ? static_cast<const uint8_t*>(carry_host[t]) : nullptr;
    p.carry_or[k] = carry_or_host ? static_cast<const uint8_t*>(carry_or_host[t]) : nullptr;
    p.step_bytes[k] = step_bytes_host[t];
    p.first_chunk[k] = chunks;
    if (packed) {
      B200RL_REQUIRE(pack->N >= 1 && pack->C >= 1 && pack->HW >= 1 && pack->Cp >= pack->C, "rollout_store_step: bad pack shape");
      B200RL_UNSUPPORTED(pack->Cp > kPackMaxCp, "rollout_store_step: Cp=%d exceeds %d", (int)pack->Cp, kPackMaxCp);
      B200RL_REQUIRE(step_bytes_host[t] == pack->N * pack->HW * pack->Cp * 4, "rollout_store_step: packed field is not [N, HW, Cp] float32");
      B200RL_REQUIRE(p.carry[k] != nullptr && p.carry_or[k] == nullptr, "rollout_store_step: the packed field carries the raw observation");
      B200RL_REQUIRE(((reinterpret_cast<uintptr_t>(p.src[k]) | reinterpret_cast<uintptr_t>(p.carry[k])) & 3u) == 0,
                     "rollout_store_step: packed field is not 4-byte aligned");
      p.pack_field = k, p.pack_C = (int)pack->C, p.pack_HW = (int)pack->HW, p.pack_Cp = (int)pack->Cp;
      p.pack_tiles = (int)((pack->HW + kPackCells - 1) / kPackCells);
      chunks += pack->N * p.pack_tiles;
    } else {
      chunks += (step_bytes_host[t] + kChunkBytes - 1) / kChunkBytes;
    }
    p.first_chunk[k + 1] = chunks;
  }
  if (advance) p.step_mut = reinterpret_cast<long long*>(step_dev), p.ticket = ticket;
  if (chunks == 0) {
    if (!advance) return B200RL_OK;
    // nothing to store, the step still advances: one CTA through the empty field list
    p.n = 0;
  }
  B200RL_UNSUPPORTED(chunks > 0x7fffffffLL, "rollout_store_step: %lld chunks", chunks);
  if (chunks == 0) {
    advance_step_kernel<<<1, 1, 0, stream>>>(p.step_mut);
    return check_launch("rollout_store_step");
  }
  store_step_kernel<<<(unsigned)chunks, kGatherBlock, 0, stream>>>(p);
  return check_launch("rollout_store_step");
}
}  // namespace b200rl

extern "C" int b200rl_rollout_store_step_carry(const void* const* src_host, void* const* dst_host,
                                               const int64_t* step_bytes_host, const void* const* carry_host,
                                               int n_tensors, const int64_t* step_dev, int64_t T,
                                               b200rl_stream_t stream) {
  return b200rl::store_step_impl(src_host, dst_host, step_bytes_host, carry_host, nullptr, n_tensors, nullptr,
                                 const_cast<int64_t*>(step_dev), T, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int b200rl_rollout_store_step_fused(const void* const* src_host, void* const* dst_host,
                                               const int64_t* step_bytes_host, const void* const* carry_host,
                                               const void* const* carry_or_host, int n_tensors,
                                               const b200rl_store_pack* pack, int64_t* step_dev, int64_t T, int advance,
                                               int32_t* ticket, b200rl_stream_t stream) {
  return b200rl::store_step_impl(src_host, dst_host, step_bytes_host, carry_host, carry_or_host, n_tensors, pack, step_dev,
                                 T, advance, ticket, (cudaStream_t)stream);
}

extern "C" int b200rl_gather_rows(const void* const* src_host, void* const* dst_host, const int64_t* row_bytes_host,
                                  int n_tensors, const int64_t* idx, int64_t B, int64_t n_src_rows,
                                  b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(src_host && dst_host && row_bytes_host && idx, "gather_rows: null pointer");
  B200RL_REQUIRE(n_tensors >= 0 && n_tensors <= B200RL_MAX_GATHER, "gather_rows: n_tensors=%d (max %d)", n_tensors,
                 B200RL_MAX_GATHER);
  B200RL_REQUIRE(B >= 0 && n_src_rows >= 0, "gather_rows: bad shape");
  if (B == 0 || n_tensors == 0) return B200RL_OK;
  GatherParams wide{}, narrow{};
  wide.idx = narrow.idx = reinterpret_cast<const long long*>(idx);
  wide.B = narrow.B = B;
  wide.n_src_rows = narrow.n_src_rows = n_src_rows;
  long long items = 0;
  for (int t = 0; t < n_tensors; ++t) {
    B200RL_REQUIRE(src_host[t] && dst_host[t] && row_bytes_host[t] >= 0, "gather_rows: tensor %d is null", t);
    if (row_bytes_host[t] == 0) continue;
    GatherParams& g = row_bytes_host[t] <= kNarrowRow ? narrow : wide;
    const int k = g.n++;
    g.src[k] = static_cast<const uint8_t*>(src_host[t]);
    g.dst[k] = static_cast<uint8_t*>(dst_host[t]);
    g.row_bytes[k] = row_bytes_host[t];
    if (&g == &wide) {
      g.chunks_per_row[k] = (int)((row_bytes_host[t] + kChunkBytes - 1) / kChunkBytes);
      g.first_item[k] = items;
      items += B * g.chunks_per_row[k];
      g.first_item[k + 1] = items;
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (wide.n) {
    B200RL_UNSUPPORTED(items > 0x7fffffffLL, "gather_rows: %lld chunks in one call", items);
    gather_wide_kernel<<<(unsigned)items, kGatherBlock, 0, s>>>(wide);
  }
  if (narrow.n) {
    const long long threads = B * narrow.n;
    const unsigned blocks = (unsigned)((threads + kGatherBlock - 1) / kGatherBlock);
    if (wide.n) {  // overlaps the wide launch (programmatic dependent launch; no data dependence between the two)
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(blocks), cfg.blockDim = dim3(kGatherBlock), cfg.stream = s;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &attr, cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, gather_narrow_kernel, narrow);
    } else {
      gather_narrow_kernel<<<blocks, kGatherBlock, 0, s>>>(narrow);
    }
  }
  return check_launch("gather_rows");
}

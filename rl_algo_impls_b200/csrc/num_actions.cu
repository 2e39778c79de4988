// a3: Batch.num_actions -- how many (cell, action plane) decisions a step actually offered.
//
// Replaces rollout/rollout.py:130-180 (num_actions / per_position_num_actions): numpy np.any / np.sum passes over the
// whole [T, N, HW, S] mask array per action plane on the host, once per rollout.  One launch here: a CTA takes a
// chunk of 256 cells of one step, stages the chunk's mask bytes in shared memory with coalesced 128-bit loads, one
// thread per cell counts the planes that have a valid entry -- with the value-dependent gating (an action plane only
// counts where the cell's reference plane chose the required value) -- and the chunk's count is added to the step's.
// Integer work: exact.  The pick_position term (log of the number of cells any pick head may choose) is finished by the
// caller from the second counter.
#include "common.cuh"

namespace b200rl {

constexpr int kNumActBlock = 256;  // = cells per chunk

struct NumActDev {
  const uint8_t* mask;
  const uint8_t* pick_mask;
  const void* actions;
  int act_dtype;
  long long R, HW;
  int A, S, n_pick, chunks, gated;
  int nvec[B200RL_MAX_HEADS], off[B200RL_MAX_HEADS], gate_ref[B200RL_MAX_HEADS], gate_val[B200RL_MAX_HEADS];
  int* cells_out;
  int* picks_out;
};

__device__ __forceinline__ int action_at(const void* base, int dtype, long long i) {
  switch (dtype) {
    case B200RL_U8: return (int)__ldg(static_cast<const uint8_t*>(base) + i);
    case B200RL_I32: return (int)__ldg(static_cast<const int32_t*>(base) + i);
    default: return (int)__ldg(static_cast<const long long*>(base) + i);
  }
}

__global__ void __launch_bounds__(kNumActBlock) num_actions_kernel(const __grid_constant__ NumActDev G) {
  extern __shared__ __align__(16) uint8_t image[];  // the chunk's mask bytes, skewed like the global address
  __shared__ int s_part[2][kNumActBlock / 32];
  const int tid = threadIdx.x;
  const long long b = blockIdx.x / G.chunks;
  const int chunk = (int)(blockIdx.x - b * G.chunks);
  const int cell0 = chunk * kNumActBlock;
  const int cells = (int)min((long long)kNumActBlock, G.HW - cell0);
  const uint8_t* g = G.mask + (b * G.HW + cell0) * G.S;
  const uint32_t bytes = (uint32_t)cells * (uint32_t)G.S;
  const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
  uint32_t head = skew ? 16u - skew : 0u;
  if (head > bytes) head = bytes;
  const uint32_t body = (bytes - head) & ~15u;
  for (uint32_t i = tid; i < head; i += kNumActBlock) image[skew + i] = g[i];
  for (uint32_t i = tid; i < body / 16u; i += kNumActBlock)
    reinterpret_cast<uint4*>(image + skew + head)[i] = ldg_stream_u4(reinterpret_cast<const uint4*>(g + head) + i);
  for (uint32_t i = head + body + tid; i < bytes; i += kNumActBlock) image[skew + i] = g[i];
  __syncthreads();
  int count = 0, pick_any = 0;
  if (tid < cells) {
    const uint8_t* m = image + skew + (uint32_t)tid * (uint32_t)G.S;
    if (!G.gated) {  // rollout.py:160-161: cells with any valid entry at all
      uint32_t any = 0u;
      for (int k = 0; k < G.S; ++k) any |= m[k];
      count = any != 0u;
    } else {  // :163-179: per action plane, gated by the reference plane's chosen value
      const long long abase = (b * G.HW + cell0 + tid) * G.A;
      for (int h = 0; h < G.A; ++h) {
        uint32_t any = 0u;
        for (int k = 0; k < G.nvec[h]; ++k) any |= m[G.off[h] + k];
        if (any && G.gate_ref[h] >= 0 && action_at(G.actions, G.act_dtype, abase + G.gate_ref[h]) != G.gate_val[h]) any = 0u;
        count += any != 0u;
      }
    }
    for (int kp = 0; kp < G.n_pick; ++kp) pick_any |= G.pick_mask[(b * G.n_pick + kp) * G.HW + cell0 + tid] != 0;
  }
  count = warp_sum(count), pick_any = warp_sum(pick_any);
  if ((tid & 31) == 0) s_part[0][tid >> 5] = count, s_part[1][tid >> 5] = pick_any;
  __syncthreads();
  if (tid == 0) {
    int c = 0, p = 0;
    for (int w = 0; w < kNumActBlock / 32; ++w) c += s_part[0][w], p += s_part[1][w];
    if (G.chunks == 1) {
      G.cells_out[b] = c;
      if (G.picks_out) G.picks_out[b] = p;
    } else {
      atomicAdd(G.cells_out + b, c);
      if (G.picks_out) atomicAdd(G.picks_out + b, p);
    }
  }
}

}  // namespace b200rl

extern "C" int b200rl_gridnet_num_actions(const b200rl_gridnet_desc* d, const uint8_t* mask, const uint8_t* pick_mask,
                                          const void* actions, int32_t* cells_out, int32_t* picks_out,
                                          b200rl_stream_t stream) {
  using namespace b200rl;
  B200RL_REQUIRE(d && mask && cells_out, "gridnet_num_actions: null pointer");
  B200RL_REQUIRE(d->B >= 0 && d->HW >= 1 && d->A >= 1 && d->A <= B200RL_MAX_HEADS && d->n_pick >= 0 && d->nvec_host,
                 "gridnet_num_actions: bad shape");
  B200RL_REQUIRE(d->n_pick == 0 || (pick_mask && picks_out), "gridnet_num_actions: pick tensors are null");
  if (d->B == 0) return B200RL_OK;
  NumActDev G{};
  G.mask = mask, G.pick_mask = pick_mask, G.actions = actions, G.act_dtype = d->act_dtype;
  G.R = d->B, G.HW = d->HW, G.A = d->A, G.n_pick = d->n_pick;
  int S = 0;
  for (int h = 0; h < d->A; ++h) {
    B200RL_REQUIRE(d->nvec_host[h] >= 1, "gridnet_num_actions: nvec[%d]=%d", h, d->nvec_host[h]);
    G.nvec[h] = d->nvec_host[h], G.off[h] = S, S += d->nvec_host[h];
    const int gr = d->gate_ref_host ? d->gate_ref_host[h] : -1;
    B200RL_REQUIRE(gr < d->A, "gridnet_num_actions: gate_ref[%d] out of range", h);
    G.gate_ref[h] = gr, G.gate_val[h] = (gr >= 0 && d->gate_val_host) ? d->gate_val_host[h] : 0;
    if (gr >= 0) G.gated = 1;
  }
  // the reference takes the per-plane path whenever a subaction mask is configured at all (rollout.py:160-163);
  // a descriptor without gates is the "cells with any valid entry" count
  B200RL_REQUIRE(!G.gated || actions != nullptr, "gridnet_num_actions: gated planes need the actions");
  B200RL_UNSUPPORTED(G.gated && d->act_dtype != B200RL_U8 && d->act_dtype != B200RL_I32 && d->act_dtype != B200RL_I64,
                     "gridnet_num_actions: action dtype %d", d->act_dtype);
  G.S = S;
  G.chunks = (int)((d->HW + kNumActBlock - 1) / kNumActBlock);
  G.cells_out = cells_out, G.picks_out = d->n_pick ? picks_out : nullptr;
  const long long ctas = d->B * G.chunks;
  B200RL_UNSUPPORTED(ctas > 0x7fffffffLL, "gridnet_num_actions: %lld chunks in one call", ctas);
  const size_t smem = (size_t)kNumActBlock * S + 32;
  B200RL_UNSUPPORTED(smem > 200 * 1024, "gridnet_num_actions: sum(nvec)=%d is too wide", S);
  cudaStream_t s = (cudaStream_t)stream;
  if (G.chunks > 1) {
    cudaMemsetAsync(cells_out, 0, (size_t)d->B * sizeof(int32_t), s);
    if (G.picks_out) cudaMemsetAsync(picks_out, 0, (size_t)d->B * sizeof(int32_t), s);
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(num_actions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("gridnet_num_actions: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
      return B200RL_ECUDA;
    }
  }
  num_actions_kernel<<<(unsigned)ctas, kNumActBlock, smem, s>>>(G);
  return check_launch("gridnet_num_actions");
}

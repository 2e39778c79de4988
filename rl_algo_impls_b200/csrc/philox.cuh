// Counter-based RNG for rollout-time sampling: Philox4x32-10 (Salmon et al., SC'11).
// Stateless: (seed, counter) -> 4 x 32 random bits, so every (sample, cell, head) draws its
// own stream without any RNG state in memory.
#pragma once
#include <stdint.h>

namespace b200rl {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ inline Philox4 philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// 24 random bits -> uniform in (0, 1), never 0 or 1
__host__ __device__ inline float u01(uint32_t bits) { return ((float)(bits >> 8) + 0.5f) * (1.0f / 16777216.0f); }

}  // namespace b200rl

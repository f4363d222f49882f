// Counter-based dropout masks (Philox4x32-10, Salmon et al. 2011; the generator torch / cuRAND use): the mask of element idx
// of a dropout site is a pure function of (seed, offset, idx), so the backward pass REGENERATES it instead of storing it, and
// the CPU test oracle can build the identical mask (tests/philox_ref.py).
//   keep(idx)  <=>  philox(seed, offset + idx / 4).word[idx % 4] >= floor(p * 2^32)
#pragma once
#include <stdint.h>

namespace ctb {

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// four random words for 64-bit counter `ctr` under 64-bit key `seed`
__host__ __device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
// keep flag of one element (one Philox call; callers that walk 4 consecutive indices should call philox4 once)
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thresh) {
  uint32_t w[4];
  philox4(seed, offset + (idx >> 2), w);
  return w[idx & 3] >= thresh;
}

}  // namespace ctb

// Philox4x32-10 counter-based uniforms (Salmon et al., SC'11).  The counter
// layout is documented in oracle/philox.py and must stay in sync with it:
//   draws     : ctr = (l, row, m | part << 16, step)   part 0 -> v 0..3, part 1 -> v 4
//   selection : ctr = (m / 4, row, 0, step | 1 << 24)  word m % 4
//   key       = (seed lo, seed hi);  u = (word >> 8) * 2^-24  in [0,1)
#pragma once
#include <stdint.h>

namespace svdd {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float philox_uniform(uint32_t word) {
  return (float)(word >> 8) * 5.9604644775390625e-08f;  // 2^-24, exact in fp32
}

}  // namespace svdd

// digits.cuh - the balanced base-256 digit recoding shared by the slicing kernel (chi2_ozaki.cuh) and by stage 2 when it
// writes the digit planes of its residual rows itself (friedmann.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cosmolike {

// Balanced base-256 digits of four fixed-point values at once.  Adding 0x80 at every digit position below the top one
// turns the carry chain of the recoding d = ((v + 128) & 255) - 128, v <- (v + 128) >> 8 into one 64-bit addition; the
// digit bytes are then the bytes of the sum with their top bit flipped, and the top digit is what remains above them.
// w[s] = the four digits of plane s (plane 0 = most significant), one byte per value: a 4 x S byte transpose (PRMT).
template <int S>
__device__ __forceinline__ void oz_digits4(const long long (&v)[4], uint32_t (&w)[S]) {
  constexpr unsigned long long kBias = 0x0080808080808080ULL >> (8 * (8 - S));   // 0x80 in bytes 0 .. S-2
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const unsigned long long u = ((unsigned long long)v[q] + kBias) ^ kBias;
    lo[q] = (uint32_t)u; hi[q] = (uint32_t)(u >> 32);
  }
  uint32_t b[8];
  {
    const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[0], lo[1], 0x7362);
    const uint32_t t2 = __byte_perm(lo[2], lo[3], 0x5140), t3 = __byte_perm(lo[2], lo[3], 0x7362);
    b[0] = __byte_perm(t0, t2, 0x5410); b[1] = __byte_perm(t0, t2, 0x7632);
    b[2] = __byte_perm(t1, t3, 0x5410); b[3] = __byte_perm(t1, t3, 0x7632);
  }
  {
    const uint32_t t0 = __byte_perm(hi[0], hi[1], 0x5140), t1 = __byte_perm(hi[0], hi[1], 0x7362);
    const uint32_t t2 = __byte_perm(hi[2], hi[3], 0x5140), t3 = __byte_perm(hi[2], hi[3], 0x7362);
    b[4] = __byte_perm(t0, t2, 0x5410); b[5] = __byte_perm(t0, t2, 0x7632);
    b[6] = __byte_perm(t1, t3, 0x5410); b[7] = __byte_perm(t1, t3, 0x7632);
  }
#pragma unroll
  for (int s = 0; s < S; s++) w[s] = b[S - 1 - s];
}

}  // namespace cosmolike

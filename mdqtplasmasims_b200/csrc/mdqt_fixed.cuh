// mdqt_fixed.cuh -- periodic fixed-point coordinates: x -> round(frac(x/L) * 2^64) mod 2^64 as a 64-bit integer.
// In this representation the two's-complement difference of two coordinates IS their minimum-image separation
// (the reference's d -= L*round(d/L), SU:218-220), exactly and for any input. Kernels that move ions (K2, the
// velocity-Verlet position kernel) store the fixed-point copy next to R; the pair kernels read only the copy.
#pragma once
#include <cuda_runtime.h>

namespace mdqt {

#define MDQT_MAGIC 6755399441055744.0 /* 1.5 * 2^52: adding it rounds to nearest integer */
#define MDQT_2P62 4611686018427387904.0
#define MDQT_2P64 18446744073709551616.0

// x/L is formed in double-double (product residual by FMA + the rounding error of 1/L), so two nearby ions keep
// their exact separation (the reference's x_i - x_j is exact for nearby ions by Sterbenz' lemma); accurate to ~L*2^-62.
__device__ __forceinline__ long long to_fixed(double x, double invL, double invL_lo) {
  double q = x * invL;
  double ql = fma(x, invL, -q) + x * invL_lo;
  double n = (q + MDQT_MAGIC) - MDQT_MAGIC;  // rint(q), |q| < 2^51
  double f = q - n;                          // exact, in [-1/2, 1/2]
  long long a = __double2ll_rn(f * MDQT_2P62) + __double2ll_rn(ql * MDQT_2P62);
  return (long long)((unsigned long long)a << 2);
}

}  // namespace mdqt

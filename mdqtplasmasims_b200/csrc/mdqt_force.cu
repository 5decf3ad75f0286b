// mdqt_force.cu -- K1 (all-pairs minimum-image Yukawa force), K3 (potential energy) and the FP64 DFMA-chain
// peak probe, hand-written for sm_100a.
//
// Replaces forces() (reference laserCoolingPlusExpansionMDQTSpeedUp.cpp:192-236), calculateAccelerations()
// (MonteCarloFollowedByMDAndTempAnisotropy.cpp:387-448) and Epotential() (SU:244-281). The reference walks the
// N(N-1)/2 unordered pairs with a racy OpenMP scatter; here every ion row i gathers over ALL j (N^2 ordered
// pair-interactions per call), which needs no scatter and sums in a fixed order.
//
// Bound: the FP64 pipe (64 DFMA/clk/SM). Memory traffic is negligible (24 B per j per CTA, staged in shared memory
// and broadcast). Per ordered pair the FP64 pipe executes ~35 instructions:
//   3 (delta) + 6 (minimum image: compare + shift) + 3 (r^2) + 5 (rsqrt: MUFU seed + one 3rd-order Newton step)
//   + 1 (r) + 10 (exp(-kappa r): magic-number range reduction to |rr|<=ln2/256, 128-entry 2^(j/128) table in shared
//   memory, degree-5 polynomial, exponent patched with integer adds) + 4 (prefactor) + 3 (accumulate).
// The cut-off / self-pair mask is done with integer compares on the bit pattern of r^2 so that it costs no FP64 issue.
#include "mdqt_internal.h"
#include <math.h>

#ifndef MDQT_PAIR_VARIANT
#define MDQT_PAIR_VARIANT 2
#endif

namespace mdqt {

__constant__ double c_exp2tab[kExpTable];

void upload_exp_table() {
  double tab[kExpTable];
  for (int j = 0; j < kExpTable; j++) tab[j] = (double)exp2l((long double)j / (long double)kExpTable);
  cudaMemcpyToSymbol(c_exp2tab, tab, sizeof(tab));
}

struct PairConsts {
  double L, halfL, invL, kappa, negkappa, nk_scale, negc, rc2;
  long long rc2bits_m1;
};

__device__ __forceinline__ PairConsts make_consts(const ForceArgs& a) {
  PairConsts c;
  c.L = a.L; c.halfL = a.halfL; c.invL = a.invL; c.kappa = a.kappa; c.negkappa = -a.kappa;
  c.nk_scale = -a.kappa * 184.66496523378731614207035916824219;  // 128 * log2(e)
  c.negc = -0.0054152123481245727298221259488920044;     // -ln2 / 128
  c.rc2bits_m1 = __double_as_longlong(a.rc2) - 1; c.rc2 = a.rc2;
  return c;
}

#define MDQT_MAGIC 6755399441055744.0 /* 1.5 * 2^52: adding it rounds to nearest integer */

template <bool WRAPPED>
__device__ __forceinline__ double min_image(double d, const PairConsts& c) {
  if (WRAPPED && MDQT_PAIR_VARIANT == 0) {
    // coordinates in [0,L] => |d| <= L => round(d/L) is -1, 0 or +1 (SU:218): one conditional shift is exact
    return (fabs(d) > c.halfL) ? d - copysign(c.L, d) : d;
  } else {
    double t = fma(d, c.invL, MDQT_MAGIC);
    double n = t - MDQT_MAGIC;
    return fma(-c.L, n, d);
  }
}

// Returns through (rinv, ef, valid): 1/r, exp(-kappa r), and whether 0 < r^2 < rcut^2.
__device__ __forceinline__ void pair_core(double dx, double dy, double dz, const PairConsts& c, const double* tab,
                                          double& rinv, double& ef, bool& valid) {
  double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
#if MDQT_PAIR_VARIANT >= 2
  valid = (r2 < c.rc2) && (__double2hiint(r2) != 0);  // one DSETP + one ISETP (r2 > 0 <=> high word != 0)
#else
  valid = (unsigned long long)(__double_as_longlong(r2) - 1) < (unsigned long long)c.rc2bits_m1;
#endif
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));  // MUFU.RSQ64H: ~2^-20 relative
  double t = r2 * y;
  double e = fma(-t, y, 1.0);                 // 1 - r2 y^2
  double p = fma(0.375, e, 0.5);
  double ye = y * e;
  y = fma(ye, p, y);                          // y (1 + e/2 + 3 e^2/8): error O(e^3) ~ 2^-60
  double r = r2 * y;
  // exp(x), x = -kappa r <= 0:  x = (128 q + idx) ln2/128 + rr
  double tt = fma(r, c.nk_scale, MDQT_MAGIC);
  int n = __double2loint(tt);
  double nd = tt - MDQT_MAGIC;
  double x = c.negkappa * r;
  double rr = fma(nd, c.negc, x);
  double q = fma(rr, 8.3333333333333332e-03, 4.1666666666666664e-02);
  q = fma(q, rr, 1.6666666666666666e-01);
  q = fma(q, rr, 0.5);
  q = fma(q, rr, 1.0);
  q = fma(q, rr, 1.0);
  double T = tab[n & (kExpTable - 1)];
  int hi = __double2hiint(T) + ((n & ~(kExpTable - 1)) << 13);  // += floor(n/128) << 20
  ef = __hiloint2double(hi, __double2loint(T)) * q;
  rinv = y;
}

constexpr int kTJ = 512;  // j positions staged per pass

template <int IPT, bool WRAPPED, bool EPOT>
__global__ void __launch_bounds__(kForceThreads) k_pairs(ForceArgs a, double* __restrict__ block_partials) {
  __shared__ double2 sxy[kTJ];
  __shared__ double sz[kTJ];
  __shared__ double stab[kExpTable];
  __shared__ double sred[kForceThreads / 32];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int b = blockIdx.z, js = blockIdx.y, tile = blockIdx.x;
  const PairConsts c = make_consts(a);
  const double* __restrict__ X = a.R + (size_t)b * 3 * a.ld;
  const double* __restrict__ Y = X + a.ld;
  const double* __restrict__ Z = Y + a.ld;
  for (int k = tid; k < kExpTable; k += kForceThreads) stab[k] = c_exp2tab[k];

  int irow[IPT];
  double xi[IPT], yi[IPT], zi[IPT], ax[IPT], ay[IPT], az[IPT];
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    irow[k] = a.row0 + tile * (kForceThreads * IPT) + k * kForceThreads + tid;
    bool ok = irow[k] < a.row0 + a.nrows;
    xi[k] = ok ? X[irow[k]] : 3e300;
    yi[k] = ok ? Y[irow[k]] : 3e300;
    zi[k] = ok ? Z[irow[k]] : 3e300;
    ax[k] = ay[k] = az[k] = 0.0;
  }
  const int jbeg = js * a.jlen;
  const int jend = min(a.N, jbeg + a.jlen);
  for (int jc = jbeg; jc < jend; jc += kTJ) {
    __syncthreads();
    for (int k = tid; k < kTJ; k += kForceThreads) {
      int j = jc + k;
      bool ok = j < jend;
      sxy[k] = make_double2(ok ? X[j] : 1e300, ok ? Y[j] : 1e300);
      sz[k] = ok ? Z[j] : 1e300;
    }
    __syncthreads();
    const int cnt = min(kTJ, jend - jc);
    const int nloop = (cnt + 3) & ~3;  // padded entries are sentinels (masked by the cut-off)
#pragma unroll 4
    for (int jj = 0; jj < nloop; jj++) {
      const double2 pxy = sxy[jj];
      const double pz = sz[jj];
#pragma unroll
      for (int k = 0; k < IPT; k++) {
        double dx = min_image<WRAPPED>(xi[k] - pxy.x, c);
        double dy = min_image<WRAPPED>(yi[k] - pxy.y, c);
        double dz = min_image<WRAPPED>(zi[k] - pz, c);
        double rinv, ef;
        bool valid;
        pair_core(dx, dy, dz, c, stab, rinv, ef, valid);
        if (EPOT) {
          double u = ef * rinv;                       // exp(-r/lDeb)/r (SU:268)
          ax[k] += valid ? u : 0.0;
        } else {
          double f = (ef * (rinv * rinv)) * (rinv + c.kappa);  // (1/r + 1/lDeb) exp(-r/lDeb)/r^2 (SU:224)
          f = valid ? f : 0.0;
          ax[k] = fma(f, dx, ax[k]);
          ay[k] = fma(f, dy, ay[k]);
          az[k] = fma(f, dz, az[k]);
        }
      }
    }
  }

  if (EPOT) {
    // fixed-order block reduction -> one partial per CTA
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; k++) s += (irow[k] < a.row0 + a.nrows) ? ax[k] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) sred[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kForceThreads / 32; w++) tot += sred[w];
      block_partials[((size_t)b * gridDim.y + js) * gridDim.x + tile] = tot;
    }
    return;
  }

  if (a.nsplit == 1) {
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (irow[k] < a.row0 + a.nrows) {
        double* Fb = a.F + (size_t)b * 3 * a.ld;
        Fb[irow[k]] = ax[k]; Fb[a.ld + irow[k]] = ay[k]; Fb[2 * a.ld + irow[k]] = az[k];
      }
    return;
  }
  // j-split: store the partial, the last CTA of this (trajectory, i-tile) sums all partials in ascending split
  // order -> deterministic, independent of arrival order and of how many ranks share the rows.
#pragma unroll
  for (int k = 0; k < IPT; k++)
    if (irow[k] < a.row0 + a.nrows) {
      double* Fp = a.Fpart + ((size_t)js * a.B + b) * 3 * a.ld;
      __stcg(&Fp[irow[k]], ax[k]); __stcg(&Fp[a.ld + irow[k]], ay[k]); __stcg(&Fp[2 * a.ld + irow[k]], az[k]);
    }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned* ctr = a.counters + (size_t)b * gridDim.x + tile;
    unsigned old = atomicAdd(ctr, 1u);
    s_last = (old == (unsigned)a.nsplit - 1);
    if (s_last) *ctr = 0;  // self-reset for the next call
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < IPT; k++)
    if (irow[k] < a.row0 + a.nrows) {
      double sx = 0.0, sy = 0.0, szz = 0.0;
      for (int s = 0; s < a.nsplit; s++) {
        const double* Fp = a.Fpart + ((size_t)s * a.B + b) * 3 * a.ld;
        sx += __ldcg(&Fp[irow[k]]); sy += __ldcg(&Fp[a.ld + irow[k]]); szz += __ldcg(&Fp[2 * a.ld + irow[k]]);
      }
      double* Fb = a.F + (size_t)b * 3 * a.ld;
      Fb[irow[k]] = sx; Fb[a.ld + irow[k]] = sy; Fb[2 * a.ld + irow[k]] = szz;
    }
}

// two rows per thread once there are enough rows to fill the machine (halves shared-memory traffic per pair);
// decided by the planner from (N, B) only
static int pick_ipt(const ForceArgs& a) { return a.ipt == 2 ? 2 : 1; }

template <bool EPOT>
static void launch_pairs(const ForceArgs& a, double* partials, cudaStream_t s) {
  int ipt = pick_ipt(a);
  dim3 grid((a.nrows + kForceThreads * ipt - 1) / (kForceThreads * ipt), a.nsplit, a.B);
  if (ipt == 2) {
    if (a.wrapped) k_pairs<2, true, EPOT><<<grid, kForceThreads, 0, s>>>(a, partials);
    else k_pairs<2, false, EPOT><<<grid, kForceThreads, 0, s>>>(a, partials);
  } else {
    if (a.wrapped) k_pairs<1, true, EPOT><<<grid, kForceThreads, 0, s>>>(a, partials);
    else k_pairs<1, false, EPOT><<<grid, kForceThreads, 0, s>>>(a, partials);
  }
}

void launch_forces(const ForceArgs& a, cudaStream_t s) { launch_pairs<false>(a, nullptr, s); }

int epot_partials_needed(const ForceArgs& a) {
  int tiles = (a.nrows + kForceThreads - 1) / kForceThreads;  // upper bound (IPT = 1)
  return tiles * a.nsplit * a.B;
}

__global__ void k_epot_final(const double* __restrict__ partials, int per_traj, double scale, double* __restrict__ out) {
  __shared__ double sred[256];
  const double* p = partials + (size_t)blockIdx.x * per_traj;
  double s = 0.0;
  for (int k = threadIdx.x; k < per_traj; k += 256) s += p[k];
  sred[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sred[0] * scale;
}

void launch_epot(const ForceArgs& a, double* partials, double* result, cudaStream_t s) {
  launch_pairs<true>(a, partials, s);
  int ipt = pick_ipt(a);
  int tiles = (a.nrows + kForceThreads * ipt - 1) / (kForceThreads * ipt);
  // ordered pairs counted twice -> 1/2; per particle -> 1/N (SU:272)
  k_epot_final<<<a.B, 256, 0, s>>>(partials, tiles * a.nsplit, 0.5 / (double)a.N, result);
}

// ---- FP64 peak probe: 8 independent DFMA chains per thread ----
__global__ void __launch_bounds__(256) k_dfma_chain(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

double run_fp64_peak(cudaStream_t s) {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, s);
    k_dfma_chain<<<blocks, threads, 0, s>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}  // namespace mdqt

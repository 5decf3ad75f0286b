// mdqt_force.cu -- K1 (all-pairs minimum-image Yukawa force), K3 (potential energy) and the FP64 DFMA-chain
// peak probe, hand-written for sm_100a.
//
// Replaces forces() (reference laserCoolingPlusExpansionMDQTSpeedUp.cpp:192-236), calculateAccelerations()
// (MonteCarloFollowedByMDAndTempAnisotropy.cpp:387-448) and Epotential() (SU:244-281). The reference walks the
// N(N-1)/2 unordered pairs with a racy OpenMP scatter; here every ion row i gathers over ALL j (N^2 ordered
// pair-interactions per call), which needs no scatter and sums in a fixed order.
//
// Bound: the FP64 pipe / the sub-partition issue port (an FP64 warp instruction holds the port for 2 cycles). Memory
// traffic is negligible (24 B per j per pass, staged in shared memory and broadcast). Pair arithmetic (23 FP64-pipe
// instructions per ordered pair):
//   PERIODIC FIXED POINT: coordinates are kept as 64-bit integers in units of L/2^64, so that the two's-complement
//     difference x_i - x_j IS the minimum image (the reference's d -= L*round(d/L), SU:218-220) -- for free, for any
//     (wrapped or unwrapped) input, and exactly: differences are formed without rounding and converted to double with
//     53-bit relative precision. Delta + minimum image cost 2 integer adds + 1 int->fp64 conversion per component and
//     NO FP64-pipe instruction.
//   r^2 (3), 1/r = MUFU.RSQ64H seed + one 3rd-order Newton step (5), r (1); the cut-off r < L/2 is ONE unsigned integer
//   compare of the high word of r^2 (r^2 < 2^126 fixed-point units; other cut-offs take one DSETP) and the self pair needs
//   no predicate (r^2 carries +1 unit^2, its f is multiplied by Delta = 0); exp(-kappa r) by a magic-number reduction in
//   units of ln2/1024 + a 1024-entry 2^(j/1024) table in shared memory (high words pre-biased: the exponent patch is one
//   integer add) + a cubic (6); prefactor (5); accumulate (3).
// Two kernels share that loop: k_pairs (CTA tiles of 128 or 32 rows, large N) and k_pairs_items (a persistent grid whose
// warps walk a static list of (trajectory, 32-row group, j chunk) items; small/medium N at any batch size).
#include "mdqt_internal.h"
#include "mdqt_fixed.cuh"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>

namespace mdqt {

// Build-time variants of the pair arithmetic (A/B'd on B200, scripts/ab_pairs.py; defaults = the measured best):
//   MDQT_EXP_SCALED 1: the exp argument is reduced in units of ln2/T directly from r (one FMA less than forming x = -kappa r)
//   MDQT_VALID_INT  1: when rcut == L/2 the test 0 < r^2 < rcut^2 is ONE unsigned compare of the high word of r^2
//                      (r^2 < 2^126 in fixed-point units), no FP64-pipe DSETP
//   kExpTable 1024 (mdqt_internal.h): degree-3 polynomial instead of degree-5 (two FMAs less)
#ifndef MDQT_EXP_SCALED
#define MDQT_EXP_SCALED 1
#endif
#ifndef MDQT_VALID_INT
#define MDQT_VALID_INT 1
#endif
//   MDQT_RSQRT_ORDER 3 (default): MUFU.RSQ64H seed + one third-order step (1/r to 1e-16).
//                    2: 1/r and r from ONE coupled second-order step on a seed for 1/sqrt(2 r^2) (4 FP64 + 1 integer instruction
//                       instead of 6 FP64). Measured on B200 (profiles/r02o_ab_pairs.log): +6.0 % at N = 1e5 (5.60e11 pairs/s),
//                       +8.3 % at 64 trajectories, K1 28.6 -> 26.9 us at N = 3500 -- but the seed only sees the HIGH word of r^2
//                       (relative input error up to 2^-20), so the second-order remainder 3/8 e^2 reaches 3.4e-13 in 1/r and the
//                       per-ion force error 5e-13 ... 9.5e-13 of sum_j |f_ij| (third order: 3e-16 ... 8e-16): no margin against the
//                       1e-12 parity budget. NOT taken; kept as a build option for the record.
//   MDQT_PRED_ACC    1: the cut-off predicate guards the three accumulating FMAs instead of zeroing f with a 64-bit select. ptxas
//                       turns the predicated FMAs back into selects on each accumulator (6 FSEL per pair instead of 2): off.
#ifndef MDQT_RSQRT_ORDER
#define MDQT_RSQRT_ORDER 3
#endif
#ifndef MDQT_PRED_ACC
#define MDQT_PRED_ACC 0
#endif
constexpr int kExpShift = (kExpTable == 1024) ? 10 : 13;  // 20 - log2(kExpTable)
static_assert(kExpTable == 128 || kExpTable == 1024, "exp table must have 128 or 1024 entries");

__device__ double c_exp2tab[kExpTable];  // global (L2-resident), not __constant__: the CTA prologue reads it with a per-thread index

int pdl_mode() {
  // Programmatic dependent launch between the force and substep kernels. With the persistent item kernel it hides the dependent's
  // launch latency and prologue behind the tail WHEN ALL WARPS FINISH TOGETHER (one trajectory whose items fill one round of the
  // resident warps; N = 3500: 54.4 -> 53.5 us per MD step) -- but whenever the force kernel's warps finish at different times
  // (N = 3653 planned for 3500: two item rounds; batches) the dependent's CTAs pile onto the SMs that free up first and the substep
  // kernel runs unbalanced: 79.9 instead of 67.6 us; with the CTA-tile kernel it was slower everywhere (114 vs 74 us). So the
  // default is decided per handle from the plan (plan_force); MDQT_PDL=1 / 0 forces it on / off.
  static const int mode = [] { const char* e = getenv("MDQT_PDL"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
  return mode;
}

bool cluster_enabled() {
  // opt-in (MDQT_CLUSTER=1). Measured on B200: -2 us at N = 2048 (64 clusters of 4, one wave), but +8 us at N = 3000
  // (94 clusters of 3) and N = 3500/4096 (more clusters than fit at once): cluster placement inside a GPC costs more balance
  // than the cheaper reduction saves, and where it does so is hard to predict -- the default stays the global-memory path.
  static const bool on = [] { const char* e = getenv("MDQT_CLUSTER"); return e && e[0] == '1'; }();
  return on;
}

void upload_exp_table() {
  double tab[kExpTable];
  // entry j holds the bit pattern of 2^(j/T) with (j << kExpShift) pre-subtracted from its high word, so that the kernel
  // patches the exponent with a single add of (n << kExpShift), n = T q + j  (high word += q << 20)
  for (int j = 0; j < kExpTable; j++) {
    tab[j] = (double)exp2l((long double)j / (long double)kExpTable);
    unsigned long long bits;
    memcpy(&bits, &tab[j], 8);
    bits -= (unsigned long long)((unsigned)j << kExpShift) << 32;
    memcpy(&tab[j], &bits, 8);
  }
  cudaMemcpyToSymbol(c_exp2tab, tab, sizeof(tab));
}

// Everything the pair loop needs, in the length unit `u` of the chosen formulation (u = 1 for variant 2,
// u = L/2^64 for the fixed-point variant): kappa*u, (rcut/u)^2, and the exp reduction constants.
struct PairConsts {
  double kappa_u, negkappa_u, nk_scale, negc, rc2_u;
  double c1, c2, c3, c4, c5;  // exp polynomial coefficients in the reduced variable
  double out_scale;  // 1/u^2 for forces, 1/u for the potential energy
};

__device__ __forceinline__ PairConsts make_consts(const ForceArgs& a, bool epot) {
  PairConsts c;
  const double u = a.L / MDQT_2P64;
  const double T = (double)kExpTable;
#if MDQT_RSQRT_ORDER == 2
  // the pair loop works with r/sqrt2 and 1/(sqrt2 r) (pair_core): kappa and the output scale absorb the factors
  const double s2 = 1.4142135623730950488016887242096981;
  c.kappa_u = a.kappa * u / s2; c.negkappa_u = -(a.kappa * u) * s2;
  c.nk_scale = -(a.kappa * u) * s2 * (T * 1.4426950408889634073599246810018921);
#else
  c.kappa_u = a.kappa * u; c.negkappa_u = -c.kappa_u;
  c.nk_scale = -c.kappa_u * (T * 1.4426950408889634073599246810018921);  // T * log2(e)
#endif
  c.negc = -0.69314718055994530941723212145817657 / T;                   // -ln2 / T
#if MDQT_EXP_SCALED
  const double w = 0.69314718055994530941723212145817657 / T;            // polynomial in rr' = rr / w
  c.c1 = w; c.c2 = w * w / 2; c.c3 = w * w * w / 6; c.c4 = w * w * w * w / 24; c.c5 = w * w * w * w * w / 120;
#else
  c.c1 = 1.0; c.c2 = 0.5; c.c3 = 1.6666666666666666e-01; c.c4 = 4.1666666666666664e-02; c.c5 = 8.3333333333333332e-03;
#endif
  c.rc2_u = a.rc2_u;                                 // (rcut/u)^2, or 2^126 = (L/2)^2 exactly: formed on the host (no sqrt /
  c.out_scale = epot ? a.inv_u : a.inv_u * a.inv_u;  // division in every thread's prologue)
#if MDQT_RSQRT_ORDER == 2
  c.out_scale *= epot ? s2 : 2.0 * s2;               // u = ef / (sqrt2 r) sqrt2;  f = ef (1/r + kappa) / r^2 = 2 sqrt2 ef y^2 (y + kappa/sqrt2)
#endif
  return c;
}

// From r2 (in units u^2): rinv = 1/r, ef = exp(-kappa r), valid = 0 < r2 < rcut^2.
template <bool HL, bool TAB32 = false>
__device__ __forceinline__ void pair_core(double r2, const PairConsts& c, const double* tab, double& rinv, double& ef,
                                          bool& valid) {
  if (HL) valid = (unsigned)(__double2hiint(r2) - 1) < 0x47D00000u - 1u;  // 0 < r2 < 2^126 on the high word alone
  else valid = (r2 < c.rc2_u) && (__double2hiint(r2) != 0);  // one DSETP + one ISETP (r2 > 0 <=> high word != 0; SU:222)
#if MDQT_RSQRT_ORDER == 2
  // Seed w ~ 1/sqrt(2 r2): MUFU.RSQ64H reads the high word only, so doubling its argument is one integer add on that word.
  // With g = r2 w ~ r/sqrt2 the residual h = 1/2 - g w = (1 - 2 r2 w^2)/2 is exact to rounding, and ONE coupled step gives
  // both r/sqrt2 = g (1 + h) and 1/(sqrt2 r) = w (1 + h) to second order (the sqrt2 factors live in the constants).
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(__double2hiint(r2) + 0x00100000, 0)));
  const double g = r2 * y;
  const double h = fma(-g, y, 0.5);
  const double r = fma(g, h, g);
  y = fma(y, h, y);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));  // MUFU.RSQ64H seed
  double t = r2 * y;
  double e = fma(-t, y, 1.0);                 // 1 - r2 y^2
  double p = fma(0.375, e, 0.5);
  double ye = y * e;
  y = fma(ye, p, y);                          // y (1 + e/2 + 3 e^2/8): the scheme of CUDA's own rsqrt(double)
  double r = r2 * y;
#endif
  // exp(x), x = -kappa r <= 0:  x = (128 q + idx) ln2/128 + rr
  double tt = fma(r, c.nk_scale, MDQT_MAGIC);
  int n = __double2loint(tt);
  double nd = tt - MDQT_MAGIC;
#if MDQT_EXP_SCALED
  double rr = fma(r, c.nk_scale, -nd);        // exact: the reduced argument in units of ln2/T, |rr| <= 1/2
#else
  double x = c.negkappa_u * r;
  double rr = fma(nd, c.negc, x);
#endif
  double q;
  if (kExpTable == 1024) {
    q = fma(rr, c.c3, c.c2);                  // |x| <= ln2/2048: x^4/24 < 6e-16
  } else {
    q = fma(rr, c.c5, c.c4);
    q = fma(q, rr, c.c3);
    q = fma(q, rr, c.c2);
  }
  q = fma(q, rr, c.c1);
  q = fma(q, rr, 1.0);
  double T;
  if (TAB32) {
    // the item kernel keeps the table 8 KB-aligned in shared memory and passes its 32-bit shared address: base | offset is ONE
    // three-input logic instruction (inside its per-warp item loop the compiler keeps the base in a vector register and would add it)
    const unsigned addr = (((unsigned)n << 3) & ((kExpTable - 1) << 3)) | (unsigned)(size_t)tab;
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(addr));
  } else {
    T = tab[n & (kExpTable - 1)];
  }
  int hi = __double2hiint(T) + (n << kExpShift);  // table high words carry -(idx << kExpShift): net += floor(n/T) << 20
  ef = __hiloint2double(hi, __double2loint(T)) * q;
  rinv = y;
}

// F += f * delta for a pair inside the cut-off (0 < r2 < rc2 as ONE predicate: r2 < 2^126 on the high word alone when HL, else
// one DSETP; the self pair needs none, see k_pairs).
template <bool HL>
__device__ __forceinline__ void accumulate_if_inside(double& ax, double& ay, double& az, double f, double dx, double dy, double dz,
                                                     double r2, double rc2) {
#if MDQT_PRED_ACC
  // the predicate guards the three FMAs (predicated DFMA): no select on the 64-bit f
  const bool inside = HL ? ((unsigned)__double2hiint(r2) < 0x47D00000u) : (r2 < rc2);
  if (inside) { ax = fma(f, dx, ax); ay = fma(f, dy, ay); az = fma(f, dz, az); }
#else
  if (HL)
    asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %1, 0x47D00000;\n\tselp.f64 %0, %0, 0d0000000000000000, q;\n\t}"
        : "+d"(f) : "r"(__double2hiint(r2)));
  else
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %0, 0d0000000000000000, p;\n\t}"
        : "+d"(f) : "d"(r2), "d"(rc2));
  ax = fma(f, dx, ax); ay = fma(f, dy, ay); az = fma(f, dz, az);
#endif
}

constexpr int kTJ = 512;  // j positions staged per pass

// Graph replay: a force launch sits between the substep kernels of two MD steps, so nobody reads the device clock while
// it runs; one thread adds the substeps of the previous MD step (t by the reference's repeated addition, SU:716). Called
// AFTER griddepcontrol.wait, so that with programmatic dependent launch the previous substep kernel has finished reading.
__device__ __forceinline__ void advance_clock(const ForceArgs& a) {
  if (a.clock_advance <= 0) return;
  double t = a.clock[0];
  for (int k = 0; k < a.clock_advance; k++) t = __dadd_rn(t, a.clock_dtq);
  a.clock[0] = t;
  unsigned long long* sub = reinterpret_cast<unsigned long long*>(a.clock + 1);
  *sub += (unsigned long long)a.clock_advance;
}

// 8-byte asynchronous global -> shared copies (LDGSTS): the exp table, the first j tile and the thread's own rows are all
// in flight together in the CTA prologue (one L2 round trip instead of three), and tiles never pass through registers
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifdef MDQT_K1_TRACE  // developer build: per-CTA phase time stamps (scripts/k1_trace.py)
__device__ long long g_trace[8 * 8192];
#define TRACE(slot)                                                                                              \
  if (threadIdx.x == 0) {                                                                                        \
    const int cta_ = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;                             \
    if (cta_ < 8192) {                                                                                           \
      long long gt_;                                                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                                    \
      g_trace[cta_ * 8 + slot] = gt_;                                                                            \
      if (slot == 0) { unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_)); g_trace[cta_ * 8 + 7] = sm_; }      \
    }                                                                                                            \
  }
#else
#define TRACE(slot)
#endif

// RG = ion rows per thread group (128, or 32 for small systems); IPT rows per thread; JS = intra-CTA split of every staged
// j tile over JS groups of RG threads (more resident warps at small N WITHOUT more partial sums in global memory: the
// groups' sums are combined through shared memory in ascending group order); UNR = unroll of the j loop.
// Small N wants RG = 32, JS = 8: the same 256-thread CTAs and the same work per warp as RG = 128, JS = 2, but a row's
// force is split over 4x fewer CTAs, so the final cross-CTA reduction (a chain of dependent L2 loads executed by the
// last CTA of a tile while the rest of the chip idles) shrinks from 20 partials to 5 at N = 3500.
// Register budget (A/B'd on B200, profiles/r01c_k1_trace.txt): the pair loop likes registers more than resident warps. With an
// explicit minimum of 4 CTAs/SM the 128-thread large-N kernel may use up to 128 registers (16 warps/SM): 5.20e11 pairs/s at
// N = 1e5 against 5.03-5.07e11 when capped at 80 or 64 registers (24-32 warps/SM). The 256-thread small-N kernel is best uncapped.
#ifndef MDQT_K1_MINB32
#define MDQT_K1_MINB32 1
#endif
#ifndef MDQT_K1_MINB128
#define MDQT_K1_MINB128 4
#endif
// CL = the j chunks of a row tile form one thread-block CLUSTER (gridDim.y = cluster size <= 8): the chunk sums are
// combined through distributed shared memory in ascending chunk order -- the same order, hence the same bits, as the
// global-memory path -- without partial-sum stores, fence, arrival counter and the last CTA's L2 round trips.
template <int IPT, int JS, bool EPOT, int UNR, bool HL, int RG, bool CL = false>
__global__ void __launch_bounds__(RG * JS, (RG == 32 && JS == 8) ? MDQT_K1_MINB32 : (RG == 128 && JS == 1) ? MDQT_K1_MINB128 : 1) k_pairs(ForceArgs a, double* __restrict__ block_partials) {
  constexpr int kForceThreads = RG;  // shadows the namespace constant: rows per group in this instantiation
#ifndef MDQT_K1_TJ32
#define MDQT_K1_TJ32 1024
#endif
  // small systems: a CTA's whole j chunk (<= 1024 at the plans chosen there) is staged in ONE pass -- the L2 round trip of a
  // second pass would be fully exposed (no other work to hide it behind): -0.7 us at N = 3500
  constexpr int kTJ = (RG == 32) ? MDQT_K1_TJ32 : mdqt::kTJ;
  __shared__ longlong2 sxy[kTJ];
  __shared__ long long sz[kTJ];
  __shared__ double stab[kExpTable];
  __shared__ double sred[kForceThreads * JS / 32];
  __shared__ double sjs[(JS > 1 ? JS - 1 : 1) * IPT * 3 * kForceThreads];
  __shared__ int s_last;
  constexpr int NT = kForceThreads * JS;

  TRACE(0)
  if (threadIdx.x == 0) stamp_time(a.stamp, 0);
  const int tid = threadIdx.x;
  const int ti = tid % kForceThreads, jh = tid / kForceThreads;
  const int b = blockIdx.z, tile = blockIdx.x;
  int js = blockIdx.y + a.js0;  // the j chunk this CTA sums (launches over a subset of the chunks: see ForceArgs)
  if (a.js_skipn && js >= a.js_skip0) js += a.js_skipn;
  const PairConsts c = make_consts(a, EPOT);
  const long long* __restrict__ X = a.Rfix + (size_t)b * 3 * a.ld;
  const long long* __restrict__ Y = X + a.ld;
  const long long* __restrict__ Z = Y + a.ld;
  for (int k = tid; k < kExpTable; k += NT) cp_async8(&stab[k], &c_exp2tab[k]);
  pdl_wait();  // positions (Rfix) come from the previous kernel in the stream
  if (!EPOT && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) advance_clock(a);

  int irow[IPT];
  long long xi[IPT], yi[IPT], zi[IPT];
  double ax[IPT], ay[IPT], az[IPT];
#pragma unroll
  for (int k = 0; k < IPT; k++) {
    irow[k] = a.row0 + tile * (kForceThreads * IPT) + k * kForceThreads + ti;
    const int ir = min(irow[k], a.row0 + a.nrows - 1);  // idle threads shadow the last row (never stored)
    xi[k] = X[ir]; yi[k] = Y[ir]; zi[k] = Z[ir];
    ax[k] = ay[k] = az[k] = 0.0;
  }
  const int jbeg = js * a.jlen;
  const int jend = min(a.N, jbeg + a.jlen);
  for (int jc = jbeg; jc < jend; jc += kTJ) {
    const int cnt = min(kTJ, jend - jc);
    __syncthreads();
    for (int k = tid; k < cnt; k += NT) {
      const int j = jc + k;
      cp_async8(&sxy[k].x, X + j); cp_async8(&sxy[k].y, Y + j); cp_async8(&sz[k], Z + j);
    }
    cp_async_wait_all();  // also covers the exp table on the first pass
    __syncthreads();
    if (jc == jbeg) { TRACE(1) }
    const int lo = (cnt * jh) / JS, hi = (cnt * (jh + 1)) / JS;
#pragma unroll UNR
    for (int jj = lo; jj < hi; jj++) {
      const longlong2 pxy = sxy[jj];
      const long long pz = sz[jj];
#pragma unroll
      for (int k = 0; k < IPT; k++) {
        // two's-complement wrap-around == minimum image; exact difference, one rounding in the conversion
        const double dx = __ll2double_rn((long long)((unsigned long long)xi[k] - (unsigned long long)pxy.x));
        const double dy = __ll2double_rn((long long)((unsigned long long)yi[k] - (unsigned long long)pxy.y));
        const double dz = __ll2double_rn((long long)((unsigned long long)zi[k] - (unsigned long long)pz));
        // Force path: + 1 (one fixed-point unit squared, L^2 2^-128) keeps the self / coincident pair finite -- its
        // f is then multiplied by dx = dy = dz = 0, contributing exactly 0 like the reference's `dr > 0` test
        // (SU:222) -- without a separate r2 != 0 predicate; for any physical separation (r > L 2^-37) the 1 is below
        // the rounding of r2 and changes nothing. The potential-energy path keeps the explicit test.
        const double r2 = EPOT ? fma(dx, dx, fma(dy, dy, dz * dz)) : fma(dx, dx, fma(dy, dy, fma(dz, dz, 1.0)));
        double rinv, ef;
        bool valid;
        pair_core<HL>(r2, c, stab, rinv, ef, valid);
        if (EPOT) {
          double u = ef * rinv;                       // exp(-r/lDeb)/r (SU:268)
          ax[k] += valid ? u : 0.0;
        } else {
          double f = (ef * (rinv * rinv)) * (rinv + c.kappa_u);  // (1/r + 1/lDeb) exp(-r/lDeb)/r^2 (SU:224)
          // 0 < r2 < rc2 as ONE predicate (DSETP, then ISETP chained with .and) and one select
          accumulate_if_inside<HL>(ax[k], ay[k], az[k], f, dx, dy, dz, r2, c.rc2_u);
        }
      }
    }
  }
  TRACE(2)
  // programmatic dependent launch: the next kernel may be scheduled now, so that its launch latency and prologue
  // overlap this kernel's epilogue (triggering at kernel START instead piles up waiting grids and was slower)
  pdl_launch_dependents();
  if (JS > 1) {  // combine the thread groups: group 0 adds the others' sums in ascending group order
    if (jh > 0) {
#pragma unroll
      for (int k = 0; k < IPT; k++) {
        double* o = sjs + (((jh - 1) * IPT + k) * 3) * kForceThreads + ti;
        o[0] = ax[k]; o[kForceThreads] = ay[k]; o[2 * kForceThreads] = az[k];
      }
    }
    __syncthreads();
    if (jh == 0) {
#pragma unroll
      for (int g = 1; g < JS; g++)
#pragma unroll
        for (int k = 0; k < IPT; k++) {
          const double* o = sjs + (((g - 1) * IPT + k) * 3) * kForceThreads + ti;
          ax[k] += o[0]; ay[k] += o[kForceThreads]; az[k] += o[2 * kForceThreads];
        }
    }
  }
  const bool owner = jh == 0;  // only group 0 holds complete sums
#pragma unroll
  for (int k = 0; k < IPT; k++) { ax[k] *= c.out_scale; ay[k] *= c.out_scale; az[k] *= c.out_scale; }

  if (EPOT) {
    // fixed-order block reduction -> one partial per CTA
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; k++) s += (owner && irow[k] < a.row0 + a.nrows) ? ax[k] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) sred[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kForceThreads / 32; w++) tot += sred[w];  // the warps of group 0
      block_partials[((size_t)b * gridDim.y + js) * gridDim.x + tile] = tot;
    }
    if (tid == 0) stamp_time(a.stamp, 1);
    return;
  }

  if (a.nsplit == 1) {
#pragma unroll
    for (int k = 0; k < IPT; k++)
      if (owner && irow[k] < a.row0 + a.nrows) {
        double* Fb = a.F + (size_t)b * 3 * a.ld;
        Fb[irow[k]] = ax[k]; Fb[a.ld + irow[k]] = ay[k]; Fb[2 * a.ld + irow[k]] = az[k];
      }
    if (tid == 0) stamp_time(a.stamp, 1);
    return;
  }
  if (CL) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    double* spart = sjs;  // the in-CTA combine is over (sjs needs 3*(JS-1)*IPT*RG >= 3*IPT*RG doubles: JS >= 2 here)
    __syncthreads();
    if (owner) {
#pragma unroll
      for (int k = 0; k < IPT; k++) {
        spart[(k * 3 + 0) * kForceThreads + ti] = ax[k]; spart[(k * 3 + 1) * kForceThreads + ti] = ay[k];
        spart[(k * 3 + 2) * kForceThreads + ti] = az[k];
      }
    }
    cluster.sync();
    if (js == 0) {  // cluster rank 0 = chunk 0 (cluster dims (1, nsplit, 1)) adds the chunks in ascending order
      for (int w = tid; w < kForceThreads * IPT * 3; w += NT) {
        const int kc = w / kForceThreads, slot = w % kForceThreads;  // kc = k*3 + comp
        const int row = a.row0 + tile * (kForceThreads * IPT) + (kc / 3) * kForceThreads + slot;
        double v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = (r < a.nsplit) ? *cluster.map_shared_rank(&spart[w], r) : 0.0;
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < 8; r++) sum += v[r];
        if (row < a.row0 + a.nrows) a.F[((size_t)b * 3 + (kc % 3)) * a.ld + row] = sum;
      }
    }
    cluster.sync();  // nobody leaves while its shared memory may still be read
    if (tid == 0) stamp_time(a.stamp, 1);
    return;
  }
  // j-split: store the partial, the last CTA of this (trajectory, i-tile) sums all partials in ascending split
  // order -> deterministic, independent of arrival order and of how many ranks share the rows.
#pragma unroll
  for (int k = 0; k < IPT; k++)
    if (owner && irow[k] < a.row0 + a.nrows) {
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
  TRACE(3)
  if (!s_last) { if (tid == 0) stamp_time(a.stamp, 1); return; }
  __threadfence();
  // the last CTA: all NT threads share the final reduction (thread -> (row slot, component) round robin)
  for (int w = tid; w < kForceThreads * IPT * 3; w += NT) {
    const int comp = w / (kForceThreads * IPT), slot = w % (kForceThreads * IPT);
    const int row = a.row0 + tile * (kForceThreads * IPT) + slot;
    if (row < a.row0 + a.nrows) {
      const double* src = a.Fpart + ((size_t)b * 3 + comp) * a.ld + row;
      const size_t stride = (size_t)a.B * 3 * a.ld;
      // the partials of a row are fetched as batches of 8 independent loads (one L2 round trip per batch; the
      // 4-per-batch version cost 4 us at N = 3500 with 20 splits, profiles/), then added in ascending split order
      double sum = 0.0;
      for (int s0 = 0; s0 < a.nsplit; s0 += 8) {  // (wider batches raise the register count of the WHOLE kernel)
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (s0 + k < a.nsplit) ? __ldcg(src + (size_t)(s0 + k) * stride) : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += v[k];  // + 0.0 for the absent splits: exact
      }
      a.F[((size_t)b * 3 + comp) * a.ld + row] = sum;
    }
  }
  TRACE(4)
  if (tid == 0) stamp_time(a.stamp, 1);
}


// ------------------------------------------------------------------------------------------------------------
// K1, item-walking form (small and medium systems, any number of batched trajectories).
//
// A few thousand ions are only ~24 us of issue work for the whole chip, and a grid of CTA tiles loses a fifth of that to
// wave quantisation (1.49 waves at N = 3500), CTA prologues and the cross-CTA reduction of partial sums (fence, arrival
// counter and dependent L2 round trips while the chip idles). Here the grid is PERSISTENT -- two 8-warp CTAs per SM -- and
// the unit of work is an ITEM = (trajectory b, group g of 32 x IPT rows, chunk c of jlen positions), handled by ONE warp:
// lane <-> row(s), the chunk staged by the warp itself into its private shared-memory buffer (cp.async, double buffered:
// the next item's tile and rows are in flight while this one computes), no CTA-wide barrier after the prologue. Items are
// dealt round-robin to the grid's warps; jlen is chosen so that the item count fills all warps evenly when one trajectory
// runs alone (N = 3500: 110 groups x 21 chunks of 168 = 2310 items on 2368 warps, one each); a batch walks B times as many
// items, two rows per lane (fewer shared-memory reads and index instructions per pair).
// Every item writes ONE partial sum and the kernel ends there: no fence, no counter, no reduction. The consumer adds the
// partials of a row in ascending chunk order -- the substep kernel while it loads its ion (mdqt_qt.cu), k_sum_partials for
// everybody else. The summation order of a row is therefore a function of the trajectory's chunk length alone -- which follows
// from the trajectory's OWN ion count (jl[b] = plan_items_jlen(nb[b]); or one nominal length, mdqt_params.plan_n > 0) -- not of
// the batch size, of the rows per lane, of the trajectory's position in the batch, or of the number of row-owning ranks:
// a job gives the same bits alone or batched. Trajectories of an ensemble hold different ion counts nb[b] (SU:299-337).
// ------------------------------------------------------------------------------------------------------------
// MDQT_TAB32 1: 8 KB-aligned table addressed as base | offset in one logic instruction (saves the base add the compiler emits inside
// the per-warp item loop: 44.6 -> 43.6 instructions per pair) -- measured on B200: no gain (28.98 vs 28.66 us at N = 3500), so off
#ifndef MDQT_TAB32
#define MDQT_TAB32 0
#endif
// 16 resident warps per SM either way: two CTAs of 8 warps, or -- for batches (two rows per lane), where it is 4 % faster
// (5.08e11 vs 4.88e11 pairs/s at 64 x 3500) -- one CTA of 16; the bits do not depend on it
#ifndef MDQT_K1_UNR1
#define MDQT_K1_UNR1 8
#endif
#ifndef MDQT_K1_UNR2
#define MDQT_K1_UNR2 4
#endif
constexpr int kItemUnroll1 = MDQT_K1_UNR1, kItemUnroll2 = MDQT_K1_UNR2;  // j-loop unroll with one / two rows per lane (A/B knobs)
constexpr int kItemResidentWarps = 16;
constexpr int kItemUnrollFat = 8;  // j-loop unroll of the one-CTA-per-SM instantiation (two rows per lane): 12 and 16 measured slower
constexpr int kItemMaxB = 512;   // per-trajectory ion counts cached in shared memory up to this batch size
constexpr int kItemMaxJ = 256;  // positions per chunk: 2 buffers x (24 B x 256 + rows) x 8 warps = 120 KB per CTA at most

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct Item { int b, g, ch, Nb, jl; };

#ifdef MDQT_K1_TRACE  // developer build: per-WARP phase time stamps of the item kernel (scripts/k1_items_trace.py)
#define WTRACE(slot)                                                                                             \
  if (lane == 0) {                                                                                               \
    const int w_ = blockIdx.x * NW + warp;                                                               \
    if (w_ < 8192) {                                                                                             \
      long long gt_;                                                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                                    \
      g_trace[w_ * 8 + slot] = gt_;                                                                              \
      if (slot == 0) { unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_)); g_trace[w_ * 8 + 7] = sm_; }        \
    }                                                                                                            \
  }
#else
#define WTRACE(slot)
#endif

template <int NW, int IPT, bool EPOT, bool HL, int RW = kItemResidentWarps>  // RW = resident warps per SM (register budget 65536 / (32 RW))
__global__ void __launch_bounds__(NW * 32, RW / NW) k_pairs_items(ForceArgs a, double* __restrict__ item_partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8192) double stab_mem[kExpTable];  // 8 KB-aligned: see pair_core<.., TAB32>
  __shared__ int snb[kItemMaxB];  // the trajectories' ion counts: read at every item decode, so not from L2
  __shared__ int sjl[kItemMaxB];  // and their chunk lengths (a function of each trajectory's own ion count: plan_items_jlen)
  constexpr int RPG = 32 * IPT;  // rows per group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tj = a.jlen;  // tile capacity (a multiple of 8, <= kItemMaxJ)
  // per warp and buffer: xy[tj] (16 B), z[tj] (8 B), then the item's rows x, y, z (3 x RPG x 8 B)
  const unsigned bufbytes = 24u * (unsigned)tj + 24u * RPG;
  unsigned char* wbase = smem_raw + (size_t)warp * (2 * (size_t)bufbytes);
  const PairConsts c = make_consts(a, EPOT);
  const double* const stab = MDQT_TAB32 ? reinterpret_cast<const double*>((size_t)(unsigned)__cvta_generic_to_shared(stab_mem)) : stab_mem;
  if (tid == 0) stamp_time(a.stamp, 0);
  WTRACE(0)
  for (int k = tid; k < kExpTable; k += NW * 32) cp_async8(&stab_mem[k], &c_exp2tab[k]);
  const bool nb_smem = a.nb && a.B <= kItemMaxB;
  if (nb_smem) for (int k = tid; k < a.B; k += NW * 32) { snb[k] = a.nb[k]; sjl[k] = a.jl ? a.jl[k] : a.jlen; }  // constant for the handle's lifetime
  __syncthreads();
  pdl_wait();  // positions (Rfix) come from the previous kernel in the stream
  if (!EPOT && tid == 0 && blockIdx.x == 0) advance_clock(a);

  const int gcap = (a.nrows + RPG - 1) / RPG;
  const int W = gridDim.x * NW;
  const bool listed = !EPOT && IPT == 2 && a.ilist != nullptr;  // walk the packed list of non-empty items instead of all slots
  const int total = listed ? a.icount : a.B * gcap * a.nsplit;
  const int rowend_cap = a.row0 + a.nrows;
  const unsigned long long mg_g = (IPT == 2) ? a.mg_gcap2 : a.mg_gcap;
  // item k -> (b, g, ch) by multiplication with host-made reciprocals (k < 2^24, divisors < 2^16: exact)
  auto decode = [&](int k, Item& it) {
    if (listed) {
      const unsigned v = __ldg(a.ilist + k);
      it.b = (int)(v >> 18); it.g = (int)((v >> 10) & 255u); it.ch = (int)(v & 1023u);
    } else {
      const unsigned t = (unsigned)(((unsigned long long)(unsigned)k * a.mg_chunk) >> 40);
      it.ch = k - (int)t * a.nsplit;
      it.b = (int)(((unsigned long long)t * mg_g) >> 40);
      it.g = (int)t - it.b * gcap;
    }
    it.Nb = nb_smem ? snb[it.b] : (a.nb ? a.nb[it.b] : a.N);
    it.jl = nb_smem ? sjl[it.b] : (a.jl ? a.jl[it.b] : a.jlen);
    // empty when the group or the chunk lies beyond the trajectory's ions
    return (a.row0 + it.g * RPG < min(rowend_cap, it.Nb)) && (it.ch * it.jl < it.Nb);
  };
  auto next_valid = [&](int k, Item& it) {
    while (k < total && !decode(k, it)) {
      if (EPOT && lane == 0) item_partials[k] = 0.0;  // the per-trajectory sum runs over all item slots
      k += W;
    }
    return k;
  };
  auto issue = [&](const Item& it, int buf) {
    const long long* __restrict__ X = a.Rfix + (size_t)it.b * 3 * a.ld;
    const long long* __restrict__ Y = X + a.ld;
    const long long* __restrict__ Z = Y + a.ld;
    unsigned char* base = wbase + (size_t)buf * bufbytes;
    longlong2* sxy = reinterpret_cast<longlong2*>(base);
    long long* sz = reinterpret_cast<long long*>(base + 16u * (unsigned)tj);
    long long* srow = reinterpret_cast<long long*>(base + 24u * (unsigned)tj);
    const int jbeg = it.ch * it.jl, cnt = min(it.jl, it.Nb - jbeg);
    for (int q = lane; q < cnt; q += 32) {
      cp_async8(&sxy[q].x, X + jbeg + q); cp_async8(&sxy[q].y, Y + jbeg + q); cp_async8(&sz[q], Z + jbeg + q);
    }
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      const int ir = min(a.row0 + it.g * RPG + r * 32 + lane, min(rowend_cap, it.Nb) - 1);  // idle lanes shadow the last row (never stored)
      cp_async8(&srow[r * 32 + lane], X + ir); cp_async8(&srow[RPG + r * 32 + lane], Y + ir); cp_async8(&srow[2 * RPG + r * 32 + lane], Z + ir);
    }
  };

  Item cur, nxt;
  int k = next_valid(warp * gridDim.x + blockIdx.x, cur);  // consecutive items go to different SMs
  if (k < total) issue(cur, 0);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();  // the exp table is complete for every warp; the last CTA-wide barrier of the kernel
  WTRACE(1)
  int buf = 0;
  while (k < total) {
    const int kn = next_valid(k + W, nxt);
    __syncwarp();  // every lane is done reading the buffer that the next tile overwrites
    if (kn < total) issue(nxt, buf ^ 1);
    cp_async_commit();
    cp_async_wait_1();  // all but the newest group: this item's tile has landed
    __syncwarp();
    const unsigned char* base = wbase + (size_t)buf * bufbytes;
    const longlong2* __restrict__ sxy = reinterpret_cast<const longlong2*>(base);
    const long long* __restrict__ sz = reinterpret_cast<const long long*>(base + 16u * (unsigned)tj);
    const long long* __restrict__ srow = reinterpret_cast<const long long*>(base + 24u * (unsigned)tj);
    long long xi[IPT], yi[IPT], zi[IPT];
    double ax[IPT], ay[IPT], az[IPT];
#pragma unroll
    for (int r = 0; r < IPT; r++) {
      xi[r] = srow[r * 32 + lane]; yi[r] = srow[RPG + r * 32 + lane]; zi[r] = srow[2 * RPG + r * 32 + lane];
      ax[r] = ay[r] = az[r] = 0.0;
    }
    const int cnt = min(cur.jl, cur.Nb - cur.ch * cur.jl);
    WTRACE(2)
#pragma unroll(IPT == 2 ? (RW == 8 ? kItemUnrollFat : kItemUnroll2) : kItemUnroll1)
    for (int jj = 0; jj < cnt; jj++) {
      const longlong2 pxy = sxy[jj];
      const long long pz = sz[jj];
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const double dx = __ll2double_rn((long long)((unsigned long long)xi[r] - (unsigned long long)pxy.x));
        const double dy = __ll2double_rn((long long)((unsigned long long)yi[r] - (unsigned long long)pxy.y));
        const double dz = __ll2double_rn((long long)((unsigned long long)zi[r] - (unsigned long long)pz));
        const double r2 = EPOT ? fma(dx, dx, fma(dy, dy, dz * dz)) : fma(dx, dx, fma(dy, dy, fma(dz, dz, 1.0)));  // see k_pairs
        double rinv, ef;
        bool valid;
        pair_core<HL, MDQT_TAB32 != 0>(r2, c, stab, rinv, ef, valid);
        if (EPOT) {
          const double u = ef * rinv;
          ax[r] += valid ? u : 0.0;
        } else {
          double f = (ef * (rinv * rinv)) * (rinv + c.kappa_u);
          accumulate_if_inside<HL>(ax[r], ay[r], az[r], f, dx, dy, dz, r2, c.rc2_u);
        }
      }
    }
    WTRACE(3)
    if (EPOT) {
      static_assert(!EPOT || IPT == 1, "the potential-energy instantiation keeps one row per lane");
      const int row = a.row0 + cur.g * RPG + lane;
      double sum = (row < min(rowend_cap, cur.Nb)) ? ax[0] * c.out_scale : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
      if (lane == 0) item_partials[k] = sum;
    } else {
      // one partial per item -- or the force itself when the whole j range is a single chunk
      double* dst = (a.nsplit == 1 ? a.F : a.Fpart + (size_t)cur.ch * ((size_t)a.B * 3 * a.ld)) + (size_t)cur.b * 3 * a.ld;
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const int row = a.row0 + cur.g * RPG + r * 32 + lane;
        if (row < min(rowend_cap, cur.Nb)) {
          dst[row] = ax[r] * c.out_scale; dst[a.ld + row] = ay[r] * c.out_scale; dst[2 * a.ld + row] = az[r] * c.out_scale;
        }
      }
    }
    k = kn; buf ^= 1; cur = nxt;
  }
  pdl_launch_dependents();
  if (lane == 0) stamp_time(a.stamp, 1);
  WTRACE(4)
}

// rows per lane of the item kernel: two once every warp has several items anyway (batches), one when a single trajectory
// must be spread over all warps. Any rule is fine -- the summation order, hence the bits, do not depend on it.
static int items_ipt(const ForceArgs& a) {
  static const int force = [] { const char* e = getenv("MDQT_ITEMS_IPT"); return e ? atoi(e) : 0; }();
  if (force == 1 || force == 2) return force;
  const long long items1 = (long long)a.B * ((a.nrows + 31) / 32) * a.nsplit;
  return items1 >= 4LL * 148 * kItemResidentWarps ? 2 : 1;
}

template <int NW, int IPT, bool EPOT, int RW = kItemResidentWarps>
static void launch_items_nw(const ForceArgs& a, double* partials, cudaStream_t s) {
  const size_t smem = (size_t)NW * 2 * (24 * (size_t)a.jlen + 24 * 32 * IPT);
  const long long total = (!EPOT && IPT == 2 && a.ilist) ? a.icount : (long long)a.B * ((a.nrows + 32 * IPT - 1) / (32 * IPT)) * a.nsplit;
  const int grid = (int)std::min<long long>(148LL * (RW / NW), (total + NW - 1) / NW);
  const bool hl = a.half_l && MDQT_VALID_INT;
  void (*kern)(ForceArgs, double*) = hl ? k_pairs_items<NW, IPT, EPOT, true, RW> : k_pairs_items<NW, IPT, EPOT, false, RW>;
  // function attributes are per DEVICE (one process may drive several GPUs from several threads: mdqt_run --gpus): set them once
  // on every device this instantiation is launched on
  static std::atomic<bool> attr_done[64][2];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !attr_done[dev][hl].load(std::memory_order_acquire)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)std::min<size_t>((size_t)NW * 2 * (24 * (size_t)kItemMaxJ + 24 * 32 * IPT), (size_t)(227 - 12) * 1024));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);  // room for all resident CTAs
    if (dev < 64) attr_done[dev][hl].store(true, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NW * 32); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  int n = 0;
  if (!EPOT && a.pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; n++; }
  cfg.attrs = at; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, a, partials);
}

template <bool EPOT>
static void launch_items(const ForceArgs& a, double* partials, cudaStream_t s) {
  const int ipt = EPOT ? 1 : items_ipt(a);  // the potential energy is a diagnostic: one instantiation
  if (EPOT) { launch_items_nw<8, 1, EPOT>(a, partials, s); return; }
  // One trajectory whose items fit ONE round of 148 x 8 warps at two rows per lane: ONE 8-warp CTA per SM with the whole register
  // file (156 registers, 16 independent pair chains per warp) instead of two CTAs of 8 one-row warps. Two co-resident CTAs do not
  // share the issue ports evenly -- the one that arrived first has strict priority, finishes early and leaves the other alone on ports
  // that two 8-chain warps per scheduler fill to 67 % only -- whereas here the two warps of a scheduler are equals
  // and each carries twice the chains: 28.6 -> 27.5 us at N = 3500, 31.0 -> 29.8 at N = 3653 (profiles/r02y_fat_ab2.log). Same items,
  // same chunk order, same bits. MDQT_K1_FAT=0 switches it off (A/B runs).
  {
    static const int fat = [] { const char* e = getenv("MDQT_K1_FAT"); return e ? atoi(e) : 1; }();  // 2: also for batches (A/B: slower)
    const long long items2 = (long long)((a.nrows + 63) / 64) * a.nsplit;
    if (fat == 2 || (fat && a.B == 1 && items2 <= 148LL * 8)) { launch_items_nw<8, 2, false, 8>(a, partials, s); return; }
  }
  if (ipt == 1) { launch_items_nw<8, 1, false>(a, partials, s); return; }
  // two rows per lane (batches): one 16-warp CTA per SM when its tiles fit (chunks of up to 192 positions), else two of 8
  if ((size_t)16 * 2 * (24 * (size_t)a.jlen + 24 * 64) + 12 * 1024 <= (size_t)227 * 1024) launch_items_nw<16, 2, false>(a, partials, s);
  else launch_items_nw<8, 2, false>(a, partials, s);
}

// F[b][comp][row] = sum of the row's partials in ascending chunk order: the one formula every consumer of item-kernel
// partials uses (here, and the substep kernel while it loads its ion), so that F has the same bits whoever adds it up
__global__ void k_sum_partials(ForceArgs a) {
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= (long long)a.B * 3 * a.nrows) return;
  const int i = a.row0 + (int)(gidx % a.nrows);
  const int bc = (int)(gidx / a.nrows), b = bc / 3;
  const int Nb = a.nb ? a.nb[b] : a.N;
  if (i >= Nb) return;
  const int jl = a.jl ? a.jl[b] : a.jlen;
  const int nch = (Nb + jl - 1) / jl;
  const size_t stride = (size_t)a.B * 3 * a.ld;
  const double* src = a.Fpart + (size_t)bc * a.ld + i;
  a.F[(size_t)bc * a.ld + i] = sum_partials(src, stride, nch);
}

template <bool EPOT, bool HL>
static void launch_pairs_hl(const ForceArgs& a, double* partials, cudaStream_t s, dim3 grid, int ipt, int jsub, bool pdl) {
  // few resident warps (small N): also unroll the j loop further so that one warp carries more independent pairs
  if (a.rg == 32) {
    // clusters must be co-scheduled inside one GPC: a win (-2 us at N = 2048) only while every CTA of the grid is resident
    // at once (two 256-thread or four 128-thread CTAs per SM at ~124 registers); beyond that the placement constraint costs
    // more balance than the reduction saves (N = 3500: 41.5 us against 32.8 us), so larger grids keep the global-memory path
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    const bool cl = !EPOT && a.js_count == 0 && a.nsplit >= 2 && a.nsplit <= 8 && ctas <= 148LL * (jsub == 8 ? 2 : 4) && cluster_enabled();
    if (cl) {
      if (jsub == 8) launch_kernel(k_pairs<1, 8, false, 8, HL, 32, true>, grid, dim3(256), s, pdl, a, partials, a.nsplit);
      else launch_kernel(k_pairs<1, 4, false, 8, HL, 32, true>, grid, dim3(128), s, pdl, a, partials, a.nsplit);
      return;
    }
    if (jsub == 8) launch_kernel(k_pairs<1, 8, EPOT, 8, HL, 32>, grid, dim3(256), s, pdl, a, partials);
    else launch_kernel(k_pairs<1, 4, EPOT, 8, HL, 32>, grid, dim3(128), s, pdl, a, partials);
    return;
  }
  if (ipt == 2 && jsub == 4) launch_kernel(k_pairs<2, 4, EPOT, 4, HL, 128>, grid, dim3(kForceThreads * 4), s, pdl, a, partials);
  else if (ipt == 2 && jsub == 2) launch_kernel(k_pairs<2, 2, EPOT, 4, HL, 128>, grid, dim3(kForceThreads * 2), s, pdl, a, partials);
  else if (ipt == 2) launch_kernel(k_pairs<2, 1, EPOT, 4, HL, 128>, grid, dim3(kForceThreads), s, pdl, a, partials);
  else if (jsub == 2) launch_kernel(k_pairs<1, 2, EPOT, 8, HL, 128>, grid, dim3(kForceThreads * 2), s, pdl, a, partials);
  else launch_kernel(k_pairs<1, 1, EPOT, 4, HL, 128>, grid, dim3(kForceThreads), s, pdl, a, partials);
}

template <bool EPOT>
static void launch_pairs(const ForceArgs& a, double* partials, cudaStream_t s) {
  if (a.items) { launch_items<EPOT>(a, partials, s); return; }
  // rows per group / per thread and the intra-CTA split: decided by the planner from (N, B) only
  const int rg = a.rg == 32 ? 32 : kForceThreads;
  const int ipt = (rg == 32) ? 1 : (a.ipt == 2 ? 2 : 1);
  const int jsub = (rg == 32) ? (a.jsub == 8 ? 8 : 4) : ((a.jsub == 2 || a.jsub == 4) ? a.jsub : 1);
  dim3 grid((a.nrows + rg * ipt - 1) / (rg * ipt), a.js_count > 0 ? a.js_count : a.nsplit, a.B);
  const bool pdl = !EPOT && a.pdl;
  if (a.half_l && MDQT_VALID_INT) launch_pairs_hl<EPOT, true>(a, partials, s, grid, ipt, jsub, pdl);
  else launch_pairs_hl<EPOT, false>(a, partials, s, grid, ipt, jsub, pdl);
}

// finalize: F must hold the complete forces on return (false: the caller's next kernel adds the item kernel's partials itself)
void launch_forces(const ForceArgs& a, cudaStream_t s, bool finalize) {
  launch_pairs<false>(a, nullptr, s);
  if (finalize && forces_are_partial(a)) {
    const long long n = (long long)a.B * 3 * a.nrows;
    k_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
  }
}

#ifdef MDQT_K1_TRACE
extern "C" int mdqt_debug_read_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * (size_t)n);
}
#endif

__global__ void k_to_fixed(const double* __restrict__ R, long long* __restrict__ Rfix, size_t n, double invL, double invL_lo) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) Rfix[k] = to_fixed(R[k], invL, invL_lo);
}
void launch_to_fixed(const double* R, long long* Rfix, size_t n, double invL, double invL_lo, cudaStream_t s) {
  k_to_fixed<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(R, Rfix, n, invL, invL_lo);
}

int epot_partials_needed(const ForceArgs& a) {
  int tiles = (a.nrows + 31) / 32;  // upper bound (32-row groups, IPT = 1)
  return tiles * a.nsplit * a.B;
}

__global__ void k_epot_final(const double* __restrict__ partials, int per_traj, double half, int N, const int* __restrict__ nb,
                             double* __restrict__ out) {
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
  if (threadIdx.x == 0) out[blockIdx.x] = sred[0] * (half / (double)(nb ? nb[blockIdx.x] : N));
}

// item kernel: partials[(b * gcap + g) * nchunk + c]. Thread t adds the groups t, t + 256, ... (each group: its chunks in
// ascending order), then the fixed tree: groups and chunks beyond a trajectory's ions hold 0 and are added last in every
// thread's sequence, so the result does not depend on the handle's capacity -- a job's energy has the same bits in any batch
__global__ void k_epot_final_items(const double* __restrict__ partials, int gcap, int nchunk, double half, int N,
                                   const int* __restrict__ nb, double* __restrict__ out) {
  __shared__ double sred[256];
  const double* p = partials + (size_t)blockIdx.x * gcap * nchunk;
  double s = 0.0;
  for (int g = threadIdx.x; g < gcap; g += 256) {
    double sg = 0.0;
    for (int c = 0; c < nchunk; c++) sg += p[(size_t)g * nchunk + c];
    s += sg;
  }
  sred[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sred[0] * (half / (double)(nb ? nb[blockIdx.x] : N));
}

void launch_epot(const ForceArgs& a, double* partials, double* result, cudaStream_t s) {
  launch_pairs<true>(a, partials, s);
  int per_traj;
  if (a.items) {
    k_epot_final_items<<<a.B, 256, 0, s>>>(partials, (a.nrows + 31) / 32, a.nsplit, 0.5, a.N, a.nb, result);
    return;
  } else {
    const int rg = a.rg == 32 ? 32 : kForceThreads;
    const int ipt = (rg == 32) ? 1 : (a.ipt == 2 ? 2 : 1);
    per_traj = (a.nrows + rg * ipt - 1) / (rg * ipt) * a.nsplit;
  }
  // ordered pairs counted twice -> 1/2; per particle -> 1/N (SU:272)
  k_epot_final<<<a.B, 256, 0, s>>>(partials, per_traj, 0.5, a.N, a.nb, result);
}

// ---- FP64 peak probe: 8 independent DFMA chains per thread, all-register operands ----
__global__ void __launch_bounds__(256) k_dfma_chain(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 16  // 128 DFMA per loop trip: the loop's own 3 instructions cost ~1 % (they cost 5 % at unroll 4)
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
  for (int rep = 0; rep < 6; rep++) {
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

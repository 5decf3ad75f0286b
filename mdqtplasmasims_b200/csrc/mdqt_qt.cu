// mdqt_qt.cu -- K2: the fused per-ion quantum-substep kernel (leap-frog position/velocity update + quantum-
// trajectory step), and K5: the MD-family velocity-Verlet kernels. Hand-written for sm_100a, fp64.
//
// Replaces step() / step_R() / step_V() (reference laserCoolingPlusExpansionMDQTSpeedUp.cpp:356-430) and
// qstep() (SU:438-717 with the operator tables of SU:1163-1215) -- and, as the 7-level instantiation, qstep()
// of MonteCarloFollowedByQTTagging408Linear.cpp:555-756 (tables MC408L:1171-1190; Quad mask MC408Q:596) --
// plus stepPositions()/stepVelocities() of MonteCarloFollowedByMDAndTempAnisotropy.cpp:452-502.
//
// Design. The reference builds dense 12x12 complex Armadillo matrices per ion per substep. The Hamiltonian is in
// fact block diagonal: the sigma+/sigma- light fields (no pi light) only connect
//     block A = { S-1/2, P+1/2, P-3/2, D+3/2, D-1/2, D-5/2 }   (0-indexed states 0,3,5,10,8,6)
//     block B = { S+1/2, P+3/2, P-1/2, D+5/2, D+1/2, D-3/2 }   (0-indexed states 1,2,4,11,9,7)
// and inside a block only 5 real couplings + 1 phase-rotating complex coupling are non-zero. The two blocks talk
// to each other only through the renormalisation prefactor 1/sqrt(1-dp) (a sum of P populations) and through
// quantum jumps. So each ion is advanced by a PAIR of adjacent lanes, one block (6 complex amplitudes, in
// registers) per lane, exchanging one double per Runge-Kutta stage with __shfl_xor. All `nsub` substeps between
// two force evaluations are fused in one launch (F is frozen in between, SU:1369-1378), so R, V, F, psi, tPart
// cross HBM once per MD step instead of once per substep. Random numbers are Philox4x32-10 keyed by
// (seed; substep, ion, trajectory) -- no state, no ordering dependence (the reference's shared drand48 races).
#include "mdqt_internal.h"
#include "mdqt_qtconsts.h"
#include "mdqt_fixed.cuh"
#include <math.h>
#include <stdlib.h>

namespace mdqt {

// ------------------------------------------------------------------------------------------------------------
// Philox4x32-10 and the uniform layout shared with oracle/mdqt_oracle.c (orc_uniforms5 / orc_collision_draws)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ double u52(unsigned hi, unsigned lo) {
  unsigned long long k = ((unsigned long long)hi << 20) | (unsigned long long)(lo >> 12);
  return ((double)k + 0.5) * 2.220446049250313080847263336181640625e-16;  // (k + 1/2) 2^-52, in (0,1)
}
__device__ __forceinline__ uint4 philox_call(uint64_t seed, unsigned traj, unsigned ion, uint64_t step, unsigned call) {
  return philox4x32_10(make_uint4((unsigned)step, (unsigned)(step >> 32), ion, (traj << 3) | call),
                       make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}

struct cplx { double re, im; };
__device__ __forceinline__ double cnorm(const cplx& a) { return fma(a.re, a.re, a.im * a.im); }
// Im(a conj(b))
__device__ __forceinline__ double im_acb(const cplx& a, const cplx& b) { return fma(a.im, b.re, -(a.re * b.im)); }

// ------------------------------------------------------------------------------------------------------------
// One lane's block of the stage map  g(w) = pref * (1 - i h H) w - w   (= h * k_stage of SU:534-535)
// local states: 0 = S, 1 = P1, 2 = P2, 3 = D coupled to P1, 4 = D coupled to P1 (static) and P2 (rotating),
// 5 = D coupled to P2.   NL = 6 (12-level) or 4 (7-level: S, P1, P2 and the uncoupled D reservoir).
// ------------------------------------------------------------------------------------------------------------
struct LaneH {
  double hc10, hc20, hc13, hc14, hc25;  // h * real couplings
  double hrr, hri;                      // h * (rotating coupling H[4][2]) re, im
  double hE1, hE2, hE3, hE4, hE5;       // h * energies
  double hg1, hg2;                      // h * Gamma/2 on P1, P2
  double G1, G2;                        // h * Gamma on P1, P2 (for dp)
};

template <int NL>
__device__ __forceinline__ void apply_M(const LaneH& H, const cplx* w, cplx* m) {
  // m = w - i h (H w):  m.re = w.re + h Im(Hw) ;  m.im = w.im - h Re(Hw). Every row is ONE fma chain that starts from
  // the amplitude itself (no separate Hw, no final add): with z = x + i y and a diagonal E - i g,
  //   m.re = x + hE y - hg x + sum_c hc y_c ,   m.im = y - hE x - hg y - sum_c hc x_c .
  // S
  m[0].re = fma(H.hc10, w[1].im, fma(H.hc20, w[2].im, w[0].re));
  m[0].im = fma(-H.hc10, w[1].re, fma(-H.hc20, w[2].re, w[0].im));
  // P1: (E1 - i g1) w1 + c10 w0 + c13 w3 + c14 w4
  {
    double re = fma(H.hE1, w[1].im, fma(-H.hg1, w[1].re, w[1].re));
    double im = fma(-H.hE1, w[1].re, fma(-H.hg1, w[1].im, w[1].im));
    re = fma(H.hc10, w[0].im, re); im = fma(-H.hc10, w[0].re, im);
    if (NL == 6) {
      re = fma(H.hc13, w[3].im, re); im = fma(-H.hc13, w[3].re, im);
      re = fma(H.hc14, w[4].im, re); im = fma(-H.hc14, w[4].re, im);
    }
    m[1].re = re; m[1].im = im;
  }
  // P2: (E2 - i g2) w2 + c20 w0 + c25 w5 + conj(rot) w4 ; conj(rot) w4 = (rr x4 + ri y4) + i (rr y4 - ri x4)
  {
    double re = fma(H.hE2, w[2].im, fma(-H.hg2, w[2].re, w[2].re));
    double im = fma(-H.hE2, w[2].re, fma(-H.hg2, w[2].im, w[2].im));
    re = fma(H.hc20, w[0].im, re); im = fma(-H.hc20, w[0].re, im);
    if (NL == 6) {
      re = fma(H.hc25, w[5].im, re); im = fma(-H.hc25, w[5].re, im);
      re = fma(H.hrr, w[4].im, fma(-H.hri, w[4].re, re));
      im = fma(-H.hrr, w[4].re, fma(-H.hri, w[4].im, im));
    }
    m[2].re = re; m[2].im = im;
  }
  if (NL == 6) {
    {  // D(3): E3 w3 + c13 w1
      m[3].re = fma(H.hE3, w[3].im, fma(H.hc13, w[1].im, w[3].re));
      m[3].im = fma(-H.hE3, w[3].re, fma(-H.hc13, w[1].re, w[3].im));
    }
    {  // D(4): E4 w4 + c14 w1 + rot w2 ; rot w2 = (rr x2 - ri y2) + i (rr y2 + ri x2)
      double re = fma(H.hE4, w[4].im, fma(H.hc14, w[1].im, w[4].re));
      double im = fma(-H.hE4, w[4].re, fma(-H.hc14, w[1].re, w[4].im));
      re = fma(H.hrr, w[2].im, fma(H.hri, w[2].re, re));
      im = fma(-H.hrr, w[2].re, fma(H.hri, w[2].im, im));
      m[4].re = re; m[4].im = im;
    }
    {  // D(5): E5 w5 + c25 w2
      m[5].re = fma(H.hE5, w[5].im, fma(H.hc25, w[2].im, w[5].re));
      m[5].im = fma(-H.hE5, w[5].re, fma(-H.hc25, w[2].re, w[5].im));
    }
  } else {
    m[3] = w[3];  // 7-level D reservoir: zero energy, no coupling
  }
}

// 1/sqrt(x) for x in (0, 1]: MUFU.RSQ64H seed + one 3rd-order Newton step (the scheme of CUDA's rsqrt(double),
// without its special-case branches: x = 1 - dp is always a normal number close to 1)
__device__ __forceinline__ double rsqrt_near1(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// sin and cos of an arbitrary-magnitude phase (|phi| < 2^50): two-term FMA Cody-Waite reduction by pi/2 and the
// fdlibm minimax kernels on [-pi/4, pi/4]; branch-free (~25 FP64 instructions, < 1 ulp for |phi| < 1e9).
__device__ __forceinline__ void sincos_fast(double phi, double& sn, double& cs) {
  const double t = fma(phi, 0.63661977236758134308, 6755399441055744.0);  // phi * 2/pi, round to nearest
  const int q = __double2loint(t);
  const double n = t - 6755399441055744.0;
  double r = fma(n, -1.5707963267948966192, phi);
  r = fma(n, -6.1232339957367658860e-17, r);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s0 = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double a = (q & 1) ? c0 : s0, b = (q & 1) ? s0 : c0;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}


// Forces of ion i, components 0 and c2: either F itself, or -- when the item force kernel left one partial sum per j chunk --
// the partials added in ascending chunk order (the same formula as k_sum_partials), written back to F by the storing lanes.
// Kept out of line: inlined into the substep kernels it perturbed ptxas' schedule of their latency-bound main loop
// (25 substeps: 23 -> 35 us at N = 3500).
#ifndef MDQT_K2_FLOAD
#define MDQT_K2_FLOAD 1
#endif
#if MDQT_K2_FLOAD == 1
__device__ __noinline__
#else
__device__ __forceinline__
#endif
double2 load_forces_impl(const double* __restrict__ fpart, double* __restrict__ Fw, const double* __restrict__ F, int Nb, int jlen, int B,
                         int ld, int b, int i, int c2, int store_mask) {  // scalars only: a by-reference QTArgs would be copied to the stack
  double2 f;
  if (fpart) {
    const int nch = (Nb + jlen - 1) / jlen;
    const size_t stride = (size_t)B * 3 * ld;
    f.x = sum_partials(fpart + (size_t)b * 3 * ld + i, stride, nch);
    f.y = sum_partials(fpart + ((size_t)b * 3 + c2) * ld + i, stride, nch);
    double* Fo = Fw + (size_t)b * 3 * ld;
    if (store_mask & 2) Fo[(size_t)c2 * ld + i] = f.y;
    if (store_mask & 1) Fo[i] = f.x;
  } else {
    const double* __restrict__ Fb = F + (size_t)b * 3 * ld;
    f.x = Fb[i]; f.y = Fb[(size_t)c2 * ld + i];
  }
  return f;
}
#define load_forces(a, b, i, c2, store_x, store_2, fx, f2)                                                                       \
  do {                                                                                                                           \
    const double2 f_ = load_forces_impl((a).fpart, (a).Fw, (a).F, (a).nb ? (a).nb[b] : (a).N, (a).jl ? (a).jl[b] : (a).fp_jlen, (a).B, (a).ld, b, i,  \
                                        c2, ((store_x) ? 1 : 0) | ((store_2) ? 2 : 0));                                          \
    fx = f_.x; f2 = f_.y;                                                                                                        \
  } while (0)

// One force component of ion i (the four-lane kernel: quad lanes 0, 1, 2 take x, y, z and share them by shuffle, so a lane has
// 21 partial loads in flight in ONE L2 round trip instead of 2 x 21 in four). Same ascending sum, same bits.
#ifndef MDQT_K2_TRIGGER
#define MDQT_K2_TRIGGER 0  // 1: the four-lane substep kernel releases its dependent (the force kernel) at its START (A/B knob)
#endif
#ifndef MDQT_K2_FLOAD1
#define MDQT_K2_FLOAD1 1
#endif
__device__ __noinline__ double load_force1_impl(const double* __restrict__ fpart, double* __restrict__ Fw, const double* __restrict__ F, int Nb,
                                                int jlen, int B, int ld, int b, int i, int comp, int store) {
  const size_t at = ((size_t)b * 3 + comp) * ld + i;
  if (!fpart) return F[at];
  const double f = sum_partials<24>(fpart + at, (size_t)B * 3 * ld, (Nb + jlen - 1) / jlen);
  if (store) Fw[at] = f;
  return f;
}

template <int NL>
__device__ __forceinline__ void stage(const LaneH& H, const cplx* w, cplx* g) {
  cplx m[NL];
  double own = fma(H.G1, cnorm(w[1]), H.G2 * cnorm(w[2]));
  double dp = own + __shfl_xor_sync(0xffffffffu, own, 1);
  double pref = rsqrt_near1(1.0 - dp);
  apply_M<NL>(H, w, m);
#pragma unroll
  for (int k = 0; k < NL; k++) {
    g[k].re = fma(pref, m[k].re, -w[k].re);
    g[k].im = fma(pref, m[k].im, -w[k].im);
  }
}

// Doppler shift, rotating phase and the coefficients that depend on them, for a substep that starts (after step()'s velocity
// update) with velocity vxs, time-since-jump tps (already advanced) and global time ts (SU:447, 481-483, 506-510). The two
// helpers are shared by the two- and the four-lane kernel: same expressions, same roundings.
struct PrepC { double e0_0, e1_0, e0_1, e1_1, e0_2, e1_2, hrot, pv2qv, two_kr, g2E, ed_num, ed_den, ed_c; bool expand; };
struct Pre { double hE0, hE1, hE2, cr, ci; };
__device__ __forceinline__ double prep_uu(const PrepC& c, double vxs, double ts) {
  double expDet = 0.0;
  if (c.expand) expDet = __ddiv_rn(__dmul_rn(c.ed_num, ts), __dmul_rn(c.ed_den, sqrt(fma(__dmul_rn(c.ed_c, ts), ts, 1.0))));
  return fma(vxs, c.pv2qv, expDet);
}
__device__ __forceinline__ void prep_rot(const PrepC& c, double uu, double tps, double& cr, double& ci) {
  const double phi = __dmul_rn(__dmul_rn(__dmul_rn(c.two_kr, uu), tps), c.g2E);  // 2 u (1 + kRat) tPart g2E (SU:508)
  double sn, cs;
  sincos_fast(phi, sn, cs);
  cr = __dmul_rn(c.hrot, cs); ci = __dmul_rn(c.hrot, sn);
}
__device__ __forceinline__ Pre prep4(const PrepC& c, double vxs, double tps, double ts) {
  const double uu = prep_uu(c, vxs, ts);
  Pre p;
  p.hE0 = fma(c.e1_0, uu, c.e0_0); p.hE1 = fma(c.e1_1, uu, c.e0_1); p.hE2 = fma(c.e1_2, uu, c.e0_2);
  prep_rot(c, uu, tps, p.cr, p.ci);
  return p;
}

// The 12-level rows of m = w - i h H w in the two-lane kernel, written with EXACTLY the operations (and their order) that the
// four-lane kernel's generic rows perform on non-zero coefficients -- so that the two lane mappings give the same bits, and a
// job's results do not depend on which mapping the batch size selects. Local states: 0 S, 1 P1, 2 P2, 3 D3, 4 D4, 5 D5.
struct LaneH6 {
  double hc10, hc20, hc13, hc14, hc25, cr, ci;   // h * couplings; rotating coupling h rot (cos, sin)
  double hE1, hE2, hE3, hE4, hE5, hg1, hg2, G1, G2;
};
__device__ __forceinline__ double stage6c(const LaneH6& H, const cplx* w, cplx* g) {
  // dp: (G1 |P1|^2) + (G2 |P2|^2), then the other block
  double own = __dadd_rn(__dmul_rn(H.G1, cnorm(w[1])), __dmul_rn(H.G2, cnorm(w[2])));
  own = __dadd_rn(own, __shfl_xor_sync(0xffffffffu, own, 1));
  const double pref = rsqrt_near1(1.0 - own);
  cplx m[6];
  double hr, hi;
  // S:  c10 P1, c20 P2
  hr = __dmul_rn(H.hc10, w[1].re); hi = __dmul_rn(H.hc10, w[1].im);
  hr = fma(H.hc20, w[2].re, hr); hi = fma(H.hc20, w[2].im, hi);
  m[0].re = __dadd_rn(w[0].re, hi); m[0].im = __dadd_rn(w[0].im, -hr);
  // P1: (E1 - i g1) P1, c10 S, c13 D3, c14 D4
  hr = fma(H.hE1, w[1].re, __dmul_rn(H.hg1, w[1].im)); hi = fma(H.hE1, w[1].im, -__dmul_rn(H.hg1, w[1].re));
  hr = fma(H.hc10, w[0].re, hr); hi = fma(H.hc10, w[0].im, hi);
  hr = fma(H.hc13, w[3].re, hr); hi = fma(H.hc13, w[3].im, hi);
  hr = fma(H.hc14, w[4].re, hr); hi = fma(H.hc14, w[4].im, hi);
  m[1].re = __dadd_rn(w[1].re, hi); m[1].im = __dadd_rn(w[1].im, -hr);
  // D3: E3 D3, c13 P1
  hr = fma(H.hE3, w[3].re, __dmul_rn(H.hc13, w[1].re)); hi = fma(H.hE3, w[3].im, __dmul_rn(H.hc13, w[1].im));
  m[3].re = __dadd_rn(w[3].re, hi); m[3].im = __dadd_rn(w[3].im, -hr);
  // P2: (E2 - i g2) P2, conj(rot) D4, c25 D5, c20 S
  hr = fma(H.hE2, w[2].re, __dmul_rn(H.hg2, w[2].im)); hi = fma(H.hE2, w[2].im, -__dmul_rn(H.hg2, w[2].re));
  hr = fma(H.cr, w[4].re, fma(H.ci, w[4].im, hr)); hi = fma(H.cr, w[4].im, fma(-H.ci, w[4].re, hi));
  hr = fma(H.hc25, w[5].re, hr); hi = fma(H.hc25, w[5].im, hi);
  hr = fma(H.hc20, w[0].re, hr); hi = fma(H.hc20, w[0].im, hi);
  m[2].re = __dadd_rn(w[2].re, hi); m[2].im = __dadd_rn(w[2].im, -hr);
  // D4: E4 D4, rot P2, c14 P1
  hr = __dmul_rn(H.hE4, w[4].re); hi = __dmul_rn(H.hE4, w[4].im);
  hr = fma(H.cr, w[2].re, fma(-H.ci, w[2].im, hr)); hi = fma(H.cr, w[2].im, fma(H.ci, w[2].re, hi));
  hr = fma(H.hc14, w[1].re, hr); hi = fma(H.hc14, w[1].im, hi);
  m[4].re = __dadd_rn(w[4].re, hi); m[4].im = __dadd_rn(w[4].im, -hr);
  // D5: E5 D5, c25 P2
  hr = fma(H.hE5, w[5].re, __dmul_rn(H.hc25, w[2].re)); hi = fma(H.hE5, w[5].im, __dmul_rn(H.hc25, w[2].im));
  m[5].re = __dadd_rn(w[5].re, hi); m[5].im = __dadd_rn(w[5].im, -hr);
#pragma unroll
  for (int k = 0; k < 6; k++) {
    g[k].re = fma(pref, m[k].re, -w[k].re);
    g[k].im = fma(pref, m[k].im, -w[k].im);
  }
  return own;
}

// ------------------------------------------------------------------------------------------------------------
// the fused kernel: nsub x { step(); qstep(); } for one ion per lane pair
// ------------------------------------------------------------------------------------------------------------
template <int NL, bool FORCED>
__global__ void __launch_bounds__(128) k_substeps(QTArgs a, QTConsts C) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = gid & 1;
  const long long slot = gid >> 1;
  const bool inrange = slot < (long long)a.nrows * a.B;
  // inactive lanes still run (the shuffles are warp-wide) on a harmless shadow of the first ion; they never store
  const int b = inrange ? (int)(slot / a.nrows) : 0;
  const int i0 = inrange ? a.row0 + (int)(slot % a.nrows) : a.row0;
  const bool active = inrange && i0 < (a.nb ? a.nb[b] : a.N);  // ensembles: trajectory b holds nb[b] <= N ions
  const int i = active ? i0 : a.row0;
  const uint64_t seed = a.seeds ? a.seeds[b] : a.seed;
  const int S = (NL == 6) ? 12 : a.S;  // compile-time stride for the 12-level hot path
  // the 12-level scheme always kicks and tracks tPart together with step() (SU); the small schemes choose at run time
  const bool do_kick = (NL == 6) ? (a.do_step != 0) : (a.do_kick != 0);
  const bool do_tpart = (NL == 6) ? (a.do_step != 0) : (a.do_tpart != 0);

  double* __restrict__ Rb = a.R + (size_t)b * 3 * a.ld;
  double* __restrict__ Vb = a.V + (size_t)b * 3 * a.ld;
  const double* __restrict__ Fb = a.F + (size_t)b * 3 * a.ld;
  double* __restrict__ Pb = a.psi + (size_t)b * 2 * S * a.ld;

  // per-lane constants into registers once (block A for even lanes, block B for odd lanes)
  const QTLane& LA = C.lane[0];
  const QTLane& LB = C.lane[1];
  int map[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) map[k] = lane ? LB.map[k] : LA.map[k];
  const double h = C.h;
  LaneH H;
  H.hc10 = h * (lane ? LB.c10 : LA.c10); H.hc20 = h * (lane ? LB.c20 : LA.c20);
  H.hc13 = h * (lane ? LB.c13 : LA.c13); H.hc14 = h * (lane ? LB.c14 : LA.c14); H.hc25 = h * (lane ? LB.c25 : LA.c25);
  const double gam1 = lane ? LB.gam1 : LA.gam1, gam2 = lane ? LB.gam2 : LA.gam2;
  H.hg1 = 0.5 * h * gam1; H.hg2 = 0.5 * h * gam2; H.G1 = h * gam1; H.G2 = h * gam2;
  H.hE3 = H.hE4 = H.hE5 = 0.0; H.hrr = H.hri = 0.0;
  const double hrot = h * (lane ? LB.rot : LA.rot);
  const double ksA = C.kick_sp * (lane ? LB.gA : LA.gA), ksB = C.kick_sp * (lane ? LB.gB : LA.gB);
  const double kd0 = C.kick_dp * (lane ? LB.gD[0] : LA.gD[0]), kd1 = C.kick_dp * (lane ? LB.gD[1] : LA.gD[1]);
  const double kd2 = C.kick_dp * (lane ? LB.gD[2] : LA.gD[2]), kd3 = C.kick_dp * (lane ? LB.gD[3] : LA.gD[3]);
  const double hG0 = h * C.gam[0], hG1 = h * C.gam[1], hG2 = h * C.gam[2], hG3 = h * C.gam[3];
  // 12-level path: the operations of the four-lane kernel (stage6c), its energy formulas and its optical-force weights
  LaneH6 H6;
  H6.hc10 = H.hc10; H6.hc20 = H.hc20; H6.hc13 = H.hc13; H6.hc14 = H.hc14; H6.hc25 = H.hc25;
  H6.hg1 = H.hg1; H6.hg2 = H.hg2; H6.G1 = H.G1; H6.G2 = H.G2;
  H6.cr = H6.ci = 0.0; H6.hE1 = H6.hE2 = H6.hE3 = H6.hE4 = H6.hE5 = 0.0;
  const double dEDP6 = -a.detuning + a.detuningDP;
  const double e0P = -h * a.detuning, e0D = h * dEDP6;
  const double e1P1 = -h, e1P2 = h, e1D3 = h * (a.kRat - 1), e1D4 = -h * (1 + a.kRat), e1D5 = h * (1 - a.kRat);
  const PrepC pc6 = {0, 0, 0, 0, 0, 0, hrot, a.pv2qv, 2. * (1 + a.kRat), a.g2E,
                     0.0126 * a.fracOfSig * a.Te, sqrt(a.density) * a.sig0, 0.00014314 * a.Te / (a.density * a.sig0 * a.sig0), a.fracOfSig != 0.0};
  const double k1a = C.kick_sp * (lane ? LB.gA : LA.gA), k2a = C.kick_dp * (lane ? LB.gD[1] : LA.gD[1]);
  const double k1b = -C.kick_dp * (lane ? LB.gD[0] : LA.gD[0]), k3b = -C.kick_sp * (lane ? LB.gB : LA.gB);
  const double k4b = -C.kick_dp * (lane ? LB.gD[2] : LA.gD[2]), k5b = -C.kick_dp * (lane ? LB.gD[3] : LA.gD[3]);

  pdl_wait();  // forces / state come from the previous kernels in the stream
  if ((threadIdx.x & 31) == 0) stamp_time(a.stamp, 0);
  cplx y[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) {
    if (map[k] >= 0) { y[k].re = Pb[(size_t)(2 * map[k]) * a.ld + i]; y[k].im = Pb[(size_t)(2 * map[k] + 1) * a.ld + i]; }
    else { y[k].re = 0.0; y[k].im = 0.0; }
  }
  // lane A carries (x, y), lane B carries (x, z); x is advanced redundantly (bitwise identically) by both
  const int c2 = 1 + lane;
  double rx = 0, r2 = 0, vx = Vb[i], v2 = 0, fx = 0, f2 = 0, tp = 0;
  if (a.do_step) {
    rx = Rb[i]; r2 = Rb[(size_t)c2 * a.ld + i];
    v2 = Vb[(size_t)c2 * a.ld + i];
    load_forces(a, b, i, c2, active && lane == 0, active, fx, f2);
  }
  if (do_tpart) tp = a.tPart[(size_t)b * a.ld + i];
  double t = a.clock ? a.clock[0] : a.t0;  // device clock inside a replayed CUDA graph
  const uint64_t substep0 = a.clock ? *reinterpret_cast<const unsigned long long*>(a.clock + 1) : a.substep0;
  const double DT = 0.5 * a.dtq;
  const double dEDP = -a.detuning + a.detuningDP;

  for (int s = 0; s < a.nsub; s++) {
    // ---------------- step(): R += V dt/2 ; V += F dt ; R += V dt/2, single wrap into [0,L] (SU:356-430) ------
    if (a.do_step) {
      const bool started = t > 0;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        if (started) {
          rx = __dadd_rn(rx, __dmul_rn(DT, vx));
          r2 = __dadd_rn(r2, __dmul_rn(DT, v2));
        } else {  // first substep of a new run: 2nd-order start (SU:370-379)
          rx = __dadd_rn(rx, __dadd_rn(__dmul_rn(DT, vx), __dmul_rn(__dmul_rn(DT, DT), fx)));
          r2 = __dadd_rn(r2, __dadd_rn(__dmul_rn(DT, v2), __dmul_rn(__dmul_rn(DT, DT), f2)));
        }
        if (rx < 0) rx = __dadd_rn(rx, a.L);
        if (rx > a.L) rx = __dadd_rn(rx, -a.L);
        if (r2 < 0) r2 = __dadd_rn(r2, a.L);
        if (r2 > a.L) r2 = __dadd_rn(r2, -a.L);
        if (half == 0) {
          vx = __dadd_rn(vx, __dmul_rn(a.dtq, fx));
          v2 = __dadd_rn(v2, __dmul_rn(a.dtq, f2));
        }
      }
    }
    // ---------------- qstep() (SU:478-714 / MC408L:578-755) ---------------------------------------------------
    double expDet = 0.0;
    if (a.fracOfSig != 0.0)  // SU:447, from the global time before it is advanced
      expDet = 0.0126 * a.fracOfSig * a.Te * t /
               (sqrt(a.density) * a.sig0 * sqrt(1 + 0.00014314 * t * t * a.Te / (a.density * a.sig0 * a.sig0)));
    const double vq = vx * a.pv2qv;
    if (do_tpart) tp = __dadd_rn(tp, a.dtq);

    // the uniforms of this (ion, substep): both lanes of an ion draw the same numbers
    double u0, u1;
    const uint64_t sidx = substep0 + (uint64_t)s;
    if (FORCED) {
      const double* up = a.forced_u + ((size_t)s * a.N + i) * 5;
      u0 = up[0]; u1 = up[1];
    } else {
      uint4 o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 0);
      u0 = u52(o.x, o.y); u1 = u52(o.z, o.w);
    }

    // P populations of both blocks in the reference's state order 2,3,4,5 -> identical jump decision in both lanes
    const double nown1 = cnorm(y[1]), nown2 = cnorm(y[2]);
    const double noth1 = __shfl_xor_sync(0xffffffffu, nown1, 1), noth2 = __shfl_xor_sync(0xffffffffu, nown2, 1);
    double n2, n3, n4, n5;
    if (NL == 6) {  // 12-level: idx2 = B.P1, idx3 = A.P1, idx4 = B.P2, idx5 = A.P2
      n2 = lane ? nown1 : noth1; n3 = lane ? noth1 : nown1; n4 = lane ? nown2 : noth2; n5 = lane ? noth2 : nown2;
    } else {        // 7-level:  idx2 = A.P1, idx3 = B.P1, idx4 = A.P2, idx5 = B.P2  (5- and 3-level: unused slots hold 0)
      n2 = lane ? noth1 : nown1; n3 = lane ? nown1 : noth1; n4 = lane ? noth2 : nown2; n5 = lane ? nown2 : noth2;
    }
    const double dp0 = hG0 * n2 + hG1 * n3 + hG2 * n4 + hG3 * n5;   // SU:484-485
    const bool jump = !(u0 > dp0);                                   // SU:487

    // ---- no-jump branch, evaluated by every lane (jumps are rare, and a warp-uniform flow keeps the cross-lane
    //      exchanges on plain full-mask shuffles); lanes that jump discard the result below ----
    double kick = 0.0;
    if (do_kick) {  // optical force from the pre-step coherences (SU:490-503; TS:170-174)
      if (NL == 6) {  // the four-lane kernel's two half-block chains, then their sum
        const double kA = fma(k2a, im_acb(y[3], y[1]), __dmul_rn(k1a, im_acb(y[0], y[1])));
        double kB = __dmul_rn(k1b, im_acb(y[2], y[4]));
        kB = fma(k3b, im_acb(y[0], y[2]), kB); kB = fma(k4b, im_acb(y[5], y[2]), kB); kB = fma(k5b, im_acb(y[4], y[1]), kB);
        kick = __dadd_rn(kA, kB);
      } else {
        kick = ksA * im_acb(y[0], y[1]) - ksB * im_acb(y[0], y[2]);
      }
    }
    if (NL == 6) {
      // energies e0 + e1 (vq + expDetuning) and the rotating coupling exactly as the four-lane kernel forms them (SU:506-510)
      const double uu = prep_uu(pc6, vx, t);
      H6.hE1 = fma(e1P1, uu, e0P); H6.hE2 = fma(e1P2, uu, e0P);
      H6.hE3 = fma(e1D3, uu, e0D); H6.hE4 = fma(e1D4, uu, e0D); H6.hE5 = fma(e1D5, uu, e0D);
      prep_rot(pc6, uu, tp, H6.cr, H6.ci);
    } else {
      H.hE1 = h * (-a.detuning - vq - expDet);                      // totalDetRightSP (MC408L:595)
      H.hE2 = h * (-a.detuning + vq + expDet);                      // totalDetLeftSP
    }
#define STAGE(v) do { if (NL == 6) stage6c(H6, v, g); else stage<NL>(H, v, g); } while (0)
    cplx yn[NL];
    {
      cplx w[NL], g[NL], acc[NL];
      STAGE(y);
#pragma unroll
      for (int k = 0; k < NL; k++) { acc[k] = g[k]; w[k].re = fma(0.5, g[k].re, y[k].re); w[k].im = fma(0.5, g[k].im, y[k].im); }
      STAGE(w);
#pragma unroll
      for (int k = 0; k < NL; k++) {
        acc[k].re = fma(3.0, g[k].re, acc[k].re); acc[k].im = fma(3.0, g[k].im, acc[k].im);
        w[k].re = fma(0.5, g[k].re, y[k].re); w[k].im = fma(0.5, g[k].im, y[k].im);
      }
      STAGE(w);
#pragma unroll
      for (int k = 0; k < NL; k++) {
        acc[k].re = fma(3.0, g[k].re, acc[k].re); acc[k].im = fma(3.0, g[k].im, acc[k].im);
        w[k].re = y[k].re + g[k].re; w[k].im = y[k].im + g[k].im;
      }
      STAGE(w);
#pragma unroll
      for (int k = 0; k < NL; k++) {
        yn[k].re = fma(0.125, acc[k].re + g[k].re, y[k].re);
        yn[k].im = fma(0.125, acc[k].im + g[k].im, y[k].im);
      }
    }
#undef STAGE
    if (!jump) {
#pragma unroll
      for (int k = 0; k < NL; k++) y[k] = yn[k];
    } else {
      // ---- quantum jump (SU:573-703 / MC408L:674-752): both lanes take identical decisions, no cross-lane traffic ----
      double u2, u3, u4;
      if (FORCED) {
        const double* up = a.forced_u + ((size_t)s * a.N + i) * 5;
        u2 = up[2]; u3 = up[3]; u4 = up[4];
      } else {
        uint4 o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 1);
        u2 = u52(o.x, o.y); u3 = u52(o.z, o.w);
        o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 2);
        u4 = u52(o.x, o.y);
      }
      tp = 0.0;
      const double tot = n2 + n3 + n4 + n5;
      const double p3 = n2 / tot, p4 = n3 / tot, p5 = n4 / tot;
      const bool sDecay = !(u2 < C.dfrac);
      kick = 0.0;
      if (do_kick && lane == 0) {  // counted once in the cross-lane sum
        double mag = sDecay ? a.vKick : a.vKickDP;
        kick = (u3 < 0.5) ? mag : -mag;
      }
      int dest;
      if (NL == 6) {
        if (u1 < p3) dest = sDecay ? 1 : (u4 < C.tD[0] ? 11 : (u4 < C.tD[1] ? 10 : 9));
        else if (u1 < p3 + p4) dest = sDecay ? (u4 < C.tS[0] ? 0 : 1) : (u4 < C.tD[2] ? 10 : (u4 < C.tD[3] ? 9 : 8));
        else if (u1 < p3 + p4 + p5) dest = sDecay ? (u4 < C.tS[1] ? 1 : 0) : (u4 < C.tD[4] ? 9 : (u4 < C.tD[5] ? 8 : 7));
        else dest = sDecay ? 0 : (u4 < C.tD[6] ? 8 : (u4 < C.tD[7] ? 7 : 6));
      } else if (a.scheme == 7) {
        // the 7-level file draws its 4th uniform (rand3) only for S decays out of states 4 and 5; with a
        // counter-based stream the slot is simply left unused otherwise
        if (u1 < p3) dest = sDecay ? 0 : 6;
        else if (u1 < p3 + p4) dest = sDecay ? (u4 < C.tS[0] ? 0 : 1) : 6;
        else if (u1 < p3 + p4 + p5) dest = sDecay ? (u4 < C.tS[1] ? 0 : 1) : 6;
        else dest = sDecay ? 1 : 6;
      } else if (a.scheme == 5) {
        // 5-level 422 nm pump (MC422L:660-720): p3 = 0 and p4 = prob3 here, so p3 + p4 is the reference's prob3
        // exactly; no randDir draw exists in that file (slot 3 unused), rand3 (slot 4) only on S decays
        if (u1 < p3 + p4) dest = sDecay ? (u4 < C.tS[0] ? 1 : 0) : 4;
        else dest = sDecay ? (u4 < C.tS[1] ? 0 : 1) : 4;
      } else {
        dest = 0;  // 3-level: back to the single ground state (TS:272-274); draws rand and randDir only
      }
#pragma unroll
      for (int k = 0; k < NL; k++) { y[k].re = (map[k] == dest) ? 1.0 : 0.0; y[k].im = 0.0; }
    }
    if (do_kick) {
      kick = kick + __shfl_xor_sync(0xffffffffu, kick, 1);
      vx = __dadd_rn(vx, kick);  // SU:705
    }
    if (a.renorm) {  // SU:706-712
      double own = 0.0;
      if (NL == 6) {  // half-block sums first, as the four-lane kernel adds them
        own = __dadd_rn(__dadd_rn(__dadd_rn(cnorm(y[0]), cnorm(y[1])), cnorm(y[3])), __dadd_rn(__dadd_rn(cnorm(y[2]), cnorm(y[4])), cnorm(y[5])));
      } else {
#pragma unroll
        for (int k = 0; k < NL; k++) own += cnorm(y[k]);
      }
      double nn = sqrt(own + __shfl_xor_sync(0xffffffffu, own, 1));
#pragma unroll
      for (int k = 0; k < NL; k++) { y[k].re /= nn; y[k].im /= nn; }
    }
    if (a.do_step) t = __dadd_rn(t, a.dtq);  // SU:716
  }

  pdl_launch_dependents();  // the force kernel's launch + prologue may overlap the stores below (it waits before reading)
  if ((threadIdx.x & 31) == 0) stamp_time(a.stamp, 1);
  if (!active) return;
#pragma unroll
  for (int k = 0; k < NL; k++)
    if (map[k] >= 0) { Pb[(size_t)(2 * map[k]) * a.ld + i] = y[k].re; Pb[(size_t)(2 * map[k] + 1) * a.ld + i] = y[k].im; }
  if (a.do_step) {
    long long* __restrict__ Xf = a.Rfix + (size_t)b * 3 * a.ld;
    Rb[(size_t)c2 * a.ld + i] = r2;
    Xf[(size_t)c2 * a.ld + i] = to_fixed(r2, a.invL, a.invL_lo);
    Vb[(size_t)c2 * a.ld + i] = v2;
    if (lane == 0) {
      Rb[i] = rx; Xf[i] = to_fixed(rx, a.invL, a.invL_lo);
      Vb[i] = vx;
    }
  } else if (do_kick && lane == 0) {
    Vb[i] = vx;
  }
  if (do_tpart && lane == 0) a.tPart[(size_t)b * a.ld + i] = tp;
}

// ------------------------------------------------------------------------------------------------------------
// K2, four lanes per ion (12-level scheme, small systems). With N ~ 3500 the two-lane kernel fills 220 of the chip's 592
// warp schedulers with ONE warp each and is bound by the latency of its own instruction stream. Here every Hamiltonian
// block is split once more -- lane (blk, 0) holds {S, P1, D3}, lane (blk, 1) holds {P2, D4, D5} -- so a warp carries
// 8 ions and fewer instructions per substep (643 vs 724 executed, ncu). Both halves run ONE uniform instruction stream with per-lane
// coefficients (no divergence):
//   row0 = D0 w0 + A01 w1 + a02 w2 + b00 r0      half0: S  (D0 = 0, A01 = c10, b00 = c20, r0 = P2)
//                                                half1: P2 (D0 = E2 - i g2, A01 = conj(rot), a02 = c25, b00 = c20, r0 = S)
//   row1 = D1 w1 + A10 w0 + a12 w2 + b11 r1      half0: P1 (D1 = E1 - i g1, A10 = c10, a12 = c13, b11 = c14, r1 = D4)
//                                                half1: D4 (D1 = E4, A10 = rot, b11 = c14, r1 = P1)
//   row2 = E2' w2 + a21 w1 + a20 w0              half0: D3 (a21 = c13)          half1: D5 (a20 = c25)
// with (r0, r1) = the partner half's (w0, w1), one shuffle pair per stage. Same physics, same uniforms, same jump
// logic as k_substeps; sums that cross lanes are formed in a different (fixed) order, so results agree to rounding.
// ------------------------------------------------------------------------------------------------------------
struct Lane4H {
  double hE0, hg0, hE1, hg1, hE2;   // h * (energy, Gamma/2) of the three local rows
  double a01r, a01i, a10r, a10i;    // h * complex couplings row0<-w1, row1<-w0
  double a02, b00, a12, b11, a21, a20;
  double G0, G1;                    // h * Gamma weights of |w0|^2, |w1|^2 in dp
  double Gr0, Gr1;                  // the partner half's weights (of r0 = its w0, r1 = its w1)
};

// Build-time variants of the four-lane kernel (A/B'd on B200; each gives the same bits as the plain form unless noted):
//   MDQT_K2_DPLOCAL 1: the block's P population is formed from the partner amplitudes every stage fetches anyway (r0, r1)
//                      instead of one more shuffle round -- same operands, same order, one dependent shuffle less per stage
//   MDQT_K2_DPSTAGE 1: the jump test uses stage 1's dp (pair sums first: (a+b)+(c+d) instead of the reference's left-to-right
//                      order: differs by <= 1 ulp, i.e. only if the uniform ties with dp); the four P populations are
//                      fetched only inside the (rare) jump branch
//   MDQT_K2_PIPE    1: the next substep's Doppler shift, rotating phase, sin/cos and coefficients are computed right after
//                      this substep's optical-force kick is known, overlapping the Runge-Kutta chains; recomputed on a jump
#ifndef MDQT_K2_DPLOCAL
#define MDQT_K2_DPLOCAL 0  // measured on B200, N = 3500, 25 substeps in the replayed graph, with __launch_bounds__(32): 21.9 us (1: 22.1). With
                           // __launch_bounds__(128) it was the other way round (1: 23.6 -> 22.6 with the one-round-trip force prologue; 0: 28.2):
                           // the variants differ in ptxas' schedule of the loop, not in work (DPSTAGE 23.6; PIPE 22.6-23.5; PHILOX4 22.1-31.2)
#endif
#ifndef MDQT_K2_DPSTAGE
#define MDQT_K2_DPSTAGE 0
#endif
#ifndef MDQT_K2_PIPE
#define MDQT_K2_PIPE 0
#endif
//   MDQT_K2_PHILOX4 1: the four lanes of an ion draw the uniforms of four consecutive substeps (one Philox call per lane every four
//                      substeps, passed round the quad by shuffle) instead of all running the same call: a quarter of the integer work,
//                      same bits -- and 29.3 instead of 22.7 us: ptxas answers the changed loop with a 168-register schedule and spills
#ifndef MDQT_K2_PHILOX4
#define MDQT_K2_PHILOX4 0
#endif

__device__ __forceinline__ double stage4(const Lane4H& H, const cplx* w, cplx* g) {
  const cplx r0 = {__shfl_xor_sync(0xffffffffu, w[0].re, 1), __shfl_xor_sync(0xffffffffu, w[0].im, 1)};
  const cplx r1 = {__shfl_xor_sync(0xffffffffu, w[1].re, 1), __shfl_xor_sync(0xffffffffu, w[1].im, 1)};
  double own = fma(H.G0, cnorm(w[0]), H.G1 * cnorm(w[1]));
#if MDQT_K2_DPLOCAL
  own += fma(H.Gr0, cnorm(r0), H.Gr1 * cnorm(r1));  // == the partner lane's `own`, bit for bit
#else
  own += __shfl_xor_sync(0xffffffffu, own, 1);
#endif
  own += __shfl_xor_sync(0xffffffffu, own, 2);
  const double pref = rsqrt_near1(1.0 - own);
  cplx m[3];
  {  // row0
    double hr = fma(H.hE0, w[0].re, H.hg0 * w[0].im), hi = fma(H.hE0, w[0].im, -(H.hg0 * w[0].re));
    hr = fma(H.a01r, w[1].re, fma(-H.a01i, w[1].im, hr)); hi = fma(H.a01r, w[1].im, fma(H.a01i, w[1].re, hi));
    hr = fma(H.a02, w[2].re, hr); hi = fma(H.a02, w[2].im, hi);
    hr = fma(H.b00, r0.re, hr); hi = fma(H.b00, r0.im, hi);
    m[0].re = w[0].re + hi; m[0].im = w[0].im - hr;
  }
  {  // row1
    double hr = fma(H.hE1, w[1].re, H.hg1 * w[1].im), hi = fma(H.hE1, w[1].im, -(H.hg1 * w[1].re));
    hr = fma(H.a10r, w[0].re, fma(-H.a10i, w[0].im, hr)); hi = fma(H.a10r, w[0].im, fma(H.a10i, w[0].re, hi));
    hr = fma(H.a12, w[2].re, hr); hi = fma(H.a12, w[2].im, hi);
    hr = fma(H.b11, r1.re, hr); hi = fma(H.b11, r1.im, hi);
    m[1].re = w[1].re + hi; m[1].im = w[1].im - hr;
  }
  {  // row2
    double hr = fma(H.hE2, w[2].re, fma(H.a21, w[1].re, H.a20 * w[0].re));
    double hi = fma(H.hE2, w[2].im, fma(H.a21, w[1].im, H.a20 * w[0].im));
    m[2].re = w[2].re + hi; m[2].im = w[2].im - hr;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    g[k].re = fma(pref, m[k].re, -w[k].re);
    g[k].im = fma(pref, m[k].im, -w[k].im);
  }
  return own;
}

// the kernel is launched as one-warp CTAs; telling ptxas so (instead of 128) gives a schedule that is 0.4 us faster
#ifndef MDQT_K2_LB
#define MDQT_K2_LB 32
#endif
#ifndef MDQT_K2_UNROLL
#define MDQT_K2_UNROLL 1
#endif
constexpr int kK2Unroll = MDQT_K2_UNROLL;
template <bool FORCED>
__global__ void __launch_bounds__(MDQT_K2_LB) k_substeps4(QTArgs a, QTConsts C) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = gid & 3, blk = q >> 1, half = q & 1;
  const long long slot = gid >> 2;
  const bool inrange = slot < (long long)a.nrows * a.B;
  const int b = inrange ? (int)(slot / a.nrows) : 0;
  const int i0 = inrange ? a.row0 + (int)(slot % a.nrows) : a.row0;
  const bool active = inrange && i0 < (a.nb ? a.nb[b] : a.N);  // ensembles: trajectory b holds nb[b] <= N ions
  const int i = active ? i0 : a.row0;
  const uint64_t seed = a.seeds ? a.seeds[b] : a.seed;

  double* __restrict__ Rb = a.R + (size_t)b * 3 * a.ld;
  double* __restrict__ Vb = a.V + (size_t)b * 3 * a.ld;
  const double* __restrict__ Fb = a.F + (size_t)b * 3 * a.ld;
  double* __restrict__ Pb = a.psi + (size_t)b * 24 * a.ld;

  // local state -> reference state: half0 = {S, P1, D3}, half1 = {P2, D4, D5} of block blk (QTLane.map order S,P1,P2,D3,D4,D5)
  const QTLane& LA = C.lane[0];
  const QTLane& LB = C.lane[1];
#define LSEL(f) (blk ? LB.f : LA.f)
  int map[3];
  map[0] = half ? LSEL(map[2]) : LSEL(map[0]);
  map[1] = half ? LSEL(map[4]) : LSEL(map[1]);
  map[2] = half ? LSEL(map[5]) : LSEL(map[3]);
  const double h = C.h;
  const double hc10 = h * LSEL(c10), hc20 = h * LSEL(c20), hc13 = h * LSEL(c13), hc14 = h * LSEL(c14), hc25 = h * LSEL(c25);
  const double hrot = h * LSEL(rot);
  const double gam1 = LSEL(gam1), gam2 = LSEL(gam2);
  const double dEDP = -a.detuning + a.detuningDP;
  Lane4H H;
  // energies as e0 + e1 * (vq + expDetuning)  (SU:506-510), pre-multiplied by h
  const double e0_0 = half ? -h * a.detuning : 0.0, e1_0 = half ? h : 0.0;                       // S | P2: -det + u
  const double e0_1 = half ? h * dEDP : -h * a.detuning, e1_1 = half ? -h * (1 + a.kRat) : -h;   // P1: -det - u | D4
  const double e0_2 = h * dEDP, e1_2 = half ? h * (1 - a.kRat) : h * (a.kRat - 1);               // D3 | D5
  H.hg0 = half ? 0.5 * h * gam2 : 0.0; H.hg1 = half ? 0.0 : 0.5 * h * gam1;
  H.G0 = half ? h * gam2 : 0.0; H.G1 = half ? 0.0 : h * gam1;
  H.Gr0 = half ? 0.0 : h * gam2; H.Gr1 = half ? h * gam1 : 0.0;
  H.a02 = half ? hc25 : 0.0; H.b00 = hc20; H.a12 = half ? 0.0 : hc13; H.b11 = hc14;
  H.a21 = half ? 0.0 : hc13; H.a20 = half ? hc25 : 0.0;
  H.a01r = hc10; H.a01i = 0.0; H.a10r = hc10; H.a10i = 0.0;  // half1: set from the rotating phase every substep
  // optical-force weights (SU:503), generic form k1 Im(w0 w1*) + k2 Im(w2 w1*) + k3 Im(r0 w0*) + k4 Im(w2 w0*) + k5 Im(w1 r1*)
  const double k1 = half ? -C.kick_dp * LSEL(gD[0]) : C.kick_sp * LSEL(gA);
  const double k2 = half ? 0.0 : C.kick_dp * LSEL(gD[1]);
  const double k3 = half ? -C.kick_sp * LSEL(gB) : 0.0;
  const double k4 = half ? -C.kick_dp * LSEL(gD[2]) : 0.0;
  const double k5 = half ? -C.kick_dp * LSEL(gD[3]) : 0.0;
#undef LSEL
  const double hG0 = h * C.gam[0], hG1 = h * C.gam[1], hG2 = h * C.gam[2], hG3 = h * C.gam[3];
  const int base = threadIdx.x & 28;  // first lane of this ion's quad within the warp

  pdl_wait();
#if MDQT_K2_TRIGGER == 1
  // Every CTA of this kernel is resident from the start (one warp per scheduler at most), so the dependent force kernel may be
  // launched NOW: one of its two CTAs per SM fits beside these warps (226 x 32 x 4 + 126 x 256 registers), loads its exp table and
  // waits at its griddepcontrol.wait -- which returns only when this grid has completed and flushed -- so launch latency and prologue
  // hide behind the 25 substeps instead of following them
  pdl_launch_dependents();
#endif
  if ((threadIdx.x & 31) == 0) stamp_time(a.stamp, 0);
  cplx y[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { y[k].re = Pb[(size_t)(2 * map[k]) * a.ld + i]; y[k].im = Pb[(size_t)(2 * map[k] + 1) * a.ld + i]; }
  // quad lane 0 carries (x, y), lane 1 (x, z); lanes 2, 3 shadow (x, y) without storing
  const int c2 = (q == 1) ? 2 : 1;
  double rx = Rb[i], r2 = Rb[(size_t)c2 * a.ld + i], vx = Vb[i], v2 = Vb[(size_t)c2 * a.ld + i];
  double fx_, f2_;
#if MDQT_K2_FLOAD1
  {
    double fown = 0.0;
    if (q < 3) fown = load_force1_impl(a.fpart, a.Fw, a.F, a.nb ? a.nb[b] : a.N, a.jl ? a.jl[b] : a.fp_jlen, a.B, a.ld, b, i, q, active ? 1 : 0);
    const double fy = __shfl_sync(0xffffffffu, fown, base + 1), fz = __shfl_sync(0xffffffffu, fown, base + 2);
    fx_ = __shfl_sync(0xffffffffu, fown, base);
    f2_ = (q == 1) ? fz : fy;
  }
#else
  load_forces(a, b, i, c2, active && q == 0, active && q < 2, fx_, f2_);
#endif
  const double fx = fx_, f2 = f2_;
  double tp = a.tPart[(size_t)b * a.ld + i];
  double t = a.clock ? a.clock[0] : a.t0;  // device clock inside a replayed CUDA graph
  const uint64_t substep0 = a.clock ? *reinterpret_cast<const unsigned long long*>(a.clock + 1) : a.substep0;
  const double DT = 0.5 * a.dtq;

  // Doppler shift, rotating phase and the coefficients that depend on them, for a substep that starts (after step()'s velocity
  // update) with velocity vxs, time-since-jump tps (already advanced) and global time ts (SU:447, 481-483, 506-510)
  const PrepC pc = {e0_0, e1_0, e0_1, e1_1, e0_2, e1_2, hrot, a.pv2qv, 2. * (1 + a.kRat), a.g2E,
                    0.0126 * a.fracOfSig * a.Te, sqrt(a.density) * a.sig0, 0.00014314 * a.Te / (a.density * a.sig0 * a.sig0), a.fracOfSig != 0.0};
#define prep(vxs, tps, ts) prep4(pc, (vxs), (tps), (ts))
#if MDQT_K2_PIPE
  // the first substep's coefficients: what step() will make of vx, and tp + dtq
  Pre cur = prep(__dadd_rn(vx, __dmul_rn(a.dtq, fx)), __dadd_rn(tp, a.dtq), t);
#endif
  const unsigned quadmask = 0xFu << base;
#if MDQT_K2_PHILOX4
  double uq0 = 0.0, uq1 = 0.0;
#endif

#pragma unroll(kK2Unroll)
  for (int s = 0; s < a.nsub; s++) {
    {  // step() (SU:356-430)
      const bool started = t > 0;
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        if (started) {
          rx = __dadd_rn(rx, __dmul_rn(DT, vx));
          r2 = __dadd_rn(r2, __dmul_rn(DT, v2));
        } else {
          rx = __dadd_rn(rx, __dadd_rn(__dmul_rn(DT, vx), __dmul_rn(__dmul_rn(DT, DT), fx)));
          r2 = __dadd_rn(r2, __dadd_rn(__dmul_rn(DT, v2), __dmul_rn(__dmul_rn(DT, DT), f2)));
        }
        if (rx < 0) rx = __dadd_rn(rx, a.L);
        if (rx > a.L) rx = __dadd_rn(rx, -a.L);
        if (r2 < 0) r2 = __dadd_rn(r2, a.L);
        if (r2 > a.L) r2 = __dadd_rn(r2, -a.L);
        if (hf == 0) {
          vx = __dadd_rn(vx, __dmul_rn(a.dtq, fx));
          v2 = __dadd_rn(v2, __dmul_rn(a.dtq, f2));
        }
      }
    }
    tp = __dadd_rn(tp, a.dtq);
#if !MDQT_K2_PIPE
    const Pre cur = prep(vx, tp, t);
#endif

    double u0, u1;
    const uint64_t sidx = substep0 + (uint64_t)s;
    if (FORCED) {
      const double* up = a.forced_u + ((size_t)s * a.N + i) * 5;
      u0 = up[0]; u1 = up[1];
    } else {
#if MDQT_K2_PHILOX4
      // the four lanes of an ion would all run the same Philox call: instead lane q draws the pair of substep s + q once every
      // four substeps and the quad passes them round (the counter-based stream is the same, a quarter of the integer work)
      if ((s & 3) == 0) {
        const uint4 o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx + (uint64_t)q, 0);
        uq0 = u52(o.x, o.y); uq1 = u52(o.z, o.w);
      }
      u0 = __shfl_sync(0xffffffffu, uq0, base + (s & 3)); u1 = __shfl_sync(0xffffffffu, uq1, base + (s & 3));
#else
      uint4 o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 0);
      u0 = u52(o.x, o.y); u1 = u52(o.z, o.w);
#endif
    }
    // P populations of the quad in the reference's state order 2,3,4,5: lanes hold P1(A)=3, P2(A)=5, P1(B)=2, P2(B)=4
    const double pn = half ? cnorm(y[0]) : cnorm(y[1]);
#if !MDQT_K2_DPSTAGE
    const double n3 = __shfl_sync(0xffffffffu, pn, base), n5 = __shfl_sync(0xffffffffu, pn, base + 1);
    const double n2 = __shfl_sync(0xffffffffu, pn, base + 2), n4 = __shfl_sync(0xffffffffu, pn, base + 3);
    const double dp0 = hG0 * n2 + hG1 * n3 + hG2 * n4 + hG3 * n5;
#endif

    H.hE0 = cur.hE0; H.hE1 = cur.hE1; H.hE2 = cur.hE2;
    if (half) { H.a01r = cur.cr; H.a01i = -cur.ci; H.a10r = cur.cr; H.a10i = cur.ci; }
    cplx yn[3];
    double kick;
#if MDQT_K2_PIPE
    Pre nxt;
#endif
    {
      cplx w[3], g[3], acc[3];
      // stage 1 also yields the partner amplitudes needed by the optical force (pre-step coherences)
      const cplx r0 = {__shfl_xor_sync(0xffffffffu, y[0].re, 1), __shfl_xor_sync(0xffffffffu, y[0].im, 1)};
      const cplx r1 = {__shfl_xor_sync(0xffffffffu, y[1].re, 1), __shfl_xor_sync(0xffffffffu, y[1].im, 1)};
      // one explicit chain (zero weights are exact no-ops): the two-lane kernel forms the same sums in the same order
      kick = __dmul_rn(k1, im_acb(y[0], y[1]));
      kick = fma(k2, im_acb(y[2], y[1]), kick); kick = fma(k3, im_acb(r0, y[0]), kick);
      kick = fma(k4, im_acb(y[2], y[0]), kick); kick = fma(k5, im_acb(y[1], r1), kick);
#if MDQT_K2_PIPE
      // the no-jump kick is final here: sum it over the quad now and start on the next substep's coefficients
      kick += __shfl_xor_sync(0xffffffffu, kick, 1);
      kick += __shfl_xor_sync(0xffffffffu, kick, 2);
      nxt = prep(__dadd_rn(__dadd_rn(vx, kick), __dmul_rn(a.dtq, fx)), __dadd_rn(tp, a.dtq), __dadd_rn(t, a.dtq));
#endif
      const double dp1 = stage4(H, y, g);
#if MDQT_K2_DPSTAGE
      const double dp0 = dp1;
#else
      (void)dp1;
#endif
#pragma unroll
      for (int k = 0; k < 3; k++) { acc[k] = g[k]; w[k].re = fma(0.5, g[k].re, y[k].re); w[k].im = fma(0.5, g[k].im, y[k].im); }
      stage4(H, w, g);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        acc[k].re = fma(3.0, g[k].re, acc[k].re); acc[k].im = fma(3.0, g[k].im, acc[k].im);
        w[k].re = fma(0.5, g[k].re, y[k].re); w[k].im = fma(0.5, g[k].im, y[k].im);
      }
      stage4(H, w, g);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        acc[k].re = fma(3.0, g[k].re, acc[k].re); acc[k].im = fma(3.0, g[k].im, acc[k].im);
        w[k].re = y[k].re + g[k].re; w[k].im = y[k].im + g[k].im;
      }
      stage4(H, w, g);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        yn[k].re = fma(0.125, acc[k].re + g[k].re, y[k].re);
        yn[k].im = fma(0.125, acc[k].im + g[k].im, y[k].im);
      }
      const bool jump = !(u0 > dp0);
      if (!jump) {
#pragma unroll
        for (int k = 0; k < 3; k++) y[k] = yn[k];
      } else {  // quantum jump (SU:573-703): all four lanes of the quad decide identically
#if MDQT_K2_DPSTAGE
        const double n3 = __shfl_sync(quadmask, pn, base), n5 = __shfl_sync(quadmask, pn, base + 1);
        const double n2 = __shfl_sync(quadmask, pn, base + 2), n4 = __shfl_sync(quadmask, pn, base + 3);
#endif
        double u2, u3, u4;
        if (FORCED) {
          const double* up = a.forced_u + ((size_t)s * a.N + i) * 5;
          u2 = up[2]; u3 = up[3]; u4 = up[4];
        } else {
          uint4 o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 1);
          u2 = u52(o.x, o.y); u3 = u52(o.z, o.w);
          o = philox_call(seed, (unsigned)(a.traj0 + b), (unsigned)i, sidx, 2);
          u4 = u52(o.x, o.y);
        }
        tp = 0.0;
        const double tot = n2 + n3 + n4 + n5;
        const double p3 = n2 / tot, p4 = n3 / tot, p5 = n4 / tot;
        const bool sDecay = !(u2 < C.dfrac);
        const double mag = sDecay ? a.vKick : a.vKickDP;
#if MDQT_K2_PIPE
        kick = (u3 < 0.5) ? mag : -mag;  // every lane of the quad holds the total (the sum below is already done)
        nxt = prep(__dadd_rn(__dadd_rn(vx, kick), __dmul_rn(a.dtq, fx)), a.dtq, __dadd_rn(t, a.dtq));  // tp restarts: 0 + dtq
#else
        kick = 0.0;
        if (q == 0) kick = (u3 < 0.5) ? mag : -mag;
#endif
        int dest;
        if (u1 < p3) dest = sDecay ? 1 : (u4 < C.tD[0] ? 11 : (u4 < C.tD[1] ? 10 : 9));
        else if (u1 < p3 + p4) dest = sDecay ? (u4 < C.tS[0] ? 0 : 1) : (u4 < C.tD[2] ? 10 : (u4 < C.tD[3] ? 9 : 8));
        else if (u1 < p3 + p4 + p5) dest = sDecay ? (u4 < C.tS[1] ? 1 : 0) : (u4 < C.tD[4] ? 9 : (u4 < C.tD[5] ? 8 : 7));
        else dest = sDecay ? 0 : (u4 < C.tD[6] ? 8 : (u4 < C.tD[7] ? 7 : 6));
#pragma unroll
        for (int k = 0; k < 3; k++) { y[k].re = (map[k] == dest) ? 1.0 : 0.0; y[k].im = 0.0; }
      }
    }
#if !MDQT_K2_PIPE
    kick += __shfl_xor_sync(0xffffffffu, kick, 1);
    kick += __shfl_xor_sync(0xffffffffu, kick, 2);
#endif
    vx = __dadd_rn(vx, kick);  // SU:705
    if (a.renorm) {
      double own = cnorm(y[0]) + cnorm(y[1]) + cnorm(y[2]);
      own += __shfl_xor_sync(0xffffffffu, own, 1);
      own += __shfl_xor_sync(0xffffffffu, own, 2);
      const double nn = sqrt(own);
#pragma unroll
      for (int k = 0; k < 3; k++) { y[k].re /= nn; y[k].im /= nn; }
    }
    t = __dadd_rn(t, a.dtq);  // SU:716
#if MDQT_K2_PIPE
    cur = nxt;
#endif
  }

#undef prep
#if MDQT_K2_TRIGGER != 1
  pdl_launch_dependents();  // the force kernel's launch and prologue may overlap the stores below (it waits before reading)
#endif
  if ((threadIdx.x & 31) == 0) stamp_time(a.stamp, 1);
  if (!active) return;
#pragma unroll
  for (int k = 0; k < 3; k++) { Pb[(size_t)(2 * map[k]) * a.ld + i] = y[k].re; Pb[(size_t)(2 * map[k] + 1) * a.ld + i] = y[k].im; }
  if (q < 2) {
    long long* __restrict__ Xf = a.Rfix + (size_t)b * 3 * a.ld;
    Rb[(size_t)c2 * a.ld + i] = r2;
    Xf[(size_t)c2 * a.ld + i] = to_fixed(r2, a.invL, a.invL_lo);
    Vb[(size_t)c2 * a.ld + i] = v2;
    if (q == 0) {
      Rb[i] = rx; Xf[i] = to_fixed(rx, a.invL, a.invL_lo);
      Vb[i] = vx; a.tPart[(size_t)b * a.ld + i] = tp;
    }
  }
}


void launch_substeps(const QTArgs& a, const QTConsts& C, int scheme, cudaStream_t s) {
  long long threads = 2LL * a.nrows * a.B;
  int block = threads >= 148LL * 4 * 128 ? 128 : 64;
  int grid = (int)((threads + block - 1) / block);
  bool forced = a.forced_u != nullptr;
  // Lanes per ion. One trajectory of a few thousand ions is a few hundred warps, at most one per SM sub-partition, each bound
  // by its own in-order instruction stream: there four lanes per ion (643 vs 724 executed instructions per warp-substep, twice the warps) win
  // -- 26.9 vs 28.5 us per 25 substeps at N = 3500 inside the replayed graph (profiles/r01c_k1_trace.txt). Once the warps
  // outnumber the sub-partitions the kernel is throughput-bound and two lanes (fewer instructions per ion) win. The switch
  // depends on N and the batch size only, so every rank of a row-decomposed run takes the same mapping.
  // MDQT_QT_LANES=2|4 forces one mapping (both stay parity-tested, tests/test_gpu_variants.py).
  static const int lanes_override = [] { const char* e = getenv("MDQT_QT_LANES"); return e ? atoi(e) : 0; }();
  const bool small = 4LL * a.N * a.B <= 148LL * 4 * 32;  // four-lane warps fit one per sub-partition
  // a.lanes pins a mapping (unused: the two mappings give the same bits, so the choice may follow N and B)
  const int want = lanes_override ? lanes_override : a.lanes;
  const bool four = scheme == 12 && a.do_step && (want == 4 || (want != 2 && small));
  if (four) {
    long long th = 4LL * a.nrows * a.B;
    int g4 = (int)((th + 31) / 32);
    if (forced) launch_kernel(k_substeps4<true>, dim3(g4), dim3(32), s, false, a, C);
    else launch_kernel(k_substeps4<false>, dim3(g4), dim3(32), s, a.pdl != 0, a, C);
    return;
  }
  if (scheme == 12) {
    if (forced) launch_kernel(k_substeps<6, true>, dim3(grid), dim3(block), s, false, a, C);
    else launch_kernel(k_substeps<6, false>, dim3(grid), dim3(block), s, a.pdl != 0, a, C);
  } else {
    if (forced) k_substeps<4, true><<<grid, block, 0, s>>>(a, C);
    else k_substeps<4, false><<<grid, block, 0, s>>>(a, C);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K5: MD-family velocity Verlet (MD:452-502)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_vv_positions(VVArgs a) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a.clock && g == 0) {  // graph replay: advance the counters while nobody reads them
    unsigned long long* c = reinterpret_cast<unsigned long long*>(a.clock);
    c[1] += (unsigned long long)a.adv_sub;
    c[2] += (unsigned long long)a.adv_vv;
  }
  if (g >= (long long)a.nrows * a.B * 3) return;
  int b = (int)(g / (3LL * a.nrows));
  int rem = (int)(g % (3LL * a.nrows));
  int c = rem / a.nrows, i = a.row0 + rem % a.nrows;
  size_t idx = ((size_t)b * 3 + c) * a.ld + i;
  // R = R + dt*V + dt*dt/2*A, each operation rounded as in the reference (MD:455)
  double r = __dadd_rn(__dadd_rn(a.R[idx], __dmul_rn(a.dt, a.V[idx])), __dmul_rn(__dmul_rn(a.dt, a.dt) / 2, a.A[idx]));
  if (r < 0) r = __dadd_rn(r, a.L);
  if (r > a.L) r = __dadd_rn(r, -a.L);
  a.R[idx] = r;
  a.Rfix[idx] = to_fixed(r, a.invL, a.invL_lo);
}

__global__ void k_vv_velocities(VVArgs a) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.nrows * a.B) return;
  int b = (int)(g / a.nrows);
  int i = a.row0 + (int)(g % a.nrows);
  size_t base = (size_t)b * 3 * a.ld + i;
  const uint64_t vstep = a.clock ? reinterpret_cast<const unsigned long long*>(a.clock)[2] : a.step;
  double v[3];
  bool collide = false;
  if (a.collisionFreq > 0.0) {
    double u, nrm[3];
    if (a.forced_u) {
      u = a.forced_u[i];
      nrm[0] = a.forced_n[3 * i]; nrm[1] = a.forced_n[3 * i + 1]; nrm[2] = a.forced_n[3 * i + 2];
      collide = u < a.dt * a.collisionFreq;
    } else {
      uint4 o = philox_call(a.seed, (unsigned)(a.traj0 + b), (unsigned)i, vstep, 3);
      u = u52(o.x, o.y);
      collide = u < a.dt * a.collisionFreq;  // MD:476-477
      if (collide) {                         // Box-Muller on stream calls 3..5 (orc_collision_draws)
        double ua = u52(o.z, o.w);
        o = philox_call(a.seed, (unsigned)(a.traj0 + b), (unsigned)i, vstep, 4);
        double ub = u52(o.x, o.y), uc = u52(o.z, o.w);
        o = philox_call(a.seed, (unsigned)(a.traj0 + b), (unsigned)i, vstep, 5);
        double ud = u52(o.x, o.y);
        double r = sqrt(-2.0 * log(ua)), sn, cs;
        sincos(6.283185307179586476925286766559 * ub, &sn, &cs);
        nrm[0] = a.sigma_v * (r * cs); nrm[1] = a.sigma_v * (r * sn);
        nrm[2] = a.sigma_v * (sqrt(-2.0 * log(uc)) * cos(6.283185307179586476925286766559 * ud));
      }
    }
    if (collide) { v[0] = nrm[0]; v[1] = nrm[1]; v[2] = nrm[2]; }
  }
  if (!collide) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      size_t idx = base + (size_t)c * a.ld;
      v[c] = __dadd_rn(a.V[idx], __dmul_rn(a.dt / 2, __dadd_rn(a.oldA[idx], a.A[idx])));  // MD:484-486
    }
  }
  if (a.laser == 2) {
    v[0] = __dadd_rn(v[0], __dmul_rn(__dmul_rn(v[0], a.dt), a.laser_coeff));          // MD:491
  } else if (a.laser == 1) {
    v[0] = __dadd_rn(v[0], __dmul_rn(__dmul_rn(v[0], a.dt), a.laser_coeff) / 2);      // MD:494
    v[1] = __dadd_rn(v[1], -(__dmul_rn(__dmul_rn(v[1], a.dt), a.laser_coeff) / 4));   // MD:495
    v[2] = __dadd_rn(v[2], -(__dmul_rn(__dmul_rn(v[2], a.dt), a.laser_coeff) / 4));   // MD:496
  }
#pragma unroll
  for (int c = 0; c < 3; c++) a.V[base + (size_t)c * a.ld] = v[c];
}

void launch_vv_positions(const VVArgs& a, cudaStream_t s) {
  long long n = 3LL * a.nrows * a.B;
  k_vv_positions<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
}
void launch_vv_velocities(const VVArgs& a, cudaStream_t s) {
  long long n = (long long)a.nrows * a.B;
  k_vv_velocities<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------------------
// Projective spin measurement after the pump: tagParticles() (MC408L:1022-1067, MC422L:992-1036) and
// measureSpinUps() (FZ408L:600-647, FZ422L:570-611). One thread per ion; the two uniforms of an ion are Philox
// call 6 of the (substep, ion, trajectory) counter, or u[N][2] when forced.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_tag(const double* __restrict__ psi, int S, int N, int ld, int traj0, uint64_t seed, uint64_t substep,
                      const double* __restrict__ forced_u, int* __restrict__ tagged, int* __restrict__ count) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double* p = psi + (size_t)b * 2 * S * ld;
  auto nrm = [&](int k) { const double re = p[(size_t)(2 * k) * ld + i], im = p[(size_t)(2 * k + 1) * ld + i]; return __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im)); };
  double ua, ub;
  if (forced_u) { ua = forced_u[2 * i]; ub = forced_u[2 * i + 1]; }
  else { uint4 o = philox_call(seed, (unsigned)(traj0 + b), (unsigned)i, substep, 6); ua = u52(o.x, o.y); ub = u52(o.z, o.w); }
  const double n1 = nrm(0), n3 = nrm(2), n4 = nrm(3);
  int up;
  if (S == 7) {  // MC408L:1034-1062
    const double n5 = nrm(4);
    const double c1 = __dadd_rn(n1, n3), c2 = __dadd_rn(c1, n4), c3 = __dadd_rn(c2, n5);
    up = (ua < c1) ? 1 : (ua < c2) ? (ub < 2. / 3) : (ua < c3) ? (ub < 1. / 3) : 0;
  } else {       // 5-level, MC422L:1004-1031
    const double c2 = __dadd_rn(n1, n3), c3 = __dadd_rn(c2, n4);
    up = (ua < n1) ? 1 : (ua < c2) ? (ub < 1. / 3) : (ua < c3) ? (ub < 2. / 3) : 0;
  }
  tagged[(size_t)b * N + i] = up;
  if (up) atomicAdd(&count[b], 1);
}
void launch_tag(const double* psi, int S, int N, int ld, int B, int traj0, uint64_t seed, uint64_t substep,
                const double* forced_u, int* tagged, int* count, cudaStream_t s) {
  cudaMemsetAsync(count, 0, sizeof(int) * B, s);
  dim3 grid((N + 255) / 256, B);
  k_tag<<<grid, 256, 0, s>>>(psi, S, N, ld, traj0, seed, substep, forced_u, tagged, count);
}

// ------------------------------------------------------------------------------------------------------------
// FZ-family leap-frog pieces (FZ408L:317-369): optional kick V += DTV*F, optional drift R += DT*V (+ DT*DT*F while
// t <= 0), single wrap into [0,L]; every operation rounded as the reference's expression tree.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_lf(LFArgs a) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)a.nrows * a.B * 3) return;
  int b = (int)(g / (3LL * a.nrows));
  int rem = (int)(g % (3LL * a.nrows));
  int c = rem / a.nrows, i = a.row0 + rem % a.nrows;
  size_t idx = ((size_t)b * 3 + c) * a.ld + i;
  double v = a.V[idx];
  const double f = a.F[idx];
  if (a.kick) { v = __dadd_rn(v, __dmul_rn(a.DTV, f)); a.V[idx] = v; }  // FZ408L:364-366
  if (a.drift) {
    double r = a.R[idx];
    if (a.first) r = __dadd_rn(r, __dadd_rn(__dmul_rn(a.DT, v), __dmul_rn(__dmul_rn(a.DT, a.DT), f)));  // FZ408L:336-338
    else r = __dadd_rn(r, __dmul_rn(a.DT, v));                                                          // FZ408L:325-327
    if (r < 0) r = __dadd_rn(r, a.L);
    if (r > a.L) r = __dadd_rn(r, -a.L);
    a.R[idx] = r;
    a.Rfix[idx] = to_fixed(r, a.invL, a.invL_lo);
  }
}
void launch_lf(const LFArgs& a, cudaStream_t s) {
  long long n = 3LL * a.nrows * a.B;
  k_lf<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
}

}  // namespace mdqt

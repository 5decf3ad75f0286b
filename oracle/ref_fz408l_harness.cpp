// oracle/ref_fz408l_harness.cpp -- TEST INFRASTRUCTURE ONLY (see ref_su_harness.cpp for the rules).
//
// Hijack include of the UNMODIFIED reference randomFrozenStartTag408Linear.cpp (FZ408L): the frozen-start pump-window
// program. Exposes its leap-frog step() (FZ408L:317-390, forces() inside step_V and inside the 2nd-order start),
// its 7-level qstep() (FZ408L:396-598, advances t), measureSpinUps() (FZ408L:600-659), Zfunc() (FZ408L:938-961) and
// init() (FZ408L:251-310). Control is taken back at the srand48 call (FZ408L:1028), after cs[]/gs[] are built.
// Use with OMP_NUM_THREADS=1 (the force loop races).
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <setjmp.h>
#include <sys/stat.h>
#include <omp.h>
#include <iostream>
#include <complex>
#include <random>
#include <armadillo>

static jmp_buf g_env;
static const double* g_uq = 0;
static long g_un = 0, g_ui = 0;
static double oracle_u() {
  if (g_ui < g_un) return g_uq[g_ui++];
  return drand48();
}
static void oracle_hook() { longjmp(g_env, 1); }
static void oracle_seed(long s) { srand48(s); }

#define main ref_main
#define srand48(x) oracle_hook()
#define drand48() oracle_u()
#define mkdir(a, b) (0)
#include "randomFrozenStartTag408Linear.cpp"
#undef main
#undef srand48
#undef drand48
#undef mkdir

extern "C" {
int ref_fz_capacity() { return N0 + 1000; }
int ref_fz_setup(double detuning_, double Om_, const char* scratch) {
  detuning = detuning_; Om = Om_;
  omp_set_num_threads(1);
  ::mkdir(scratch, 0777);
  strcpy(saveDirectory, scratch);
  static char a0[] = "ref", a1[] = "1";
  char* av[] = {a0, a1, 0};
  if (setjmp(g_env) == 0) { ref_main(2, av); return 1; }
  strcpy(saveDirectory, scratch);
  lDeb = 1. / sqrt(3. * Ge);
  L = pow(N0 * 4. * M_PI / 3., 0.333333333);
  t = 0;
  return 0;
}
// out = {L, lDeb, quantumTimestep, gamToEinsteinFreq, plasVelToQuantVel, ratio, decayRatio, tpump, tendV0, TIMESTEP}
void ref_fz_get_consts(double* out) {
  out[0] = L; out[1] = lDeb; out[2] = quantumTimestep; out[3] = gamToEinsteinFreq; out[4] = plasVelToQuantVel;
  out[5] = plasmaToQuantumTimestepRatio; out[6] = decayRatio; out[7] = tpump; out[8] = tendV0; out[9] = TIMESTEP;
}
int ref_fz_init(long seed) { oracle_seed(seed); t = 0; init(); return (int)N; }
int ref_fz_get_N() { return (int)N; }
double ref_fz_get_t() { return t; }
void ref_fz_set_t(double t_) { t = t_; }
void ref_fz_set_state(int n, const double* R_, const double* V_, const double* psi) {
  N = n;
  for (int i = 0; i < n; i++) {
    for (int c = 0; c < 3; c++) {
      if (R_) R[c][i] = R_[c * n + i];
      if (V_) V[c][i] = V_[c * n + i];
    }
    if (psi) {
      cx_mat w = cx_mat(mat(7, 1, fill::zeros), mat(7, 1, fill::zeros));
      for (int k = 0; k < 7; k++) w(k, 0) = std::complex<double>(psi[(i * 7 + k) * 2], psi[(i * 7 + k) * 2 + 1]);
      wvFns[i] = w;
    }
  }
}
void ref_fz_get_state(double* R_, double* V_, double* F_, double* psi) {
  int n = (int)N;
  for (int i = 0; i < n; i++) {
    for (int c = 0; c < 3; c++) {
      if (R_) R_[c * n + i] = R[c][i];
      if (V_) V_[c * n + i] = V[c][i];
      if (F_) F_[c * n + i] = F[c][i];
    }
    if (psi)
      for (int k = 0; k < 7; k++) {
        psi[(i * 7 + k) * 2] = wvFns[i](k, 0).real();
        psi[(i * 7 + k) * 2 + 1] = wvFns[i](k, 0).imag();
      }
  }
}
void ref_fz_set_uniforms(const double* u, long nu) { g_uq = u; g_un = nu; g_ui = 0; }
long ref_fz_uniforms_used() { return g_ui; }
void ref_fz_step() { step(); }
void ref_fz_qstep() { qstep(); }
int ref_fz_measure(int* out) {
  measureSpinUps();
  for (unsigned i = 0; i < N; i++) out[i] = SpinUpList[i];
  return NSpinUp;
}
double ref_fz_zfunc(int c1V) { Zfunc(c1V); return VAF; }
int ref_fz_get_c0() { return c0; }
void ref_fz_set_c0(int c) { c0 = c; }
// The body of main()'s time loop (FZ408L:1040-1072) re-typed around the reference's OWN functions and globals, with the
// compile-time tmax / tstartV0 and the derived tendV0 as parameters, and the file-writing output()/printVAF() calls left
// out (recordedSpinUps starts at 0: a new run). Returns the number of loop iterations; vaf[0], vaf[1] = Zfunc values at the
// measurement and at the last sampling point (NaN if none); spin = SpinUpList.
long ref_fz_run_loop(double tmax_, double tstart_, double tend_, int sampleFreq_, double* vaf, int* spin, int* nspin) {
  int timeStepCounter = plasmaToQuantumTimestepRatio;
  int recorded = 0;
  long iters = 0;
  vaf[0] = vaf[1] = NAN;
  *nspin = -1;
  while (t <= tmax_ + 0.0009) {
    if (recorded == 0 && t >= tend_) {
      measureSpinUps();
      recorded = 1;
      for (unsigned i = 0; i < N; i++) spin[i] = SpinUpList[i];
      *nspin = NSpinUp;
      Zfunc(0);
      vaf[0] = VAF;
    }
    if ((c0 + 1) % sampleFreq_ == 0 && timeStepCounter == 1 && recorded == 1) {
      Zfunc(1);
      vaf[1] = VAF;
    }
    if (timeStepCounter == plasmaToQuantumTimestepRatio) {
      step();
      c0++;
      timeStepCounter = 0;
    }
    if (t < tend_ && t > tstart_) qstep();
    else t += quantumTimestep;
    timeStepCounter++;
    iters++;
  }
  return iters;
}
}

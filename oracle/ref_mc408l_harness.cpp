// oracle/ref_mc408l_harness.cpp -- TEST INFRASTRUCTURE ONLY (see ref_su_harness.cpp for the rules).
//
// Hijack include of the UNMODIFIED reference MonteCarloFollowedByQTTagging408Linear.cpp (MC408L): its 7-level
// qstep() (MC408L:555-756) with the channel tables built inside its main (MC408L:1171-1190), plus the MD core
// (calculateAccelerations MC408L:437-498, MDStep). main has no srand48 call to hook, so
// control is taken back at the first `cout << k` of the Monte-Carlo loop (MC408L:1205), i.e. after the tables,
// init() and calculatePotentialEnergyForParticles() have run. N=4096 is a compile-time constant, so qstep()
// always sweeps 4096 ions serially; uniforms are injected as ONE sequential stream in that serial order.
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
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

#define main ref_main
#define drand48() oracle_u()
#define cout (oracle_hook(), std::cout)
#include "MonteCarloFollowedByQTTagging408Linear.cpp"
#undef main
#undef drand48
#undef cout

extern "C" {
int ref_mc_N() { return N; }
// p = {detuning, Om}
int ref_mc_setup(const double* p, const char* scratch) {
  detuning = p[0]; Om = p[1];
  ::mkdir(scratch, 0777);
  strcpy(saveDirectory, scratch);
  static char a0[] = "ref", a1[] = "1";
  char* av[] = {a0, a1, 0};
  if (setjmp(g_env) == 0) { ref_main(2, av); return 1; }
  return 0;
}
// out = {L, rCut, kappa, Gamma, n, timeStep, g2E, ratio, quantumTimestep, pv2qv, decayRatio, pumpMDTimeSteps}
void ref_mc_get_consts(double* out) {
  out[0] = L; out[1] = rCut; out[2] = kappa; out[3] = Gamma; out[4] = n; out[5] = timeStep;
  out[6] = gamToEinsteinFreq; out[7] = plasmaToQuantumTimestepRatio; out[8] = quantumTimestep;
  out[9] = plasVelToQuantVel; out[10] = decayRatio; out[11] = pumpMDTimeSteps;
}
void ref_mc_seed(unsigned s) { rng.seed(s); velocityDistribution.reset(); uni.reset(); }
void ref_mc_set_controls(double collFreq) { collisionFreq = collFreq; }
void ref_mc_set_state(const double* R_, const double* V_, const double* A_, const double* psi) {
  for (int i = 0; i < N; i++) {
    for (int c = 0; c < 3; c++) {
      if (R_) R[c][i] = R_[c * N + i];
      if (V_) V[c][i] = V_[c * N + i];
      if (A_) A[c][i] = A_[c * N + i];
    }
    if (psi) {
      cx_mat w = cx_mat(mat(7, 1, fill::zeros), mat(7, 1, fill::zeros));
      for (int k = 0; k < 7; k++) w(k, 0) = std::complex<double>(psi[(i * 7 + k) * 2], psi[(i * 7 + k) * 2 + 1]);
      wvFns[i] = w;
    }
  }
}
void ref_mc_get_state(double* R_, double* V_, double* A_, double* psi) {
  for (int i = 0; i < N; i++) {
    for (int c = 0; c < 3; c++) {
      if (R_) R_[c * N + i] = R[c][i];
      if (V_) V_[c * N + i] = V[c][i];
      if (A_) A_[c * N + i] = A[c][i];
    }
    if (psi)
      for (int k = 0; k < 7; k++) {
        psi[(i * 7 + k) * 2] = wvFns[i](k, 0).real();
        psi[(i * 7 + k) * 2 + 1] = wvFns[i](k, 0).imag();
      }
  }
}
void ref_mc_set_uniforms(const double* u, long nu) { g_uq = u; g_un = nu; g_ui = 0; }
long ref_mc_uniforms_used() { return g_ui; }
void ref_mc_qstep() { qstep(); }
void ref_mc_tag(int* out) { tagParticles(); for (int i = 0; i < N; i++) out[i] = tagged[i] ? 1 : 0; }  // MC408L:1022-1067
void ref_mc_accelerations() { calculateAccelerations(0); }
void ref_mc_mdstep() { MDStep(0); }
// init() again under a chosen seed (its mt19937 draws and, for the wavefunctions, drand48), then main()'s stages 4-6
// (MC408L:1211-1244) with run-time step counts (the reference's are compile-time constants): every call the reference's own function in
// the reference's order -- collisional MD, collisionFreq = 0, pump {ratio x qstep(); MDStep(k)}, tagParticles(), then per recorded
// step recordTaggedParticleMoments(k); k % 100 == 0 -> recordPairPairCorr(k); recordTemperature(); MDStep(k);
// recordVelsForAutocorrelations(k). Files go to `dir` (trailing '/'). Returns the number of tagged ions.
void ref_mc_init(unsigned s) { rng.seed(s); velocityDistribution.reset(); uni.reset(); srand48(s); init(); memset(A, 0, sizeof(A)); }
int ref_mc_run_stages(const char* dir, int npre, int npump, int nrec, double collFreq) {
  ::mkdir(dir, 0777);
  strcpy(saveDirectory, dir);
  collisionFreq = collFreq;
  for (int k = 0; k < npre; k++) MDStep(k);
  collisionFreq = 0;
  for (int k = 0; k < npump; k++) {
    for (int l = 0; l < plasmaToQuantumTimestepRatio; l++) qstep();
    MDStep(k);
  }
  tagParticles();
  int ntag = 0;
  for (int i = 0; i < N; i++) ntag += tagged[i] ? 1 : 0;
  for (int k = 0; k < nrec; k++) {
    recordTaggedParticleMoments(k);
    if (k % 100 == 0) recordPairPairCorr(k);
    recordTemperature();
    MDStep(k);
    recordVelsForAutocorrelations(k);
  }
  return ntag;
}
}

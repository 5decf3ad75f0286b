// oracle/ref_md_harness.cpp -- TEST INFRASTRUCTURE ONLY (see ref_su_harness.cpp for the rules).
//
// Hijack include of the UNMODIFIED reference MonteCarloFollowedByMDAndTempAnisotropy.cpp (MD; no Armadillo):
// its calculateAccelerations() (MD:387-448), stepPositions() (MD:452-467), stepVelocities() (MD:469-502)
// and MDStep() (MD:504-511) become callable on controlled inputs. N=4096, kappa=0.5, L, rCut are compile-time
// constants of the reference (MD:66-74) and are therefore fixed here too. Use OMP_NUM_THREADS=1 (Q1, Q3).
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <sys/stat.h>
#include <omp.h>
#include <iostream>
#include <random>

#define main ref_main
#define mkdir(a, b) (0)
#include "MonteCarloFollowedByMDAndTempAnisotropy.cpp"
#undef main
#undef mkdir

extern "C" {
int ref_md_N() { return N; }
// out = {L, rCut, kappa, Gamma, n, timeStep, beta}
void ref_md_get_consts(double* out) {
  out[0] = L; out[1] = rCut; out[2] = kappa; out[3] = Gamma; out[4] = n; out[5] = timeStep; out[6] = beta;
}
void ref_md_seed(unsigned s) { rng.seed(s); velocityDistribution.reset(); uni.reset(); }
void ref_md_init() { init(); }
void ref_md_set_controls(double collFreq, int laser, int oneAxis) {
  collisionFreq = collFreq; addLaserForce = laser; applyForceAlongOneAxisOnly = (oneAxis != 0);
}
void ref_md_set_state(const double* R_, const double* V_, const double* A_) {
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < N; i++) {
      if (R_) R[c][i] = R_[c * N + i];
      if (V_) V[c][i] = V_[c * N + i];
      if (A_) A[c][i] = A_[c * N + i];
    }
}
void ref_md_get_state(double* R_, double* V_, double* A_) {
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < N; i++) {
      if (R_) R_[c * N + i] = R[c][i];
      if (V_) V_[c * N + i] = V[c][i];
      if (A_) A_[c * N + i] = A[c][i];
    }
}
void ref_md_accelerations() { calculateAccelerations(0); }
void ref_md_step_positions() { stepPositions(); }
void ref_md_mdstep() { MDStep(0); }
}

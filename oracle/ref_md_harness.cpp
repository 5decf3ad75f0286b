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
// recordPairPairCorr(stepNum) (MD:584-652) writes <saveDirectory>pairPairCorrStepNum<k>.dat: run it into `scratch`
// (must end with '/') and read the two columns back; returns the number of rows
int ref_md_pair_correlation(const char* scratch, double* r_out, double* g_out, int cap) {
  ::mkdir(scratch, 0777);
  strcpy(saveDirectory, scratch);
  recordPairPairCorr(7);
  char fn[512];
  snprintf(fn, sizeof fn, "%spairPairCorrStepNum7.dat", scratch);
  FILE* f = fopen(fn, "r");
  if (!f) return -1;
  int k = 0;
  while (k < cap && fscanf(f, "%lg %lg", &r_out[k], &g_out[k]) == 2) k++;
  fclose(f);
  return k;
}
// main()'s stages 5, 7 and 8 (MD:1090-1165) with run-time step counts (the reference's are compile-time constants), every
// call the reference's own function in the reference's order, files written by its own recorders into `dir` (trailing '/'):
//   stage 5: collisionFreq = 0; tagParticles(); k < nrec: recordTaggedParticleMoments(k); k % 100 == 0 -> recordPairPairCorr(k);
//            recordTemperature(); MDStep(k); recordVelsForAutocorrelations(k)
//   stage 7: anisotropizeVelocities(); k < ninst: recordTempForEachAxis(Instantaneous, k); MDStep(k)   (its cout prints muted)
//   stage 8: addLaserForce = 1; k < nest: recordTempForEachAxis(DuringForcePeriod, k); MDStep(k);
//            addLaserForce = 0; k < nrelax: recordTempForEachAxis(AfterForcePeriod, k); MDStep(k)
void ref_md_run_stages(const char* dir, int nrec, int ninst, int nest, int nrelax) {
  ::mkdir(dir, 0777);
  strcpy(saveDirectory, dir);
  collisionFreq = 0;
  tagParticles();
  for (int k = 0; k < nrec; k++) {
    recordTaggedParticleMoments(k);
    if (k % 100 == 0) recordPairPairCorr(k);
    recordTemperature();
    MDStep(k);
    recordVelsForAutocorrelations(k);
  }
  std::streambuf* old = std::cout.rdbuf(nullptr);  // anisotropizeVelocities() prints every velocity twice
  anisotropizeVelocities();
  std::cout.rdbuf(old);
  char fileName[256];
  strcpy(fileName, saveDirectory); strcat(fileName, "TemperaturesAlongAxesInstantaneous.dat");
  for (int k = 0; k < ninst; k++) { recordTempForEachAxis(fileName, k); MDStep(k); }
  addLaserForce = 1; collisionFreq = 0;
  strcpy(fileName, saveDirectory); strcat(fileName, "TemperaturesAlongAxesDuringForcePeriod.dat");
  for (int k = 0; k < nest; k++) { recordTempForEachAxis(fileName, k); MDStep(k); }
  addLaserForce = 0;
  strcpy(fileName, saveDirectory); strcat(fileName, "TemperaturesAlongAxesAfterForcePeriod.dat");
  for (int k = 0; k < nrelax; k++) { recordTempForEachAxis(fileName, k); MDStep(k); }
}
void ref_md_get_tags(unsigned char* out) {
  for (int i = 0; i < N; i++) out[i] = (taggedOne[i] ? 1 : 0) | (taggedTwo[i] ? 2 : 0) | (taggedThree[i] ? 4 : 0) | (taggedFour[i] ? 8 : 0);
}
int ref_md_autocorr_steps() { return numVelAutoCorrsSteps; }
double ref_md_pair_step() { return pairPairStep; }
double ref_md_pair_max() { return pairPairMax; }
// vStore[c][i][:] = v[c][i][:] for i < n, zero beyond (MD:121)
void ref_md_set_vstore(int n, const double* v) {
  const int T = numVelAutoCorrsSteps;
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < T; j++) vStore[c][i][j] = (i < n) ? v[((size_t)c * n + i) * T + j] : 0.0;
}
// which: 1 recordVAF, 2 recordLongViscAutoCorr, 3 recordVCubeAutoCorr, 4 recordVFourthAutoCorr (MD:654-823); the files go to scratch
void ref_md_autocorr(int which, const char* scratch, double* out) {
  ::mkdir(scratch, 0777);
  strcpy(saveDirectory, scratch);
  const double* src = 0;
  if (which == 1) { recordVAF(); src = VAF; }
  else if (which == 2) { recordLongViscAutoCorr(); src = longViscAutoCorr; }
  else if (which == 3) { recordVCubeAutoCorr(); src = vCubeAutoCorr; }
  else { recordVFourthAutoCorr(); src = vFourthAutoCorr; }
  for (int j = 0; j < numVelAutoCorrsSteps; j++) out[j] = src[j];
}
}

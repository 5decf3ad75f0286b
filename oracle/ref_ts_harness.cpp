// oracle/ref_ts_harness.cpp -- TEST INFRASTRUCTURE ONLY (see ref_su_harness.cpp for the rules).
//
// Hijack include of the UNMODIFIED reference laserCoolNoPlasmaThreeState.cpp (TS): its 3-level qstep() (TS:140-293)
// with cs[]/gs[] built inside its main (TS:379-382). Control is taken back at the srand48 call (TS:384). N0 = 1000 is
// a compile-time constant. Uniforms are injected as ONE sequential stream in ion order -- use one OpenMP thread
// (ref_ts_setup sets it), the reference's shared drand48 is racy otherwise.
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

#define main ref_main
#define srand48(x) oracle_hook()
#define drand48() oracle_u()
#define mkdir(a, b) (0)
#include "laserCoolNoPlasmaThreeState.cpp"
#undef main
#undef srand48
#undef drand48
#undef mkdir

extern "C" {
int ref_ts_N() { return N0; }
int ref_ts_setup(double detuning_, double Om_, double dt_, int applyForce_) {
  detuning = detuning_; Om = Om_; applyForce = applyForce_ != 0;
  omp_set_num_threads(1);
  static char a0[] = "ref", a1[] = "1";
  char* av[] = {a0, a1, 0};
  if (setjmp(g_env) == 0) { ref_main(2, av); return 1; }
  dt = dt_; t = 0;  // main sets dt = 0.01 after the hook (TS:390)
  return 0;
}
double ref_ts_vkick() { return vKick; }
void ref_ts_set_state(const double* Vx, const double* psi, const double* tp) {
  for (int i = 0; i < N0; i++) {
    if (Vx) V[0][i] = Vx[i];
    if (tp) tPart[i] = tp[i];
    if (psi) {
      cx_mat w = cx_mat(mat(3, 1, fill::zeros), mat(3, 1, fill::zeros));
      for (int k = 0; k < 3; k++) w(k, 0) = std::complex<double>(psi[(i * 3 + k) * 2], psi[(i * 3 + k) * 2 + 1]);
      wvFns[i] = w;
    }
  }
}
void ref_ts_get_state(double* Vx, double* psi, double* tp) {
  for (int i = 0; i < N0; i++) {
    if (Vx) Vx[i] = V[0][i];
    if (tp) tp[i] = tPart[i];
    if (psi)
      for (int k = 0; k < 3; k++) {
        psi[(i * 3 + k) * 2] = wvFns[i](k, 0).real();
        psi[(i * 3 + k) * 2 + 1] = wvFns[i](k, 0).imag();
      }
  }
}
void ref_ts_set_uniforms(const double* u, long nu) { g_uq = u; g_un = nu; g_ui = 0; }
long ref_ts_uniforms_used() { return g_ui; }
void ref_ts_qstep() { qstep(); }
}

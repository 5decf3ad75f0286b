// examples/su_dropin.cpp -- the drop-in boundary, EXECUTED: the reference program
// laserCoolingPlusExpansionMDQTSpeedUp.cpp (SU), unmodified and #include'd from where it lies, runs its OWN main() -- directory
// setup, operator tables, init(), the while (t <= tmax + 0.0009) loop, output(), writeConditions() -- while the hot
// functions it calls are routed through the C ABI of include/mdqt.h, exactly the wiring INTEGRATION.md section 2 describes:
//
//     forces()       SU:192   ->  mdqt_forces
//     step()         SU:418   ->  (fused into the next call)
//     qstep()        SU:438   ->  mdqt_substeps(gpu, 1)   = { step(); qstep(); } on the device, then t <- mdqt_get_time
//     Epotential()   SU:244   ->  mdqt_epot                (also the call inside init() and output())
//     output()       SU:917   ->  download the device state into the reference's globals, then the reference's own output()
//
// How the calls are re-routed without touching the source: `#define forces(...) forces_##__VA_ARGS__()` turns the
// reference's DEFINITION `void forces(void)` into `void forces_void()` (its CPU code, kept but unused) and every CALL
// `forces()` into `forces_()`, which is defined here. Built by oracle/Makefile into oracle/_ref/su_dropin (the reference
// sources exist only in the build container); needs an Armadillo header -- the real one, or oracle/arma_shim.
//
//   su_dropin <job> [--tmax x] [--seed n] [--Om x] [--OmDP x]
//
// tmax is a compile-time #define of the reference (SU:63, 30); a shorter run is obtained by pushing the global `t` past it
// once the requested end time is reached (the loop condition is the only reader).
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <sys/stat.h>
#include <omp.h>
#include <iostream>
#include <complex>
#include <random>
#include <armadillo>
#include "mdqt.h"

static void forces_();
static void step_();
static void qstep_();
static void Epotential_();
static void output_();
static long g_seed = -1;
static long dropin_seed(long from_reference) { return g_seed >= 0 ? g_seed : from_reference; }

#define forces(...) forces_##__VA_ARGS__()
#define step(...) step_##__VA_ARGS__()
#define qstep(...) qstep_##__VA_ARGS__()
#define Epotential(...) Epotential_##__VA_ARGS__()
#define output(...) output_##__VA_ARGS__()
#define main ref_main
#define srand48(x) srand48(dropin_seed((long)(x)))
#include "laserCoolingPlusExpansionMDQTSpeedUp.cpp"
#undef forces
#undef step
#undef qstep
#undef Epotential
#undef output
#undef main
#undef srand48

// ---- the stub of INTEGRATION.md section 2 -----------------------------------------------------------------------------------
static mdqt_handle* gpu = NULL;
static double psi_host[(N0 + 1000) * 12 * 2];  // wvFns[] marshalled as [ion][state][re,im]
static double g_tend = tmax;                   // requested end time (<= the reference's compile-time tmax)

static void gpu_check(int rc) {
  if (rc) { fprintf(stderr, "su_dropin: mdqt: %s\n", mdqt_last_error()); exit(1); }
}
static void gpu_upload(void) {  // globals -> device
  for (unsigned i = 0; i < N; i++)
    for (int k = 0; k < 12; k++) { psi_host[(i * 12 + k) * 2] = wvFns[i](k, 0).real(); psi_host[(i * 12 + k) * 2 + 1] = wvFns[i](k, 0).imag(); }
  gpu_check(mdqt_upload_state(gpu, &R[0][0], &V[0][0], psi_host, tPart, N0 + 1000));  // ld = N0+1000 (SU:126)
  gpu_check(mdqt_set_time(gpu, t, (uint64_t)llround(t / quantumTimestep)));
}
static void gpu_download(void) {  // device -> globals
  gpu_check(mdqt_download_state(gpu, &R[0][0], &V[0][0], psi_host, tPart, N0 + 1000));
  for (unsigned i = 0; i < N; i++)
    for (int k = 0; k < 12; k++) wvFns[i](k, 0) = std::complex<double>(psi_host[(i * 12 + k) * 2], psi_host[(i * 12 + k) * 2 + 1]);
}
static void gpu_ensure(void) {  // first hot call after init() / readConditions(c0): N and the globals are final
  if (gpu) return;
  mdqt_params p;
  gpu_check(mdqt_params_su(&p, Ge, density, sig0, Te, fracOfSig, detuning, detuningDP, Om, OmDP, N0, (int)N));
  p.traj0 = (int)job;
  p.seed = (uint64_t)dropin_seed((long)((unsigned)time(NULL) + job));  // was srand48(time(NULL)+job), SU:1219
  p.renormalize = reNormalizewvFns;
  gpu_check(mdqt_create(&p, &gpu));
  gpu_upload();
}

static void forces_() { gpu_ensure(); gpu_check(mdqt_forces(gpu)); }
static void step_() {}  // step() and qstep() are one fused device call: see qstep_()
static void qstep_() {
  gpu_ensure();
  gpu_check(mdqt_substeps(gpu, 1));
  gpu_check(mdqt_get_time(gpu, &t, NULL));
  if (!(t <= g_tend + 0.0009)) {  // the run ends here: bring the state home for writeConditions(c0) and leave the loop
    gpu_download();
    t += 2.0 * tmax + 1.0;
  }
}
static void Epotential_() { gpu_ensure(); gpu_check(mdqt_epot(gpu, &Epot)); }
static void output_() {
  gpu_ensure();
  gpu_download();
  output_void();  // the reference's own output(): host observables from the downloaded globals; its Epotential() call is routed above
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: su_dropin <job> [--tmax x] [--seed n] [--Om x] [--OmDP x]\n"); return 2; }
  for (int i = 2; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--tmax")) g_tend = atof(argv[i + 1]);
    else if (!strcmp(argv[i], "--seed")) g_seed = atol(argv[i + 1]);
    else if (!strcmp(argv[i], "--Om")) Om = atof(argv[i + 1]);
    else if (!strcmp(argv[i], "--OmDP")) OmDP = atof(argv[i + 1]);
    else { fprintf(stderr, "su_dropin: bad option %s\n", argv[i]); return 2; }
  }
  char* av[] = {argv[0], argv[1], 0};
  int rc = ref_main(2, av);  // the reference's main(), SU:1139-1383
  if (gpu) mdqt_destroy(gpu);
  return rc;
}

/* examples/abi_client.c -- a plain C99 caller of the C ABI (include/mdqt.h): what a maintainer of the reference programs
 * links against. Builds a small frozen-start system, runs `nsteps` iterations of the main-loop body
 * { forces(); ratio x { step(); qstep(); } } (SU:1369-1378) on the GPU and prints observables.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/abi_client.c -Lmdqtplasmasims_b200 -lmdqt_b200 -Wl,-rpath,$PWD/mdqtplasmasims_b200 -lm
 *   ./a.out [n_ions] [nsteps]        exit code 0 = ran; 3 = no CUDA device (the library has no CPU fallback)
 */
#include "mdqt.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static unsigned long long lcg = 88172645463325252ULL;
static double urand(void) { /* xorshift64: a self-contained stream for the example's initial state */
  lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17;
  return (double)(lcg >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000, nsteps = argc > 2 ? atoi(argv[2]) : 5, ld = n + 1000; /* ld as SU:126 */
  mdqt_params p;
  if (mdqt_params_su(&p, 0.1, 2.0, 4.0, 19.0, 0.0, -1.0, 1.0, 1.0, 1.0, n, n)) { fprintf(stderr, "%s\n", mdqt_last_error()); return 1; }
  p.seed = 2024; p.traj0 = 1;
  double* R = calloc((size_t)3 * ld, sizeof(double));
  double* V = calloc((size_t)3 * ld, sizeof(double));
  double* psi = calloc((size_t)n * 24, sizeof(double));
  double* tPart = calloc((size_t)n, sizeof(double));
  for (int i = 0; i < n; i++) {
    for (int c = 0; c < 3; c++) R[c * ld + i] = p.L * urand();
    const double r1 = urand(), r2 = urand(); /* random S-manifold superposition, as init() builds it (SU:317-332) */
    psi[i * 24 + 0] = sqrt(r1);
    psi[i * 24 + 2] = sqrt(1 - r1) * sqrt(r2);
    psi[i * 24 + 3] = sqrt(1 - r1) * sqrt(1 - r2);
  }
  mdqt_handle* h = NULL;
  int rc = mdqt_create(&p, &h);
  if (rc == MDQT_ENODEVICE) { fprintf(stderr, "abi_client: %s\n", mdqt_last_error()); return 3; }
  if (rc) { fprintf(stderr, "abi_client: %s\n", mdqt_last_error()); return 1; }
  if (mdqt_upload_state(h, R, V, psi, tPart, ld) || mdqt_set_time(h, 0.0, 0) || mdqt_md_steps(h, nsteps) ||
      mdqt_download_state(h, R, V, psi, tPart, ld)) { fprintf(stderr, "abi_client: %s\n", mdqt_last_error()); return 1; }
  mdqt_diag d;
  if (mdqt_diagnostics(h, &d)) { fprintf(stderr, "abi_client: %s\n", mdqt_last_error()); return 1; }
  double norm = 0.0;
  for (int k = 0; k < n * 24; k++) norm += psi[k] * psi[k];
  printf("t=%.17g ekin_x=%.17g ekin_y=%.17g ekin_z=%.17g epot=%.17g vx_avg=%.17g norm=%.17g\n", d.t, d.ekin_x, d.ekin_y, d.ekin_z,
         d.epot, d.vx_avg, norm / n);
  mdqt_destroy(h);
  free(R); free(V); free(psi); free(tPart);
  return 0;
}

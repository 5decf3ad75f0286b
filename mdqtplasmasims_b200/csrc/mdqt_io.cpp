// mdqt_io.cpp -- host-side driver pieces around the hot path (include/mdqt_io.h): directory naming, init(),
// restart files, output() files and the main-loop schedule of the reference
// laserCoolingPlusExpansionMDQTSpeedUp.cpp, with the same format strings so files are byte-compatible
// (SURVEY.md App. B). No CUDA here.
#include "../../include/mdqt_io.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <string>
#include <vector>

namespace {
const int kStates = 12;
const int kBins = 2001;
// (unsigned)(negative double) is undefined behaviour in the reference (SU:1153); x86-64 gcc yields the wrapped
// two's-complement value, which "%d" prints as the negative number. Reproduce that.
int as_printed(double x) { return (int)(long long)x; }
std::string join(const char* dir, const char* name) { return std::string(dir) + name; }
}  // namespace

extern "C" {

int mdqt_io_dirname(char* out, int cap, const char* saveDirectory, double Ge, double density, double sig0, double Te,
                    double fracOfSig, double detuning, double detuningDP, double Om, double OmDP, int N0, unsigned job,
                    int create) {
  char namebuf[256], namebuf2[64];
  snprintf(namebuf, sizeof(namebuf), "Ge%dDensity%dE+11Sig0%dTe%dSigFrac%dDetSP%dDetDP%dOmSP%dOmDP%dNumIons%d",
           as_printed(100 * Ge), as_printed(density * 1000), as_printed(10 * sig0), as_printed(Te), as_printed(fracOfSig * 100),
           as_printed(detuning * 100), as_printed(detuningDP * 100), as_printed(Om * 100), as_printed(OmDP * 100), N0);
  snprintf(namebuf2, sizeof(namebuf2), "/job%d/", (int)job);
  std::string top(saveDirectory), mid = top + namebuf, full = mid + namebuf2;
  if ((int)full.size() + 1 > cap) return -1;
  if (create) {
    mkdir(top.c_str(), 0777);  // ACCESSPERMS
    mkdir(mid.c_str(), 0777);
    mkdir(full.c_str(), 0777);
  }
  strcpy(out, full.c_str());
  return 0;
}

int mdqt_io_init_su(long seed, int N0, double Ge, int ld, double* R, double* V, double* psi, double* tPart, double* L_out,
                    double* lDeb_out) {
  const double lDeb = 1. / sqrt(3. * Ge);                    // SU:295
  const double L = pow(N0 * 4. * M_PI / 3., 0.333333333);    // SU:297
  const double N9L = (unsigned)(9. * 9. * 9. * (L * L * L) * 3. / (4. * M_PI));  // SU:299
  // srand48(seed) + drand48() on a LOCAL generator state (erand48: the same 48-bit LCG and the same numbers as the
  // reference's global stream), so that the jobs of an ensemble can be initialised from several threads at once
  unsigned short xs[3] = {0x330E, (unsigned short)(seed & 0xFFFF), (unsigned short)((seed >> 16) & 0xFFFF)};
  auto drand48 = [&xs]() { return erand48(xs); };
  int N = 0;
  for (int i = 0; i < N9L; i++) {
    double x = 9. * L * drand48() - 4. * L;  // SU:305-307
    double y = 9. * L * drand48() - 4. * L;
    double z = 9. * L * drand48() - 4. * L;
    if (x <= L && y <= L && z <= L && x > 0 && y > 0 && z > 0) {
      if (N >= ld) return -1;
      R[N] = x; R[ld + N] = y; R[2 * ld + N] = z;
      V[N] = 0.; V[ld + N] = 0.; V[2 * ld + N] = 0.;
      double rand1 = drand48(), rand2 = drand48(), rand3 = drand48();  // SU:317-319
      double sign = 1;
      if (rand3 < 0.5) sign = -1;
      double rand4 = drand48();
      double sign2 = 1;
      if (rand4 < 0.5) sign2 = -1;
      double* w = psi + (size_t)N * kStates * 2;
      for (int k = 0; k < kStates * 2; k++) w[k] = 0.0;
      w[0] = sqrt(rand1);                                   // S mJ=-1/2 (SU:329)
      w[2] = sign2 * sqrt(1 - rand1) * sqrt(rand2);         // S mJ=+1/2, real part (SU:330)
      w[3] = sign * sqrt(1 - rand1) * sqrt(1 - rand2);      // imaginary part (SU:331)
      tPart[N] = 0;
      N++;
    }
  }
  if (L_out) *L_out = L;
  if (lDeb_out) *lDeb_out = lDeb;
  return N;
}

int mdqt_io_write_conditions(const char* dir, int c0, int N, unsigned counter, const double* R, const double* V,
                             const double* psi, int ld, const double* vholder) {
  char buffer[256];
  FILE* fa;
  snprintf(buffer, sizeof(buffer), "ions_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "w"))) return -2;
  fprintf(fa, "%i\t%i", N, counter);  // SU:737
  fclose(fa);
  snprintf(buffer, sizeof(buffer), "conditions_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "w"))) return -2;
  for (int i = 0; i < N; i++)
    fprintf(fa, "%lg\t%lg\t%lg\t%lg\t%lg\t%lg\t\n", R[i], R[ld + i], R[2 * ld + i], V[i], V[ld + i], V[2 * ld + i]);  // SU:747
  fclose(fa);
  for (int c2V = 0; c2V < MDQT_NUM_VINTERVALS; c2V++) {
    snprintf(buffer, sizeof(buffer), "VZERO_timestep%06d_interval%d.dat", c0, c2V);
    if (!(fa = fopen(join(dir, buffer).c_str(), "w"))) return -2;
    for (int i = 0; i < N; i++) {
      double a = 0, b = 0, c = 0;
      if (vholder) {
        a = vholder[((size_t)0 * MDQT_NUM_VINTERVALS + c2V) * ld + i];
        b = vholder[((size_t)1 * MDQT_NUM_VINTERVALS + c2V) * ld + i];
        c = vholder[((size_t)2 * MDQT_NUM_VINTERVALS + c2V) * ld + i];
      }
      fprintf(fa, "%lg\t%lg\t%lg\n", a, b, c);  // SU:760
    }
    fclose(fa);
  }
  snprintf(buffer, sizeof(buffer), "wvFns_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "w"))) return -2;
  for (int j = 0; j < N; j++) {
    for (int k = 0; k < kStates; k++) fprintf(fa, "%lg\t%lg\t", psi[((size_t)j * kStates + k) * 2], psi[((size_t)j * kStates + k) * 2 + 1]);  // SU:777
    fprintf(fa, "\n");
  }
  fclose(fa);
  return 0;
}

int mdqt_io_read_conditions(const char* dir, int c0, int ld, double* R, double* V, double* psi, unsigned* counter,
                            double* t_out, double* vholder) {
  char buffer[256];
  FILE* fa;
  int N = 0, j, m;
  if (t_out) *t_out = ((double)c0 - 9.) * 0.002 + 0.02;  // SU:789
  snprintf(buffer, sizeof(buffer), "ions_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "r"))) return -2;
  while (fscanf(fa, "%i\t%i", &j, &m) == 2) { N = j; if (counter) *counter = (unsigned)m; }  // SU:809-813
  fclose(fa);
  if (N > ld) return -1;
  snprintf(buffer, sizeof(buffer), "conditions_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "r"))) return -2;
  double a, b, z, d, e, f;
  int i = 0;
  while (fscanf(fa, "%lg\t%lg\t%lg\t%lg\t%lg\t%lg\n", &a, &b, &z, &d, &e, &f) == 6) {  // SU:822
    if (i >= ld) { fclose(fa); return -1; }
    R[i] = a; R[ld + i] = b; R[2 * ld + i] = z; V[i] = d; V[ld + i] = e; V[2 * ld + i] = f;
    i++;
  }
  fclose(fa);
  snprintf(buffer, sizeof(buffer), "wvFns_timestep%06d.dat", c0);
  if (!(fa = fopen(join(dir, buffer).c_str(), "r"))) return -2;
  i = 0;
  for (;;) {  // SU:863: 12 x "%lg%lg\t" then "\n"
    double w[kStates * 2];
    int got = 0;
    for (int k = 0; k < kStates; k++) {
      if (fscanf(fa, "%lg%lg", &w[2 * k], &w[2 * k + 1]) != 2) break;
      got += 2;
    }
    if (got != kStates * 2) break;
    if (i >= ld) { fclose(fa); return -1; }
    memcpy(psi + (size_t)i * kStates * 2, w, sizeof(w));
    i++;
  }
  fclose(fa);
  for (int c2V = 0; c2V < MDQT_NUM_VINTERVALS; c2V++) {
    snprintf(buffer, sizeof(buffer), "VZERO_timestep%06d_interval%d.dat", c0, c2V);
    if (!(fa = fopen(join(dir, buffer).c_str(), "r"))) return -2;  // the reference would crash here (SU:904-905)
    i = 0;
    while (fscanf(fa, "%lg\t%lg\t%lg", &a, &b, &z) == 3) {
      if (vholder && i < ld) {
        vholder[((size_t)0 * MDQT_NUM_VINTERVALS + c2V) * ld + i] = a;
        vholder[((size_t)1 * MDQT_NUM_VINTERVALS + c2V) * ld + i] = b;
        vholder[((size_t)2 * MDQT_NUM_VINTERVALS + c2V) * ld + i] = z;
      }
      i++;
    }
    fclose(fa);
  }
  return N;
}

int mdqt_io_append_energies(const char* dir, double t, double ekx, double eky, double ekz, double epot, double epot0,
                            double vx_avg) {
  FILE* fa = fopen(join(dir, "energies.dat").c_str(), "a");
  if (!fa) return -2;
  fprintf(fa, "%lg\t%lg\t%lg\t%lg\t%lg\t%lg\t%lg\n", t, ekx, eky, ekz, epot, ekx + eky + ekz + epot - epot0, vx_avg);  // SU:954
  fclose(fa);
  return 0;
}

int mdqt_io_write_vel_dist(const char* dir, unsigned counter, const double* pvel, double vx_avg) {
  static const char* names[3] = {"vel_distX_time%06d.dat", "vel_distY_time%06d.dat", "vel_distZ_time%06d.dat"};
  for (int c = 0; c < 3; c++) {
    char buffer[256];
    snprintf(buffer, sizeof(buffer), names[c], counter);
    FILE* fa = fopen(join(dir, buffer).c_str(), "w");
    if (!fa) return -2;
    for (int i = 0; i < kBins; i++) {
      double vel = (double)i * 0.0025;  // SU:342
      fprintf(fa, "%lg\t%lg\n", c == 0 ? vel + vx_avg : vel, pvel[c * kBins + i]);  // SU:1000-1002
    }
    fclose(fa);
  }
  return 0;
}

int mdqt_io_write_populations(const char* dir, unsigned counter, int N, const double* Vx, const double* pops) {
  char buffer[256];
  snprintf(buffer, sizeof(buffer), "statePopulationsVsVTime%06d.dat", counter);
  FILE* fa = fopen(join(dir, buffer).c_str(), "w");
  if (!fa) return -2;
  for (int i = 0; i < N; i++) fprintf(fa, "%lg\t%lg\t%lg\t%lg\n", Vx[i], pops[3 * i], pops[3 * i + 1], pops[3 * i + 2]);  // SU:1022
  fclose(fa);
  return 0;
}

int mdqt_schedule_next(int* c0, int* tsc, double* t, int ratio, int sampleFreq, double dtq, double tmax, int* do_output,
                       int* do_forces) {
  *do_output = 0; *do_forces = 0;
  if (!(*t <= tmax + 0.0009)) return 0;                                   // SU:1248
  if ((*c0 + 1) % sampleFreq == 0 && *tsc == 1) *do_output = 1;           // SU:1365
  if (*tsc == ratio) { *do_forces = 1; (*c0)++; *tsc = 0; }               // SU:1369-1375
  int k = 0;
  do {                                                                    // step(); qstep(); timeStepCounter++
    *t += dtq; (*tsc)++; k++;
  } while (*t <= tmax + 0.0009 && *tsc != ratio && !((*c0 + 1) % sampleFreq == 0 && *tsc == 1));
  return k;
}

}  // extern "C"

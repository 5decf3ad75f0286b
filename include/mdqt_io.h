/* include/mdqt_io.h -- host-side (no GPU) pieces of the reference's driver that surround the hot path, kept
 * byte-compatible so that the engine drops in for the simulation loop:
 *
 *   directory naming            SU:1147-1159   -> mdqt_io_dirname
 *   init()  (random frozen start, drand48 draw order)   SU:289-348   -> mdqt_io_init_su
 *   writeConditions(c0)         SU:725-784     -> mdqt_io_write_conditions
 *   readConditions(c0)          SU:785-916     -> mdqt_io_read_conditions
 *   output() file formats       SU:951-1024    -> mdqt_io_append_energies / mdqt_io_write_vel_dist / mdqt_io_write_populations
 *   main-loop schedule          SU:1248, 1365-1378 -> mdqt_schedule_next
 *
 * Same library (libmdqt_b200.so), plain C ABI, host arrays in the reference's layouts: R, V = double [3][ld];
 * psi = double [n][12][2]; Vholder = double [3][13][ld] (the VZERO files; all zeros in SU since the VAF code is
 * commented out, but they must exist or readConditions crashes, SU:898-913).
 */
#ifndef MDQT_IO_H
#define MDQT_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MDQT_NUM_VINTERVALS 13 /* numberOfIntervalV, SU:105 */

/* saveDirectory + "Ge%dDensity%dE+11Sig0%dTe%dSigFrac%dDetSP%dDetDP%dOmSP%dOmDP%dNumIons%d" + "/job%d/" (SU:1153-1158).
 * Creates the three directory levels when `create` != 0. Returns 0, or -1 if `out` (capacity cap) is too small. */
int mdqt_io_dirname(char* out, int cap, const char* saveDirectory, double Ge, double density, double sig0, double Te,
                    double fracOfSig, double detuning, double detuningDP, double Om, double OmDP, int N0, unsigned job,
                    int create);

/* init() (SU:289-348): Bernoulli thinning of 729*N0 trial points drawn with drand48() after srand48(seed), in the
 * reference's draw order (x, y, z, then rand1..rand4 for accepted ions). V = 0, tPart = 0. Returns N (>= 0), or -1 when
 * N would exceed ld. L_out / lDeb_out receive the box length and Debye length. */
int mdqt_io_init_su(long seed, int N0, double Ge, int ld, double* R, double* V, double* psi, double* tPart,
                    double* L_out, double* lDeb_out);

/* writeConditions(c0) (SU:725-784): ions_, conditions_, 13 x VZERO_, wvFns_ files in `dir` (trailing slash included).
 * vholder may be NULL (zeros are written, as the reference does with its commented-out VAF code). */
int mdqt_io_write_conditions(const char* dir, int c0, int N, unsigned counter, const double* R, const double* V,
                             const double* psi, int ld, const double* vholder);
/* readConditions(c0) (SU:785-916). Returns N (>= 0) or a negative error (missing file -> -2, too many ions -> -1);
 * *t_out = (c0-9)*0.002 + 0.02 (SU:789). vholder may be NULL. tPart is NOT restored by the reference (Q7): the caller zeroes it. */
int mdqt_io_read_conditions(const char* dir, int c0, int ld, double* R, double* V, double* psi, unsigned* counter,
                            double* t_out, double* vholder);

/* output() files (SU:951-1024). */
int mdqt_io_append_energies(const char* dir, double t, double ekx, double eky, double ekz, double epot, double epot0,
                            double vx_avg);
int mdqt_io_write_vel_dist(const char* dir, unsigned counter, const double* pvel /*[3][2001]*/, double vx_avg);
int mdqt_io_write_populations(const char* dir, unsigned counter, int N, const double* Vx, const double* pops /*[N][3]*/);

/* Main-loop schedule (SU:1248, 1365-1378) as a pure function. State: c0, timeStepCounter, t. Given the state at the
 * top of a loop iteration it reports which reference calls fall due BEFORE the next substep and how many substeps can
 * then be fused until the next due call or until `t <= tmax + 0.0009` fails:
 *   *do_output  : output() is due now            ((c0+1) % sampleFreq == 0 && timeStepCounter == 1)
 *   *do_forces  : forces(); c0++; counter reset   (timeStepCounter == ratio)  -- applied to *c0 / *tsc by this call
 *   return value: number of { step(); qstep(); } to run (0 = the loop has ended); *tsc and *t are advanced accordingly
 *                 with the reference's repeated addition t += dtq. */
int mdqt_schedule_next(int* c0, int* tsc, double* t, int ratio, int sampleFreq, double dtq, double tmax, int* do_output,
                       int* do_forces);

#ifdef __cplusplus
}
#endif
#endif

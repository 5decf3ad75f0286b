/* oracle/mdqt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Independent CPU restatement (plain C99, scalar, no FMA contraction) of the reference's per-timestep
 * MDQT hot path, written long-hand from the reference algorithm; every function cites the reference
 * file:line it follows (SU = laserCoolingPlusExpansionMDQTSpeedUp.cpp, MD =
 * MonteCarloFollowedByMDAndTempAnisotropy.cpp, MC408L = MonteCarloFollowedByQTTagging408Linear.cpp).
 *
 * Parity pinning: the reference holds no golden vectors or tests, so this
 * restatement is pinned against the reference ITSELF compiled unmodified (oracle/_ref/libref_*.so, built
 * by oracle/Makefile from /root/reference) in tests/test_oracle_vs_reference.py, and against the vectors
 * that run produced, committed under tests/golden/ by oracle/gen_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 * The product (mdqtplasmasims_b200/) never does.
 */
#ifndef MDQT_ORACLE_H
#define MDQT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- counter-based RNG shared by oracle and product spec: Philox4x32-10 (Salmon et al., SC'11) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the five uniforms (rand, rand2, randDOrS, randDir, rand3) of ion `ion` of trajectory `traj` at global
 * quantum-substep index `substep`; u in (0,1) with 53 random bits: u = (k + 0.5) * 2^-53 */
void orc_uniforms5(uint64_t seed, uint32_t traj, uint32_t ion, uint64_t substep, double u[5]);
/* stream 1 of the same generator, used by the MD-family Andersen thermostat: one uniform + 3 normals */
void orc_collision_draws(uint64_t seed, uint32_t traj, uint32_t ion, uint64_t step, double* u, double nrm[3]);

/* ---- Yukawa forces / potential energy: arrays are [3][n] contiguous ---- */
void orc_forces_su(int n, const double* R, double L, double lDeb, double* F);                 /* SU:192-236 */
void orc_forces_md(int n, const double* R, double L, double kappa, double rCut, double* A);   /* MD:161-169, 387-448 */
double orc_epot_su(int n, const double* R, double L, double lDeb);                            /* SU:244-281 */

/* ---- SU integrator: one step() = step_R(dt/2); step_V(dt); step_R(dt/2) with dt = dtq (SU:356-430) ---- */
void orc_step_su(int n, double* R, double* V, const double* F, double L, double dtq, double t);

/* ---- MD-family velocity Verlet (MD:452-502). coll_u / coll_n ([n] / [n][3]) replace the mt19937 draws;
 *      coll_u may be NULL (no collisions). laser: 0 none, 1 three-axis, 2 x-axis only. ---- */
void orc_vv_positions(int n, double* R, const double* V, const double* A, double L, double dt);
void orc_vv_velocities(int n, double* V, const double* oldA, const double* A, double dt, double collisionFreq,
                       const double* coll_u, const double* coll_n, int laser, double beta, double dens);

/* ---- quantum-trajectory step ---- */
typedef struct {
  double detuning, detuningDP, Om, OmDP; /* SU:70-73 */
  double dR, kRat;                       /* decayRatioD5Halves, kRat SU:146-147 */
  double vKick, vKickDP;                 /* SU:148-149 */
  double g2E, pv2qv, dtq;                /* gamToEinsteinFreq, plasVelToQuantVel, quantumTimestep SU:79-85 */
  double fracOfSig, Te, sig0, density;   /* expansion detuning SU:447 */
  int renorm;                            /* reNormalizewvFns SU:74 */
  int quad;                              /* 7-level only: MC408Q coupling mask (MC408Q:596) */
} orc_qt_params;

/* uniforms: umode 0 = per-ion table u[n][5]; umode 1 = one sequential stream u[] consumed in ion order
 * (cursor advanced). jumped[i] (may be NULL) = number of uniforms ion i consumed (1 no jump, 4/5 jump). */
void orc_qstep12(int n, double* psi, double* Vx, double* tPart, double* t, const orc_qt_params* p,
                 const double* u, int umode, long* cursor, int* used);                         /* SU:438-717, 1163-1215 */
void orc_qstep7(int n, double* psi, const double* Vx, const orc_qt_params* p,
                const double* u, int umode, long* cursor, int* used);                          /* MC408L:555-756, 1171-1190 */

void orc_qstep5(int n, double* psi, const double* Vx, const orc_qt_params* p,
                const double* u, int umode, long* cursor, int* used);                          /* MC422L:552-727, 1144-1155 */

void orc_qstep3(int n, double* psi, double* Vx, double* tPart, const orc_qt_params* p, int applyForce,
                const double* u, int umode, long* cursor, int* used);                          /* TS:140-293, 379-382 */
int orc_tag(int n, int S, const double* psi, const double* u, int umode, long* cursor, int* tagged); /* MC408L:1022-1067, MC422L:992-1036 */
void orc_lf_drift(int n, double* R, const double* V, const double* F, double L, double DT, int first); /* FZ408L:317-350 */
void orc_lf_kick(int n, double* V, const double* F, double DT);                                 /* FZ408L:358-369 */
double orc_vaf(int n, const double* Vhold, const double* Vx);                                   /* FZ408L:938-961 */

void orc_pair_correlation(int n, const double* R, double L, double step, double rmax, double* counts, double* g); /* MD:584-652 */
void orc_autocorr(int which, int n, int nnorm, int T, const double* vstore, double Gamma, double* out);         /* MD:654-823 */

#ifdef __cplusplus
}
#endif
#endif

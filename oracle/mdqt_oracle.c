/* oracle/mdqt_oracle.c -- TEST INFRASTRUCTURE ONLY (see mdqt_oracle.h for scope, pinning and usage rules).
 * Plain scalar C99; build with -ffp-contract=off so that every operation rounds once, like the x86-64
 * reference build. Arrays: R,V,F,A = [3][n] contiguous; psi = [n][S][2] (re,im); S = 12 or 7. */
#define _GNU_SOURCE /* M_PI under -std=c99 */
#include "mdqt_oracle.h"
#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u52(uint32_t hi, uint32_t lo) { /* (k+1/2)*2^-52, k = 52 random bits: exact, in (0,1) */
  uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
  return ((double)k + 0.5) * 2.220446049250313080847263336181640625e-16;
}

static void philox_call(uint64_t seed, uint32_t traj, uint32_t ion, uint64_t step, uint32_t call, uint32_t o[4]) {
  uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), ion, (traj << 3) | call};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, o);
}

void orc_uniforms5(uint64_t seed, uint32_t traj, uint32_t ion, uint64_t substep, double u[5]) {
  uint32_t o[4];
  philox_call(seed, traj, ion, substep, 0, o); u[0] = u52(o[0], o[1]); u[1] = u52(o[2], o[3]);
  philox_call(seed, traj, ion, substep, 1, o); u[2] = u52(o[0], o[1]); u[3] = u52(o[2], o[3]);
  philox_call(seed, traj, ion, substep, 2, o); u[4] = u52(o[0], o[1]);
}

void orc_collision_draws(uint64_t seed, uint32_t traj, uint32_t ion, uint64_t step, double* u, double nrm[3]) {
  uint32_t o[4];
  double ua, ub, uc, ud, r;
  philox_call(seed, traj, ion, step, 3, o); *u = u52(o[0], o[1]); ua = u52(o[2], o[3]);
  philox_call(seed, traj, ion, step, 4, o); ub = u52(o[0], o[1]); uc = u52(o[2], o[3]);
  philox_call(seed, traj, ion, step, 5, o); ud = u52(o[0], o[1]);
  r = sqrt(-2.0 * log(ua));
  nrm[0] = r * cos(6.283185307179586476925286766559 * ub);
  nrm[1] = r * sin(6.283185307179586476925286766559 * ub);
  nrm[2] = sqrt(-2.0 * log(uc)) * cos(6.283185307179586476925286766559 * ud);
}

/* ------------------------------------------------------------------------------------------------ */
/* forces and potential energy                                                                        */
/* ------------------------------------------------------------------------------------------------ */
void orc_forces_su(int n, const double* R, double L, double lDeb, double* F) { /* SU:192-236 */
  const double* X = R; const double* Y = R + n; const double* Z = R + 2 * n;
  double* FX = F; double* FY = F + n; double* FZ = F + 2 * n;
  double Rcut = L / 2.;
  for (int i = 0; i < 3 * n; i++) F[i] = 0.;
  for (int i = 0; i < n - 1; i++) {
    double rx = X[i], ry = Y[i], rz = Z[i];
    for (int j = i + 1; j < n; j++) {
      double dx = rx - X[j], dy = ry - Y[j], dz = rz - Z[j];
      dx -= L * round(dx / L); dy -= L * round(dy / L); dz -= L * round(dz / L);
      double dr = sqrt(dx * dx + dy * dy + dz * dz);
      if (dr > 0 && dr < Rcut) {
        double ft = (1. / dr + 1. / lDeb) * exp(-dr / lDeb) / (dr * dr);
        double fx = dx * ft, fy = dy * ft, fz = dz * ft;
        FX[i] += fx; FX[j] -= fx; FY[i] += fy; FY[j] -= fy; FZ[i] += fz; FZ[j] -= fz;
      }
    }
  }
}

void orc_forces_md(int n, const double* R, double L, double kappa, double rCut, double* A) { /* MD:387-448 */
  const double* X = R; const double* Y = R + n; const double* Z = R + 2 * n;
  for (int i = 0; i < 3 * n; i++) A[i] = 0.;
  for (int i = 0; i < n; i++) {
    for (int j = i + 1; j < n; j++) {
      double dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
      dx -= L * round(dx / L); dy -= L * round(dy / L); dz -= L * round(dz / L);
      double d = sqrt(dx * dx + dy * dy + dz * dz);
      double pre = 0.;
      if (d < rCut) pre = exp(-1 * kappa * d) * (pow(d, -3) + kappa / (d * d)); /* calcAIJ MD:161-169 */
      double a;
      a = pre * dx; A[i] += a; A[j] -= a;
      a = pre * dy; A[n + i] += a; A[n + j] -= a;
      a = pre * dz; A[2 * n + i] += a; A[2 * n + j] -= a;
    }
  }
}

double orc_epot_su(int n, const double* R, double L, double lDeb) { /* SU:244-281 */
  const double* X = R; const double* Y = R + n; const double* Z = R + 2 * n;
  double Rcut = L / 2., E = 0.;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      double dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
      dx -= L * round(dx / L); dy -= L * round(dy / L); dz -= L * round(dz / L);
      double dr = sqrt(dx * dx + dy * dy + dz * dz);
      if (dr > 0 && dr < Rcut) E += exp(-dr / lDeb) / dr;
    }
  return E / (double)n;
}

/* ------------------------------------------------------------------------------------------------ */
/* integrators                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
static void su_step_R(int n, double* R, const double* V, const double* F, double L, double DT, double t) { /* SU:356-390 */
  for (int k = 0; k < 3 * n; k++) {
    if (t > 0) R[k] += DT * V[k];
    else R[k] += DT * V[k] + DT * DT * F[k];
    if (R[k] < 0) R[k] += L;
    if (R[k] > L) R[k] -= L;
  }
}
void orc_step_su(int n, double* R, double* V, const double* F, double L, double dtq, double t) { /* SU:418-430 */
  su_step_R(n, R, V, F, L, 0.5 * dtq, t);
  for (int k = 0; k < 3 * n; k++) V[k] += dtq * F[k]; /* step_V SU:398-409 */
  su_step_R(n, R, V, F, L, 0.5 * dtq, t);
}

void orc_vv_positions(int n, double* R, const double* V, const double* A, double L, double dt) { /* MD:452-467 */
  for (int k = 0; k < 3 * n; k++) {
    R[k] = R[k] + dt * V[k] + dt * dt / 2 * A[k];
    if (R[k] < 0) R[k] += L;
    if (R[k] > L) R[k] -= L;
  }
}
void orc_vv_velocities(int n, double* V, const double* oldA, const double* A, double dt, double collisionFreq,
                       const double* coll_u, const double* coll_n, int laser, double beta, double dens) { /* MD:469-502 */
  for (int i = 0; i < n; i++) {
    if (coll_u && coll_u[i] < dt * collisionFreq) {
      for (int c = 0; c < 3; c++) V[c * n + i] = coll_n[3 * i + c];
    } else {
      for (int c = 0; c < 3; c++) V[c * n + i] = V[c * n + i] + dt / 2 * (oldA[c * n + i] + A[c * n + i]);
    }
    if (laser == 2) {
      V[i] += V[i] * dt * 1.234 * pow(10, -6) * beta / sqrt(dens);
    } else if (laser == 1) {
      V[i] += V[i] * dt * 1.234 * pow(10, -6) * beta / sqrt(dens) / 2;
      V[n + i] += V[n + i] * dt * 1.234 * pow(10, -6) * beta / sqrt(dens) / 4 * (-1);
      V[2 * n + i] += V[2 * n + i] * dt * 1.234 * pow(10, -6) * beta / sqrt(dens) / 4 * (-1);
    }
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* quantum trajectories                                                                               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { double re, im; } cplx;
static cplx cmul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static double cnorm(cplx a) { return a.re * a.re + a.im * a.im; }

/* sparse Hamiltonian: diagonal (complex) + list of off-diagonal entries (row, col, value); the Hermitian
 * conjugates of the couplings are listed explicitly. */
typedef struct { int r, c; cplx v; } hent;
typedef struct { int S; cplx diag[12]; int nent; hent ent[32]; double gam[12]; } sparse_h;

static void h_apply(const sparse_h* H, const cplx* y, cplx* out) {
  for (int k = 0; k < H->S; k++) out[k] = cmul(H->diag[k], y[k]);
  for (int e = 0; e < H->nent; e++) {
    cplx t = cmul(H->ent[e].v, y[H->ent[e].c]);
    out[H->ent[e].r].re += t.re; out[H->ent[e].r].im += t.im;
  }
}
static double dp_of(const sparse_h* H, const cplx* y, double h) { /* h * y^dagger D y, D diagonal (SU:484, 530) */
  double s = 0.;
  for (int k = 0; k < H->S; k++) if (H->gam[k] != 0.) s += h * H->gam[k] * cnorm(y[k]);
  return s;
}
/* slope f(y) = ( M y / sqrt(1-dp(y)) - y ) / h with M = 1 - i h H  (SU:526-535) */
static void slope(const sparse_h* H, const cplx* y, double h, cplx* k) {
  cplx Hy[12];
  double pre = 1 / sqrt(1 - dp_of(H, y, h));
  h_apply(H, y, Hy);
  for (int s = 0; s < H->S; s++) {
    /* (1 - i h H) y = y + h*(Hy.im) - i h*(Hy.re) */
    cplx my = {y[s].re + h * Hy[s].im, y[s].im - h * Hy[s].re};
    k[s].re = 1. / h * (pre * my.re - y[s].re);
    k[s].im = 1. / h * (pre * my.im - y[s].im);
  }
}
/* the reference's 4-stage scheme: RK4 nodes, (1,3,3,1)/8 weights (SU:523-567; MC408L:614-672) */
static void rk_step(const sparse_h* H, cplx* y, double h) {
  cplx k1[12], k2[12], k3[12], k4[12], w[12];
  int S = H->S;
  double hh = h / 2;
  slope(H, y, h, k1);
  for (int s = 0; s < S; s++) { w[s].re = y[s].re + hh * k1[s].re; w[s].im = y[s].im + hh * k1[s].im; }
  slope(H, w, h, k2);
  for (int s = 0; s < S; s++) { w[s].re = y[s].re + hh * k2[s].re; w[s].im = y[s].im + hh * k2[s].im; }
  slope(H, w, h, k3);
  for (int s = 0; s < S; s++) { w[s].re = y[s].re + h * k3[s].re; w[s].im = y[s].im + h * k3[s].im; }
  slope(H, w, h, k4);
  for (int s = 0; s < S; s++) {
    y[s].re = y[s].re + (k1[s].re + 3 * k2[s].re + 3 * k3[s].re + k4[s].re) / 8 * h;
    y[s].im = y[s].im + (k1[s].im + 3 * k2[s].im + 3 * k3[s].im + k4[s].im) / 8 * h;
  }
}
static void add_pair(sparse_h* H, int r, int c, cplx v) { /* entry and its Hermitian conjugate (SU:517) */
  H->ent[H->nent].r = r; H->ent[H->nent].c = c; H->ent[H->nent].v = v; H->nent++;
  H->ent[H->nent].r = c; H->ent[H->nent].c = r; H->ent[H->nent].v.re = v.re; H->ent[H->nent].v.im = -v.im; H->nent++;
}

static double next_u(const double* u, int umode, long* cursor, int i, int* k) {
  double x = (umode == 0) ? u[5 * i + *k] : u[(*cursor)++];
  (*k)++;
  return x;
}

/* 12-level channel amplitudes gs[k] (SU:1181-1198) */
static void su_gs(double dR, double gs[18]) {
  gs[0] = sqrt(1.); gs[1] = sqrt(2. / 3); gs[2] = sqrt(1. / 3); gs[3] = sqrt(2. / 3); gs[4] = sqrt(1. / 3); gs[5] = sqrt(1.);
  gs[6] = sqrt(dR * 2. / 3); gs[7] = sqrt(dR * 4. / 15); gs[8] = sqrt(dR * 1. / 15); gs[9] = sqrt(dR * 2. / 5);
  gs[10] = sqrt(dR * 2. / 5); gs[11] = sqrt(dR * 1. / 5); gs[12] = sqrt(dR * 1. / 5); gs[13] = sqrt(dR * 2. / 5);
  gs[14] = sqrt(dR * 2. / 5); gs[15] = sqrt(dR * 1. / 15); gs[16] = sqrt(dR * 4. / 15); gs[17] = sqrt(dR * 2. / 3);
}
/* upper (decaying) state of channel k, 0-indexed (cs[k] = |lower><upper|, SU:1163-1180) */
static const int su_upper[18] = {2, 3, 3, 4, 4, 5, 5, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2};
static const int su_lower[18] = {1, 1, 0, 0, 1, 0, 6, 7, 8, 7, 8, 9, 8, 9, 10, 9, 10, 11};

void orc_qstep12(int n, double* psi, double* Vx, double* tPart, double* t, const orc_qt_params* p,
                 const double* u, int umode, long* cursor, int* used) {
  double gs[18];
  sparse_h H;
  const double dtq = p->dtq, g2E = p->g2E, dR = p->dR, kRat = p->kRat;
  const double h = dtq * g2E;
  /* expansion detuning from the global time BEFORE it is advanced (SU:447, 716) */
  const double tt = *t;
  const double expDet = 0.0126 * p->fracOfSig * p->Te * tt /
                        (sqrt(p->density) * p->sig0 * sqrt(1 + 0.00014314 * tt * tt * p->Te / (p->density * p->sig0 * p->sig0)));
  su_gs(dR, gs);
  memset(&H, 0, sizeof(H));
  H.S = 12;
  for (int k = 0; k < 18; k++) H.gam[su_upper[k]] += gs[k] * gs[k]; /* decayMatrix diagonal (SU:1203) */

  for (int i = 0; i < n; i++) {
    cplx y[12];
    int nu = 0;
    double kick;
    for (int s = 0; s < 12; s++) { y[s].re = psi[(i * 12 + s) * 2]; y[s].im = psi[(i * 12 + s) * 2 + 1]; }
    double vq = Vx[i] * p->pv2qv;  /* SU:481-482 */
    tPart[i] += dtq;               /* SU:483 */
    double dp = dp_of(&H, y, h);   /* SU:484-485 */
    double rnd = next_u(u, umode, cursor, i, &nu); /* SU:486 */
    if (rnd > dp) {
      /* optical force from the PRE-step coherences, Im(rho_ab) = Im(psi_a conj(psi_b)) (SU:490-503) */
#define IMRHO(a, b) (y[(a) - 1].im * y[(b) - 1].re - y[(a) - 1].re * y[(b) - 1].im)
      kick = 1 * p->vKick * p->Om * (IMRHO(2, 3) * gs[0] + IMRHO(1, 4) * gs[2] - IMRHO(2, 5) * gs[4] - IMRHO(1, 6) * gs[5]) * dtq * g2E +
             p->vKickDP * (p->OmDP / dR) *
                 (IMRHO(9, 6) * gs[8] + IMRHO(10, 5) * gs[11] + IMRHO(11, 4) * gs[14] + IMRHO(12, 3) * gs[17] - IMRHO(7, 6) * gs[6] -
                  IMRHO(8, 5) * gs[9] - IMRHO(9, 4) * gs[12] - IMRHO(10, 3) * gs[15]) * dtq * g2E;
#undef IMRHO
      /* Hamiltonian (SU:506-521) */
      double uu = vq + expDet;
      double detR = -p->detuning - vq - expDet, detL = -p->detuning + vq + expDet;
      double eD1 = -p->detuning + p->detuningDP + (1 - kRat) * uu;          /* states 7,8  (idx 6,7)  */
      double eD3 = -p->detuning + p->detuningDP + (kRat - 1) * uu;          /* states 11,12 (idx 10,11) */
      double eD2 = -p->detuning + p->detuningDP - vq - expDet - kRat * uu;  /* states 9,10 (idx 8,9)  */
      for (int s = 0; s < 12; s++) { H.diag[s].re = 0; H.diag[s].im = 0; }
      H.diag[2].re = detR; H.diag[3].re = detR; H.diag[4].re = detL; H.diag[5].re = detL;
      H.diag[6].re = eD1; H.diag[7].re = eD1; H.diag[8].re = eD2; H.diag[9].re = eD2; H.diag[10].re = eD3; H.diag[11].re = eD3;
      for (int s = 2; s < 6; s++) H.diag[s].im = -1. / 2 * H.gam[s]; /* hamDecayTerm (SU:1202) */
      H.nent = 0;
      /* static couplings -cs[k]^dagger gs[k] Om/2, k=0,2,4,5 and -cs[k]^dagger gs[k] OmDP/2/sqrt(dR), k=6,9,12,14,15,17 (SU:1206-1215) */
      static const int ksp[4] = {0, 2, 4, 5}, kdp[6] = {6, 9, 12, 14, 15, 17};
      for (int a = 0; a < 4; a++) { cplx v = {-1. * gs[ksp[a]] * p->Om / 2, 0}; add_pair(&H, su_upper[ksp[a]], su_lower[ksp[a]], v); }
      for (int a = 0; a < 6; a++) { cplx v = {-1. * gs[kdp[a]] * p->OmDP / 2 / sqrt(dR), 0}; add_pair(&H, su_upper[kdp[a]], su_lower[kdp[a]], v); }
      /* phase-rotating couplings on |9><6| and |10><5| (SU:508) */
      double phi = 2. * uu * (1 + kRat) * tPart[i] * g2E;
      cplx eph = {cos(phi), sin(phi)};
      double a8 = p->OmDP / 2 * gs[8] / sqrt(dR), a11 = p->OmDP / 2 * gs[11] / sqrt(dR);
      cplx v8 = {-(a8 * eph.re), -(a8 * eph.im)}, v11 = {-(a11 * eph.re), -(a11 * eph.im)};
      add_pair(&H, 8, 5, v8);
      add_pair(&H, 9, 4, v11);
      rk_step(&H, y, h);
    } else { /* quantum jump (SU:573-703) */
      tPart[i] = 0;
      double rand2 = next_u(u, umode, cursor, i, &nu);
      double n3 = cnorm(y[2]), n4 = cnorm(y[3]), n5 = cnorm(y[4]), n6 = cnorm(y[5]);
      double tot = n3 + n4 + n5 + n6;
      double p3 = n3 / tot, p4 = n4 / tot, p5 = n5 / tot;
      double randDOrS = next_u(u, umode, cursor, i, &nu);
      double randDir = next_u(u, umode, cursor, i, &nu);
      int sDecay = 1, dest;
      if (randDOrS < (dR / (dR + 1))) { sDecay = 0; kick = (randDir < 0.5) ? p->vKickDP : -p->vKickDP; }
      else kick = (randDir < 0.5) ? p->vKick : -p->vKick;
      if (rand2 < p3) {
        if (sDecay) dest = 1;
        else { double r3 = next_u(u, umode, cursor, i, &nu);
          if (r3 < gs[17] * gs[17] / dR) dest = 11; else if (r3 < gs[17] * gs[17] / dR + gs[16] * gs[16] / dR) dest = 10; else dest = 9; }
      } else if (rand2 < p3 + p4) {
        double r3 = next_u(u, umode, cursor, i, &nu);
        if (sDecay) dest = (r3 < gs[2] * gs[2]) ? 0 : 1;
        else { if (r3 < gs[14] * gs[14] / dR) dest = 10; else if (r3 < gs[14] * gs[14] / dR + gs[13] * gs[13] / dR) dest = 9; else dest = 8; }
      } else if (rand2 < p3 + p4 + p5) {
        double r3 = next_u(u, umode, cursor, i, &nu);
        if (sDecay) dest = (r3 < gs[4] * gs[4]) ? 1 : 0;
        else { if (r3 < gs[11] * gs[11] / dR) dest = 9; else if (r3 < gs[11] * gs[11] / dR + gs[10] * gs[10] / dR) dest = 8; else dest = 7; }
      } else {
        if (sDecay) dest = 0;
        else { double r3 = next_u(u, umode, cursor, i, &nu);
          if (r3 < gs[8] * gs[8] / dR) dest = 8; else if (r3 < gs[8] * gs[8] / dR + gs[7] * gs[7] / dR) dest = 7; else dest = 6; }
      }
      for (int s = 0; s < 12; s++) { y[s].re = 0; y[s].im = 0; }
      y[dest].re = 1;
    }
    Vx[i] = Vx[i] + kick; /* SU:705 */
    if (p->renorm) {      /* SU:706-712 */
      double popS = cnorm(y[0]) + cnorm(y[1]);
      double popP = cnorm(y[2]) + cnorm(y[3]) + cnorm(y[4]) + cnorm(y[5]);
      double popD = cnorm(y[6]) + cnorm(y[7]) + cnorm(y[8]) + cnorm(y[9]) + cnorm(y[10]) + cnorm(y[11]);
      double nn = sqrt(popS + popP + popD);
      for (int s = 0; s < 12; s++) { y[s].re /= nn; y[s].im /= nn; }
    }
    for (int s = 0; s < 12; s++) { psi[(i * 12 + s) * 2] = y[s].re; psi[(i * 12 + s) * 2 + 1] = y[s].im; }
    if (used) used[i] = nu;
  }
  *t += dtq; /* SU:716 */
}

/* 7-level 408 nm pump (MC408L:555-756): gs are RATES here (MC408L:1181-1190); no kick, no tPart, no t. */
void orc_qstep7(int n, double* psi, const double* Vx, const orc_qt_params* p,
                const double* u, int umode, long* cursor, int* used) {
  static const int up7[10] = {2, 3, 4, 3, 4, 5, 2, 3, 4, 5}; /* cs[k]=|lower><upper| (MC408L:1171-1180) */
  double gs[10] = {1, 2. / 3, 1. / 3, 1. / 3, 2. / 3, 1, p->dR, p->dR, p->dR, p->dR};
  const double h = p->dtq * p->g2E;
  sparse_h H;
  memset(&H, 0, sizeof(H));
  H.S = 7;
  for (int k = 0; k < 10; k++) H.gam[up7[k]] += gs[k]; /* sum_j gs_j c_j^dagger c_j (MC408L:584-588, 603-606) */
  for (int i = 0; i < n; i++) {
    cplx y[12];
    int nu = 0;
    for (int s = 0; s < 7; s++) { y[s].re = psi[(i * 7 + s) * 2]; y[s].im = psi[(i * 7 + s) * 2 + 1]; }
    double vq = Vx[i] * p->pv2qv; /* MC408L:581-582 */
    double dp = dp_of(&H, y, h);
    double rnd = next_u(u, umode, cursor, i, &nu); /* MC408L:589 */
    if (rnd > dp) {
      double detR = -p->detuning - vq, detL = -p->detuning + vq; /* MC408L:595-596 */
      for (int s = 0; s < 7; s++) { H.diag[s].re = 0; H.diag[s].im = -1. / 2 * H.gam[s]; }
      H.diag[2].re = detR; H.diag[3].re = detR; H.diag[4].re = detL; H.diag[5].re = detL;
      H.nent = 0;
      /* -Om/2 |2><4| sqrt(gs3) - Om/2 |2><6| sqrt(gs5) - Om/2 |1><3| sqrt(gs0) - Om/2 |1><5| sqrt(gs2)  (1-indexed kets, MC408L:597);
       * the Quad variant keeps only |2><6| and |1><5| (MC408Q:596) */
      cplx v;
      v.im = 0;
      if (!p->quad) { v.re = -p->Om / 2 * sqrt(gs[3]); add_pair(&H, 1, 3, v); }
      v.re = -p->Om / 2 * sqrt(gs[5]); add_pair(&H, 1, 5, v);
      if (!p->quad) { v.re = -p->Om / 2 * sqrt(gs[0]); add_pair(&H, 0, 2, v); }
      v.re = -p->Om / 2 * sqrt(gs[2]); add_pair(&H, 0, 4, v);
      rk_step(&H, y, h);
    } else { /* MC408L:674-752 */
      double rand2 = next_u(u, umode, cursor, i, &nu);
      double n3 = cnorm(y[2]), n4 = cnorm(y[3]), n5 = cnorm(y[4]), n6 = cnorm(y[5]);
      double tot = n3 + n4 + n5 + n6;
      double p3 = n3 / tot, p4 = n4 / tot, p5 = n5 / tot;
      double randDOrS = next_u(u, umode, cursor, i, &nu);
      (void)next_u(u, umode, cursor, i, &nu); /* randDir: drawn, unused (MC408L:691) */
      int sDecay = !(randDOrS < (p->dR / (p->dR + 1))), dest;
      if (rand2 < p3) dest = sDecay ? 0 : 6;
      else if (rand2 < p3 + p4) { if (sDecay) { double r3 = next_u(u, umode, cursor, i, &nu); dest = (r3 < gs[1]) ? 0 : 1; } else dest = 6; }
      else if (rand2 < p3 + p4 + p5) { if (sDecay) { double r3 = next_u(u, umode, cursor, i, &nu); dest = (r3 < gs[2]) ? 0 : 1; } else dest = 6; }
      else dest = sDecay ? 1 : 6;
      for (int s = 0; s < 7; s++) { y[s].re = 0; y[s].im = 0; }
      y[dest].re = 1;
    }
    for (int s = 0; s < 7; s++) { psi[(i * 7 + s) * 2] = y[s].re; psi[(i * 7 + s) * 2 + 1] = y[s].im; }
    if (used) used[i] = nu;
  }
}

/* 5-level 422 nm pump (MonteCarloFollowedByQTTagging422Linear.cpp:552-727; tables MC422L:1144-1155): states
 * 0 S-1/2, 1 S+1/2, 2 P+1/2, 3 P-1/2, 4 D. gs are rates; draws: rand, rand2, randDOrS, [rand3 on S decays] -- there is
 * no randDir draw in this file. In the per-ion table mode (umode 0) the slots are u[0]=rand, u[1]=rand2, u[2]=randDOrS,
 * u[4]=rand3 (slot 3 unused), matching the engine's Philox layout. */
static double u_slot(const double* u, int umode, long* cursor, int i, int slot, int* count) {
  (*count)++;
  return (umode == 0) ? u[5 * i + slot] : u[(*cursor)++];
}
void orc_qstep5(int n, double* psi, const double* Vx, const orc_qt_params* p,
                const double* u, int umode, long* cursor, int* used) {
  static const int up5[6] = {2, 3, 3, 2, 2, 3}; /* cs[k]=|lower><upper| (MC422L:1144-1149) */
  double gs[6] = {2. / 3, 1. / 3, 2. / 3, 1. / 3, p->dR, p->dR};
  const double h = p->dtq * p->g2E;
  sparse_h H;
  memset(&H, 0, sizeof(H));
  H.S = 5;
  for (int k = 0; k < 6; k++) H.gam[up5[k]] += gs[k];
  for (int i = 0; i < n; i++) {
    cplx y[12];
    int nu = 0;
    for (int s = 0; s < 5; s++) { y[s].re = psi[(i * 5 + s) * 2]; y[s].im = psi[(i * 5 + s) * 2 + 1]; }
    double vq = Vx[i] * p->pv2qv;
    double dp = dp_of(&H, y, h);
    double rnd = u_slot(u, umode, cursor, i, 0, &nu); /* MC422L:586 */
    if (rnd > dp) {
      double detR = -p->detuning - vq, detL = -p->detuning + vq; /* MC422L:592-593 */
      for (int s = 0; s < 5; s++) { H.diag[s].re = 0; H.diag[s].im = -1. / 2 * H.gam[s]; }
      H.diag[2].re = detR; H.diag[3].re = detL; /* MC422L:595 */
      H.nent = 0;
      cplx v;
      v.im = 0;
      v.re = -p->Om / 2 * sqrt(gs[0]); add_pair(&H, 1, 2, v); /* -Om/2 |2><3| sqrt(gs0) (MC422L:594) */
      v.re = -p->Om / 2 * sqrt(gs[2]); add_pair(&H, 0, 3, v); /* -Om/2 |1><4| sqrt(gs2) */
      rk_step(&H, y, h);
    } else { /* MC422L:660-720 */
      double rand2 = u_slot(u, umode, cursor, i, 1, &nu);
      double n3 = cnorm(y[2]), n4 = cnorm(y[3]);
      double tot = n3 + n4;
      double p3 = n3 / tot;
      double randDOrS = u_slot(u, umode, cursor, i, 2, &nu);
      int sDecay = !(randDOrS < (p->dR / (p->dR + 1))), dest;
      if (rand2 < p3) {
        if (sDecay) { double r3 = u_slot(u, umode, cursor, i, 4, &nu); dest = (r3 < gs[0]) ? 1 : 0; } else dest = 4;
      } else {
        if (sDecay) { double r3 = u_slot(u, umode, cursor, i, 4, &nu); dest = (r3 < gs[2]) ? 0 : 1; } else dest = 4;
      }
      for (int s = 0; s < 5; s++) { y[s].re = 0; y[s].im = 0; }
      y[dest].re = 1;
    }
    for (int s = 0; s < 5; s++) { psi[(i * 5 + s) * 2] = y[s].re; psi[(i * 5 + s) * 2 + 1] = y[s].im; }
    if (used) used[i] = nu;
  }
}

/* 3-level test system (laserCoolNoPlasmaThreeState.cpp:140-293; tables TS:379-382): states 0 ground (J=0), 1 m=+1,
 * 2 m=-1; gs = {1,1}; cs[0] = |0><1|, cs[1] = |0><2|. dt is used directly (velocities and times already in quantum
 * units: pass dtq = dt, g2E = 1, pv2qv = 1). Draws: rand, and randDir on a jump (slots 0 and 3 of the 5-table). */
void orc_qstep3(int n, double* psi, double* Vx, double* tPart, const orc_qt_params* p, int applyForce,
                const double* u, int umode, long* cursor, int* used) {
  double gs[2] = {1, 1};
  const double dt = p->dtq * p->g2E;
  sparse_h H;
  memset(&H, 0, sizeof(H));
  H.S = 3;
  H.gam[1] += gs[0]; H.gam[2] += gs[1];
  for (int i = 0; i < n; i++) {
    cplx y[12];
    int nu = 0;
    double kick;
    for (int s = 0; s < 3; s++) { y[s].re = psi[(i * 3 + s) * 2]; y[s].im = psi[(i * 3 + s) * 2 + 1]; }
    double vq = Vx[i] * p->pv2qv; /* TS:154 */
    tPart[i] += dt;               /* TS:155 */
    double dp = dp_of(&H, y, dt);
    double rnd = u_slot(u, umode, cursor, i, 0, &nu); /* TS:165 */
    if (rnd > dp) {
      /* p13 = <1|rho|3> = y0 conj(y2), p12 = y0 conj(y1) (TS:170-171) */
      double im13 = y[0].im * y[2].re - y[0].re * y[2].im, im12 = y[0].im * y[1].re - y[0].re * y[1].im;
      kick = 1 * p->vKick * p->Om * (im13 * sqrt(gs[0]) - im12 * sqrt(gs[1])) * dt; /* TS:174 */
      double detR = -p->detuning - vq, detL = -p->detuning + vq; /* TS:177-178 */
      for (int s = 0; s < 3; s++) { H.diag[s].re = 0; H.diag[s].im = -1. / 2 * H.gam[s]; }
      H.diag[2].re = detR; H.diag[1].re = detL; /* TS:181 */
      H.nent = 0;
      cplx v;
      v.im = 0;
      v.re = -p->Om / 2 * sqrt(gs[0]); add_pair(&H, 0, 2, v); /* -Om/2 |1><3| sqrt(gs0) (TS:179) */
      v.re = -p->Om / 2 * sqrt(gs[1]); add_pair(&H, 0, 1, v); /* -Om/2 |1><2| sqrt(gs1) */
      rk_step(&H, y, dt);
    } else { /* TS:270-282 */
      tPart[i] = 0;
      for (int s = 0; s < 3; s++) { y[s].re = 0; y[s].im = 0; }
      y[0].re = 1;
      double randDir = u_slot(u, umode, cursor, i, 3, &nu);
      kick = (randDir < 0.5) ? p->vKick : -p->vKick;
    }
    if (applyForce) Vx[i] = Vx[i] + kick; /* TS:283-285 */
    for (int s = 0; s < 3; s++) { psi[(i * 3 + s) * 2] = y[s].re; psi[(i * 3 + s) * 2 + 1] = y[s].im; }
    if (used) used[i] = nu;
  }
}

/* projective spin measurement: tagParticles (MC408L:1022-1067 for S = 7, MC422L:992-1036 for S = 5) ==
 * measureSpinUps (FZ408L:600-647). umode 0: u[n][2] (second slot used only on the mixed branches);
 * umode 1: one sequential stream. Returns the number tagged. */
int orc_tag(int n, int S, const double* psi, const double* u, int umode, long* cursor, int* tagged) {
  int count = 0;
  for (int i = 0; i < n; i++) {
    double nrm[5];
    for (int k = 0; k < 5 && k < S; k++) {
      double re = psi[(i * S + k) * 2], im = psi[(i * S + k) * 2 + 1];
      nrm[k] = re * re + im * im;
    }
    double rnd = (umode == 0) ? u[2 * i] : u[(*cursor)++];
    int up;
    if (S == 7) {
      if (rnd < nrm[0] + nrm[2]) up = 1;
      else if (rnd < nrm[0] + nrm[2] + nrm[3]) { double r2 = (umode == 0) ? u[2 * i + 1] : u[(*cursor)++]; up = r2 < 2. / 3; }
      else if (rnd < nrm[0] + nrm[2] + nrm[3] + nrm[4]) { double r3 = (umode == 0) ? u[2 * i + 1] : u[(*cursor)++]; up = r3 < 1. / 3; }
      else up = 0;
    } else {
      if (rnd < nrm[0]) up = 1;
      else if (rnd < nrm[0] + nrm[2]) { double r2 = (umode == 0) ? u[2 * i + 1] : u[(*cursor)++]; up = r2 < 1. / 3; }
      else if (rnd < nrm[0] + nrm[2] + nrm[3]) { double r3 = (umode == 0) ? u[2 * i + 1] : u[(*cursor)++]; up = r3 < 2. / 3; }
      else up = 0;
    }
    tagged[i] = up;
    count += up;
  }
  return count;
}

/* FZ-family leap-frog pieces (FZ408L:317-369); the caller supplies F = forces() where the reference recomputes it */
void orc_lf_drift(int n, double* R, const double* V, const double* F, double L, double DT, int first) {
  for (int k = 0; k < 3 * n; k++) {
    if (first) R[k] += DT * V[k] + DT * DT * F[k]; /* FZ408L:336-338 */
    else R[k] += DT * V[k];                        /* FZ408L:325-327 */
    if (R[k] < 0) R[k] += L;
    if (R[k] > L) R[k] -= L;
  }
}
void orc_lf_kick(int n, double* V, const double* F, double DT) { /* FZ408L:364-366 */
  for (int k = 0; k < 3 * n; k++) V[k] += DT * F[k];
}
double orc_vaf(int n, const double* Vhold, const double* Vx) { /* FZ408L:955-960 */
  double vaf = 0.0;
  for (int j = 0; j < n; j++) vaf += 1 / ((double)(n)) * (Vhold[j] * Vx[j]);
  return vaf;
}

/* recordPairPairCorr (MD:584-652): counts[nbins] of ordered pairs, then the reference's normalisation (with its
 * integer sub-expressions N*4/3 and N*3). nbins = (int)(rmax/step). */
void orc_pair_correlation(int n, const double* R, double L, double step, double rmax, double* counts, double* g) {
  const int nbins = (int)(rmax / step);
  const double* X = R; const double* Y = R + n; const double* Z = R + 2 * n;
  for (int k = 0; k < nbins; k++) counts[k] = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      if (j == i) continue;
      double xd = X[i] - X[j], yd = Y[i] - Y[j], zd = Z[i] - Z[j];
      xd -= L * round(xd / L); yd -= L * round(yd / L); zd -= L * round(zd / L);
      double dist = sqrt(xd * xd + yd * yd + zd * zd);
      int bin = (int)floor((int)(dist / step));
      if (bin < nbins) counts[bin]++;
    }
  if (g)
    for (int i = 0; i < nbins; i++)
      g[i] = (i == 0) ? counts[i] / (n * 4 / 3 * M_PI * step * step * step) : counts[i] / (n * 3 * step * step * step * i * i);
}

/* recordVAF / recordLongViscAutoCorr / recordVCubeAutoCorr / recordVFourthAutoCorr (MD:654-823) on vStore[3][n][T];
 * nnorm = the N used in the normalisation (the reference's compile-time N; series beyond n are taken as zero).
 * which: 1 VAF, 2 long. viscosity, 3 v^3, 4 v^4. */
void orc_autocorr(int which, int n, int nnorm, int T, const double* vstore, double Gamma, double* out) {
  for (int tDiff = 0; tDiff < T; tDiff++) {
    double sum = 0;
    /* all-zero series beyond n contribute only the subtracted constant: added in closed form (test speed) */
    if (which == 2) sum += (double)(nnorm - n) * (T - tDiff) * (0.0 - 3 / (Gamma * Gamma));
    if (which == 4) sum += (double)(nnorm - n) * (T - tDiff) * (0.0 - 3 * 9 / (Gamma * Gamma * Gamma * Gamma));
    for (int i = 0; i < n; i++)
      for (int j = 0; j < T - tDiff; j++) {
        double a[3], b[3];
        for (int c = 0; c < 3; c++) {
          a[c] = vstore[((size_t)c * n + i) * T + j];
          b[c] = vstore[((size_t)c * n + i) * T + j + tDiff];
        }
        if (which == 1) sum += a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
        else if (which == 2)
          sum += pow(a[0], 2) * pow(b[0], 2) + pow(a[1], 2) * pow(b[1], 2) + pow(a[2], 2) * pow(b[2], 2) - 3 / (Gamma * Gamma);
        else if (which == 3)
          sum += pow(a[0], 2) * pow(b[0], 2) * a[0] * b[0] + pow(a[1], 2) * pow(b[1], 2) * a[1] * b[1] +
                 pow(a[2], 2) * pow(b[2], 2) * a[2] * b[2];
        else
          sum += pow(a[0], 2) * pow(b[0], 2) * pow(a[0], 2) * pow(b[0], 2) + pow(a[1], 2) * pow(b[1], 2) * pow(a[1], 2) * pow(b[1], 2) +
                 pow(a[2], 2) * pow(b[2], 2) * pow(a[2], 2) * pow(b[2], 2) - 3 * 9 / (Gamma * Gamma * Gamma * Gamma);
      }
    out[tDiff] = sum / (nnorm * (T - tDiff));
  }
}

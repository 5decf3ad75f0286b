// mdqt_qtconsts.h -- per-lane constants of the block-diagonal atom-light Hamiltonian, computed on the host once per
// handle with exactly the expressions (and summation order) the reference uses when it builds cs[], gs[],
// hamDecayTerm, decayMatrix and hamCouplingTermNoTimeDep (SU:1163-1215; MC408L:1171-1190, 597, 603-606).
#pragma once
#include <math.h>

namespace mdqt {

struct QTLane {
  double c10, c20, c13, c14, c25;  // real couplings H[P1][S], H[P2][S], H[P1][D3], H[P1][D4], H[P2][D5]
  double rot;                      // amplitude of the phase-rotating coupling H[D4][P2] = rot * exp(i phi) (SU:508)
  double gA, gB;                   // S-P optical-force weights (SU:503)
  double gD[4];                    // D-P optical-force weights (SU:503)
  double gam1, gam2;               // decayMatrix diagonal on P1, P2
  int map[6];                      // local state -> reference state index (0-based), -1 = unused
};

struct QTConsts {
  QTLane lane[2];
  double gam[4];   // decayMatrix diagonal on reference states 2,3,4,5
  double dfrac;    // dR/(dR+1): D-vs-S branching (SU:592)
  double tS[2];    // S-destination thresholds for jumps out of states 3 and 4 (SU:635, 660 / MC408L:710, 727)
  double tD[8];    // cumulative D-destination thresholds for jumps out of states 2,3,4,5 (SU:619-690)
  double h;        // dtQuant * gamToEinsteinFreq (SU:525)
  double kick_sp, kick_dp;  // vKick*Om*h and vKickDP*(OmDP/dR)*h (SU:503)
};

inline void fill_qt_consts(QTConsts& C, int scheme, double Om, double OmDP, double dR, double vKick, double vKickDP,
                           double dtq, double g2E, int quad) {
  QTConsts z = {};
  C = z;
  C.h = dtq * g2E;
  C.dfrac = dR / (dR + 1);
  if (scheme == 12) {
    double gs[18];
    gs[0] = sqrt(1.); gs[1] = sqrt(2. / 3); gs[2] = sqrt(1. / 3); gs[3] = sqrt(2. / 3); gs[4] = sqrt(1. / 3); gs[5] = sqrt(1.);
    gs[6] = sqrt(dR * 2. / 3); gs[7] = sqrt(dR * 4. / 15); gs[8] = sqrt(dR * 1. / 15); gs[9] = sqrt(dR * 2. / 5);
    gs[10] = sqrt(dR * 2. / 5); gs[11] = sqrt(dR * 1. / 5); gs[12] = sqrt(dR * 1. / 5); gs[13] = sqrt(dR * 2. / 5);
    gs[14] = sqrt(dR * 2. / 5); gs[15] = sqrt(dR * 1. / 15); gs[16] = sqrt(dR * 4. / 15); gs[17] = sqrt(dR * 2. / 3);
    static const int upper[18] = {2, 3, 3, 4, 4, 5, 5, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2};  // cs[k] = |lower><upper|
    double gam[12] = {0};
    for (int k = 0; k < 18; k++) gam[upper[k]] += gs[k] * gs[k];
    for (int m = 0; m < 4; m++) C.gam[m] = gam[2 + m];
    const double sd = sqrt(dR);
    QTLane& A = C.lane[0];
    QTLane& B = C.lane[1];
    const int mapA[6] = {0, 3, 5, 10, 8, 6}, mapB[6] = {1, 2, 4, 11, 9, 7};
    for (int k = 0; k < 6; k++) { A.map[k] = mapA[k]; B.map[k] = mapB[k]; }
    A.c10 = -1. * gs[2] * Om / 2;  A.c20 = -1. * gs[5] * Om / 2;
    A.c13 = -1. * gs[14] * OmDP / 2 / sd; A.c14 = -1. * gs[12] * OmDP / 2 / sd; A.c25 = -1. * gs[6] * OmDP / 2 / sd;
    A.rot = -(OmDP / 2 * gs[8] / sd);
    A.gA = gs[2]; A.gB = gs[5]; A.gD[0] = gs[8]; A.gD[1] = gs[14]; A.gD[2] = gs[6]; A.gD[3] = gs[12];
    A.gam1 = gam[3]; A.gam2 = gam[5];
    B.c10 = -1. * gs[0] * Om / 2;  B.c20 = -1. * gs[4] * Om / 2;
    B.c13 = -1. * gs[17] * OmDP / 2 / sd; B.c14 = -1. * gs[15] * OmDP / 2 / sd; B.c25 = -1. * gs[9] * OmDP / 2 / sd;
    B.rot = -(OmDP / 2 * gs[11] / sd);
    B.gA = gs[0]; B.gB = gs[4]; B.gD[0] = gs[11]; B.gD[1] = gs[17]; B.gD[2] = gs[9]; B.gD[3] = gs[15];
    B.gam1 = gam[2]; B.gam2 = gam[4];
    C.tS[0] = gs[2] * gs[2]; C.tS[1] = gs[4] * gs[4];
    C.tD[0] = gs[17] * gs[17] / dR; C.tD[1] = gs[17] * gs[17] / dR + gs[16] * gs[16] / dR;
    C.tD[2] = gs[14] * gs[14] / dR; C.tD[3] = gs[14] * gs[14] / dR + gs[13] * gs[13] / dR;
    C.tD[4] = gs[11] * gs[11] / dR; C.tD[5] = gs[11] * gs[11] / dR + gs[10] * gs[10] / dR;
    C.tD[6] = gs[8] * gs[8] / dR;   C.tD[7] = gs[8] * gs[8] / dR + gs[7] * gs[7] / dR;
    C.kick_sp = 1 * vKick * Om * dtq * g2E;
    C.kick_dp = vKickDP * (OmDP / dR) * dtq * g2E;
  } else if (scheme == 5) {  // 5-level 422 nm pump: gs are rates (MC422L:1144-1155); states S-1/2, S+1/2, P+1/2, P-1/2, D
    double gs[6] = {2. / 3, 1. / 3, 2. / 3, 1. / 3, dR, dR};
    static const int upper[6] = {2, 3, 3, 2, 2, 3};  // cs[k] = |lower><upper| (MC422L:1144-1149)
    double gam[5] = {0};
    for (int k = 0; k < 6; k++) gam[upper[k]] += gs[k];
    QTLane& A = C.lane[0];
    QTLane& B = C.lane[1];
    // lane A: S-1/2 <-> P-1/2 (energy -det + vq: the P2 slot); lane B: S+1/2 <-> P+1/2 (energy -det - vq: the P1 slot)
    const int mapA[6] = {0, -1, 3, 4, -1, -1}, mapB[6] = {1, 2, -1, -1, -1, -1};
    for (int k = 0; k < 6; k++) { A.map[k] = mapA[k]; B.map[k] = mapB[k]; }
    A.c20 = -Om / 2 * sqrt(gs[2]);  // |1><4| (MC422L:594)
    B.c10 = -Om / 2 * sqrt(gs[0]);  // |2><3|
    A.gam2 = gam[3]; B.gam1 = gam[2];
    C.gam[1] = gam[2]; C.gam[2] = gam[3];  // slots (A.P1, B.P1, A.P2, B.P2) = (-, state 2, state 3, -)
    C.tS[0] = gs[0]; C.tS[1] = gs[2];      // MC422L:685, 701
  } else if (scheme == 3) {  // 3-level J=0 <-> J=1 sigma+/- test system (TS:95-101, 379-382): ground, m=+1, m=-1
    double gs[2] = {1, 1};
    QTLane& A = C.lane[0];
    QTLane& B = C.lane[1];
    // one block: ground <-> state 2 (energy -det - v: P1 slot) and ground <-> state 1 (energy -det + v: P2 slot)
    const int mapA[6] = {0, 2, 1, -1, -1, -1}, mapB[6] = {-1, -1, -1, -1, -1, -1};
    for (int k = 0; k < 6; k++) { A.map[k] = mapA[k]; B.map[k] = mapB[k]; }
    A.c10 = -Om / 2 * sqrt(gs[0]);  // |1><3| sqrt(gs[0]) (TS:179)
    A.c20 = -Om / 2 * sqrt(gs[1]);  // |1><2| sqrt(gs[1])
    A.gam1 = gs[1]; A.gam2 = gs[0]; // cs[0] = |1><2|, cs[1] = |1><3| (TS:379-380)
    C.gam[0] = gs[1]; C.gam[2] = gs[0];
    A.gA = sqrt(gs[0]); A.gB = sqrt(gs[1]);
    C.kick_sp = 1 * vKick * Om * dtq * g2E;  // TS:174 (dt; the caller passes g2E = 1)
  } else {  // 7-level pump: gs are rates (MC408L:1181-1190)
    double gs[10] = {1, 2. / 3, 1. / 3, 1. / 3, 2. / 3, 1, dR, dR, dR, dR};
    static const int upper[10] = {2, 3, 4, 3, 4, 5, 2, 3, 4, 5};
    double gam[7] = {0};
    for (int k = 0; k < 10; k++) gam[upper[k]] += gs[k];
    for (int m = 0; m < 4; m++) C.gam[m] = gam[2 + m];
    QTLane& A = C.lane[0];
    QTLane& B = C.lane[1];
    const int mapA[6] = {0, 2, 4, 6, -1, -1}, mapB[6] = {1, 3, 5, -1, -1, -1};
    for (int k = 0; k < 6; k++) { A.map[k] = mapA[k]; B.map[k] = mapB[k]; }
    A.c10 = quad ? 0.0 : -Om / 2 * sqrt(gs[0]);  // |1><3|
    A.c20 = -Om / 2 * sqrt(gs[2]);               // |1><5|
    B.c10 = quad ? 0.0 : -Om / 2 * sqrt(gs[3]);  // |2><4|
    B.c20 = -Om / 2 * sqrt(gs[5]);               // |2><6|
    A.gam1 = gam[2]; A.gam2 = gam[4]; B.gam1 = gam[3]; B.gam2 = gam[5];
    C.tS[0] = gs[1]; C.tS[1] = gs[2];
  }
}

}  // namespace mdqt

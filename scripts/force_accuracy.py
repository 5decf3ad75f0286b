"""Developer aid: accuracy of the force kernel's pair arithmetic against a numpy float64 evaluation of SU:207-233 carried out in
extended precision (np.longdouble) for sampled rows, and against the C oracle for whole systems.
Usage: MDQT_LIB_PATH=... python scripts/force_accuracy.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqtplasmasims_b200 import Engine, su_params, synthetic


def rows_longdouble(R, L, lDeb, rows):
    out = {}
    Rl = R.astype(np.longdouble)
    for i in rows:
        d = Rl[:, i:i + 1] - Rl
        d -= L * np.round(d / L)
        r = np.sqrt((d ** 2).sum(axis=0))
        m = (r > 0) & (r < np.longdouble(L) / 2)
        ft = np.zeros(R.shape[1], dtype=np.longdouble)
        ft[m] = (1 / r[m] + 1 / np.longdouble(lDeb)) * np.exp(-r[m] / np.longdouble(lDeb)) / (r[m] * r[m])
        u = np.zeros_like(ft); u[m] = np.exp(-r[m] / np.longdouble(lDeb)) / r[m]
        out[i] = ((d * ft).sum(axis=1), (ft * r).sum(), u.sum())
    return out


for n, ge in ((3500, 0.1), (4096, 0.5), (20000, 0.1)):
    p = su_params(Ge=ge, n_ions=n, N0=n)
    R = synthetic.random_positions(n, p.L, seed=n + 1)
    eng = Engine(p)
    eng.upload(R=R)
    eng.forces()
    F = eng.download_forces()
    rows = np.random.default_rng(1).choice(n, 64, replace=False)
    ref = rows_longdouble(R, p.L, 1 / p.kappa, rows)
    e_abs = max(float(np.abs(F[:, i] - ref[i][0]).max() / ref[i][1]) for i in rows)      # per ion, relative to sum_j |f_ij|
    e_rel = max(float(np.abs(F[:, i] - ref[i][0]).max() / np.abs(ref[i][0]).max()) for i in rows)
    e_max = max(float(np.abs(F[:, i] - ref[i][0]).max()) for i in rows) / np.abs(F).max()
    print("N=%d Ge=%g kappa*L/2=%.2f: per-ion |dF|/sum_j|f_ij| %.2e, |dF|/|F_i| %.2e, |dF|/max|F| %.2e" %
          (n, ge, p.kappa * p.L / 2, e_abs, e_rel, e_max), flush=True)

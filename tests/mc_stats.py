"""Shared by tests/test_oracle_mc_programs.py (CPU) and tests/test_gpu_mc_programs.py (GPU): the statistical comparison of a set of
runs of the MC-tagging programs' stages 4-6 with the reference-run fixture tests/golden/mc_programs.npz (oracle/gen_golden.py
--mcprograms: 8 seeds per program of the reference's own collisional MD, pump stage, tagParticles() and recording stage at its
compile-time N = 4096).

What is compared (all intensive, so runs at another N can be held against the fixture with a wider tolerance):
  * the tagged fraction numTagged / N (the net result of the velocity-selective pump and the spin measurement),
  * the tagged ions' velocity moments <v_x>, <v_x^2> of taggedMoments.dat (MC408L:1069-1115) in three time windows of the
    recording stage -- the pump burns a velocity-dependent hole, the collisionless MD relaxes it,
  * temperature.dat (MC408L:771-790) in the same windows.
"""
import numpy as np

WINDOWS = ((0, 20), (90, 110), (180, 200))


def welch(a, b):
    """t statistic of the difference of the means of two sets of runs (first axis = run)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
    return (a.mean(axis=0) - b.mean(axis=0)) / se


def observables(frac, moments, temperature):
    """[runs][11]: tagged fraction, then per window (<v_x>, <v_x^2>, T). moments: [runs][steps][5], temperature: [runs][steps]."""
    cols = [np.asarray(frac, dtype=float)]
    names = ["tagged_fraction"]
    for lo, hi in WINDOWS:
        cols += [moments[:, lo:hi, 1].mean(axis=1), moments[:, lo:hi, 2].mean(axis=1), temperature[:, lo:hi].mean(axis=1)]
        names += ["m1[%d:%d]" % (lo, hi), "m2[%d:%d]" % (lo, hi), "T[%d:%d]" % (lo, hi)]
    return np.stack(cols, axis=1), names


def fixture(golden_dir, program):
    import os
    g = np.load(os.path.join(golden_dir, "mc_programs.npz"))
    n_ref = 4096
    obs, names = observables(g[program + "_ntag"] / n_ref, g[program + "_moments"], g[program + "_temperature"])
    return dict(obs=obs, names=names, npre=int(g["npre"]), nrec=int(g["nrec"]), npump=int(g[program + "_npump"]),
                ratio=int(g[program + "_ratio"]), pops=g[program + "_pops"], vel_dist0=g[program + "_vel_dist0"].astype(float))


def compare(ours, ref, names, tmax, rel):
    """Every observable: |Welch t| < tmax AND the means within `rel` (relative; a dict overrides per name prefix). The second
    bound keeps the test meaningful when the run-to-run scatter is tiny or huge. Returns the table for the failure message."""
    t = welch(ours, ref)
    mo, mr = ours.mean(axis=0), ref.mean(axis=0)
    rows = []
    ok = True
    for k, nm in enumerate(names):
        r = rel.get(nm.split("[")[0], rel["default"]) if isinstance(rel, dict) else rel
        scale = max(abs(mr[k]), 1e-3)
        good = abs(t[k]) < tmax and abs(mo[k] - mr[k]) <= r * scale
        ok = ok and good
        rows.append((nm, float(mo[k]), float(mr[k]), float(t[k]), bool(good)))
    return ok, rows

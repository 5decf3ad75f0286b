"""GPU suite: statistical parity of the COUPLED run (north-star: "jump statistics, temperatures and state populations must
agree statistically over ensembles"). Fixture tests/golden/su_coupled.npz = 16 trajectories of the reference's own main loop
(lasers on, forces on, its drand48 stream, 1 thread; oracle/gen_golden.py --coupled): N0 = 500, tmax = 2.4, energies.dat rows
(SU:934-955) and the mean S/P/D populations of statePopulationsVsVTime (SU:1012-1024) at its 30 output() calls.
The engine runs 64 jobs through `mdqt_run --jobs` (the first 16 start from the very same init() states; the stochastic part
-- Philox instead of drand48 -- and the chaotic N-body dynamics make the trajectories statistically independent)."""
import os
import subprocess

import numpy as np
import pytest

from mdqtplasmasims_b200 import Engine, hostio, su_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run")

OBS = ["ekin_x", "ekin_y", "ekin_z", "epot", "vx_avg", "pop_S", "pop_P", "pop_D", "std_vx"]


def _welch(a, b):
    """t statistic of the difference of the ensemble means, a: [seeds_a, ...], b: [seeds_b, ...]."""
    se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
    return (a.mean(axis=0) - b.mean(axis=0)) / se


def test_coupled_run_statistics_match_the_reference(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "su_coupled.npz"))
    N0, tmax, nrow = int(g["N0"]), float(g["tmax"]), g["energies"].shape[1]
    ref = np.concatenate([g["energies"][:, :, [1, 2, 3, 4, 6]], g["pops"]], axis=2)  # [16][30][9]
    save = str(tmp_path) + "/"
    first, njobs = int(g["seeds"][0]), 64
    r = subprocess.run([DRIVER, "--jobs", "%d-%d" % (first, first + njobs - 1), "--batch", "64", "--seed", "0", "--N0", str(N0),
                        "--tmax", str(tmax), "--saveDirectory", save, "--quiet"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    ours = np.zeros((njobs, nrow, 9))
    for k in range(njobs):
        d = hostio.dirname(save, N0=N0, job=first + k)
        en = np.loadtxt(os.path.join(d, "energies.dat"), ndmin=2)
        assert en.shape == (nrow, 7)
        if k < 16:  # same init() states as the reference trajectories: same N, same output times
            assert int(open(os.path.join(d, [f for f in os.listdir(d) if f.startswith("ions_")][0])).read().split()[0]) == int(g["N"][k])
            assert np.array_equal(en[:, 0], g["energies"][k, :, 0])
        ours[k, :, :5] = en[:, [1, 2, 3, 4, 6]]
        for j in range(nrow):
            a = np.loadtxt(os.path.join(d, "statePopulationsVsVTime%06d.dat" % j), ndmin=2)
            ours[k, j, 5:] = [a[:, 1].mean(), a[:, 2].mean(), a[:, 3].mean(), a[:, 0].std()]
    # (1) every observable at every output time: Welch t of the ensemble means (16 reference vs 64 engine trajectories)
    t_all = _welch(ours, ref)
    assert np.abs(t_all).max() < 5.5, [(OBS[c], j, t_all[j, c]) for j, c in zip(*np.where(np.abs(t_all) >= 5.5))]
    # (2) time-window averages (disorder-induced heating t < 0.5, the kinetic-energy oscillation, the late plateau): tighter
    for lo, hi in ((0, 6), (6, 16), (16, nrow)):
        tw = _welch(ours[:, lo:hi].mean(axis=1), ref[:, lo:hi].mean(axis=1))
        assert np.abs(tw).max() < 4.5, (lo, hi, dict(zip(OBS, tw)))
    # (3) no systematic offset: the t values of an observable scatter about 0 along the run
    assert np.abs(t_all.mean(axis=0)).max() < 3.0, dict(zip(OBS, t_all.mean(axis=0)))
    # (4) the physics both must show: heating out of the frozen start, P population ~0.19, D filling up, T_x < T_y,z late (x is cooled)
    m = ours.mean(axis=0)
    assert m[10, 0] > 5 * m[0, 0] and 0.15 < m[-1, 6] < 0.22 and m[-1, 7] > 2 * m[0, 7]
    rel = np.abs(ours.mean(axis=0) - ref.mean(axis=0))[:, [0, 1, 2, 3, 5, 6, 7]] / np.abs(ref.mean(axis=0))[:, [0, 1, 2, 3, 5, 6, 7]]
    assert rel[3:].max() < 0.12  # ensemble means within 12 % everywhere after the first outputs (statistical scatter ~5 %)


def test_coupled_run_jump_rate_and_norm_match_the_reference(golden_dir):
    """The same 16 jobs through the Engine (one batched handle): the fraction of ions that jumped within the last MD step
    and the mean final norm (the reference lets it drift, reNormalizewvFns = false) against the reference trajectories."""
    g = np.load(os.path.join(golden_dir, "su_coupled.npz"))
    N0 = int(g["N0"])
    sts = [hostio.init_su(int(s), N0=N0) for s in g["seeds"]]
    B, cap = len(sts), max(s["N"] for s in sts)
    assert [s["N"] for s in sts] == [int(n) for n in g["N"]]
    p = su_params(n_ions=cap, N0=N0, n_traj=B, traj0=int(g["seeds"][0]), plan_n=N0)
    R, V, psi, tp = np.zeros((B, 3, cap)), np.zeros((B, 3, cap)), np.zeros((B, cap, 12, 2)), np.zeros((B, cap))
    psi[:, :, 0, 0] = 1.0
    for b, s in enumerate(sts):
        n = s["N"]
        R[b, :, :n], V[b, :, :n], psi[b, :n] = s["R"], s["V"], s["psi"]
    e = Engine(p)
    e.set_ion_counts([s["N"] for s in sts])
    e.set_traj_seeds(np.asarray(g["seeds"], dtype=np.uint64))
    e.upload(R=R, V=V, psi=psi, tPart=tp, t=0.0, substep=0)
    nmd = int(g["nsub"][0]) // 25
    for _ in range(nmd // 40):
        e.md_steps(40)
    if nmd % 40:
        e.md_steps(nmd % 40)
    s = e.download()
    recent = np.array([(s["tPart"][b, :sts[b]["N"]] < 25 * p.dtq * 0.999).mean() for b in range(B)])
    norm = np.array([(s["psi"][b, :sts[b]["N"]] ** 2).sum(axis=(1, 2)).mean() for b in range(B)])
    t_recent = _welch(recent[:, None], g["recent_jump_frac"][:, None])[0]
    t_norm = _welch(norm[:, None], g["norm_final"][:, None])[0]
    assert abs(t_recent) < 4.0 and abs(t_norm) < 4.0, (t_recent, t_norm, recent.mean(), g["recent_jump_frac"].mean())
    assert abs(recent.mean() - 0.049) < 0.01  # 1 - exp(-25 h Gamma <popP>) with <popP> ~ 0.19
    e.close()

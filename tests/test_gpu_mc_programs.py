"""GPU suite: `mdqt_run --program mc408l | mc422l` (csrc/mdqt_programs.cpp) against runs of the REFERENCE's own programs.

Fixture tests/golden/mc_programs.npz (oracle/gen_golden.py --mcprograms): 8 seeds per program of main()'s stages 4-6 executed
by the reference's own functions at its compile-time N = 4096 -- 200 collisional MD steps, the pump stage with the reference's
pumpMDTimeSteps x plasmaToQuantumTimestepRatio sweeps (46 x 62 for 408 nm, 12 x 55 for 422 nm), tagParticles(), 200 recorded
collisionless steps with taggedMoments.dat / temperature.dat / vel_distX_timestep%06d.dat (MC408L:1211-1244, MC422L:1178-1211).
The reference draws from mt19937 and drand48, the engine from Philox, so the comparison is statistical (tests/mc_stats.py):
16 engine runs at the same N and step counts against the 8 reference runs -- tagged fraction, the tagged ions' <v_x>, <v_x^2>
and the temperature in three windows of the recording stage, and the tagged ions' velocity distribution right after the
measurement. The CPU twin (tests/test_oracle_mc_programs.py) holds the oracle against the same fixture."""
import os
import subprocess

import numpy as np
import pytest

import mc_stats

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run")
NRUNS, N = 16, 4096


def _table(path):
    return np.loadtxt(path, ndmin=2)


@pytest.mark.parametrize("program", ["mc408l", "mc422l"])
def test_mc_program_runs_match_the_reference_runs_statistically(tmp_path, golden_dir, program):
    fx = mc_stats.fixture(golden_dir, program)
    frac, mom, temp, vd0 = [], [], [], []
    for k in range(NRUNS):
        save = str(tmp_path / ("run%d" % k)) + "/"
        r = subprocess.run([DRIVER, "--program", program, str(k + 1), "--N", str(N), "--seed", str(7000 + 13 * k), "--preSteps", str(fx["npre"]),
                            "--recordSteps", str(fx["nrec"]), "--dateSuffix", "0", "--saveDirectory", save, "--quiet"],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        top = os.listdir(save)
        assert len(top) == 1
        d = os.path.join(save, top[0], "job%d" % (k + 1))
        tm = _table(os.path.join(d, "taggedMoments.dat"))
        assert tm.shape == (fx["nrec"], 5)
        mom.append(tm)
        temp.append(_table(os.path.join(d, "temperature.dat"))[:, 0])
        pv = _table(os.path.join(d, "vel_distX_timestep%06d.dat" % 0))[:, 1]
        ntag = pv.sum() * 0.0025 * 6                 # the KDE integrates to numTagged / 6 (MC408L:1097-1125)
        assert abs(ntag - round(ntag)) < 1e-3
        frac.append(round(ntag) / N)
        vd0.append(pv)
    obs, names = mc_stats.observables(np.array(frac), np.stack(mom), np.stack(temp))
    # the default pump length is the reference's pumpMDTimeSteps (MC408L:119-120): nothing is passed on the command line
    ok, rows = mc_stats.compare(obs, fx["obs"], names, tmax=6.0, rel=dict(default=0.05, m1=0.25, T=0.03))
    print("\n".join("%-18s ours %.6g  reference %.6g  t %+.2f  %s" % r for r in rows))
    assert ok, rows
    # the tagged ions' velocity distribution at the first recorded step, pooled over the runs: Kolmogorov-Smirnov distance of
    # the two cumulative distributions (16 x ~2000 against 8 x ~2000 ions: D(alpha = 1e-3) = 0.017)
    a, b = np.stack(vd0).sum(axis=0), fx["vel_dist0"].sum(axis=0)
    ca, cb = np.cumsum(a) / a.sum(), np.cumsum(b) / b.sum()
    assert np.abs(ca - cb).max() < 0.02, np.abs(ca - cb).max()

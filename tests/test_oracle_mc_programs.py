"""CPU suite: the oracle restatement, sequenced as stages 4-6 of the MC-tagging programs' main() (MC408L:1211-1244,
MC422L:1178-1211: collisional MD, the pump stage {ratio x qstep(); MDStep(k)}, tagParticles(), the recording stage), against the
reference-run fixture tests/golden/mc_programs.npz -- statistically (tests/mc_stats.py), because the reference draws from
mt19937 / drand48 and everything here from numpy streams. It pins the oracle pieces (7- and 5-level sweeps, spin measurement,
velocity Verlet with Andersen collisions) TOGETHER, in the programs' order, to what the reference's programs produce; the GPU
twin of this test (tests/test_gpu_mc_programs.py) holds `mdqt_run --program mc408l|mc422l` against the same fixture.

The oracle runs at N = 512 (8 seeds in a few seconds per program); the fixture is the reference's compile-time N = 4096. All
compared quantities are intensive; the smaller box (cut-off L/2 = 6.4 instead of 12.9 screening lengths/2) moves the temperature
by a fraction of a percent, which the tolerance below covers."""
import os

import numpy as np
import pytest

import mc_stats
from oracle import pyoracle as po

N, SEEDS = 512, range(8)
GAMMA, KAPPA, DT, COLL = 3.0, 0.5, 0.005, 0.25


def oracle_stages(orc, program, seed, npre, npump, nrec):
    rng = np.random.default_rng(1000 + seed)
    p, ratio = po.mc408_params() if program == "mc408l" else po.mc422_params()
    S = 7 if program == "mc408l" else 5
    sweep = orc.qstep7 if program == "mc408l" else orc.qstep5
    L = (N * 4. * np.pi / 3.) ** (1. / 3)
    side = round(N ** (1. / 3))
    g = np.arange(side) * L / side + 0.5
    R = np.ascontiguousarray(np.stack(np.meshgrid(g, g, g, indexing="ij")).reshape(3, N))          # init(): MC408L:209-250
    V = rng.normal(0, np.sqrt(1 / GAMMA), size=(3, N))
    r1, r2 = rng.uniform(size=N), rng.uniform(size=N)
    s1, s2 = np.where(rng.uniform(size=N) < 0.5, -1.0, 1.0), np.where(rng.uniform(size=N) < 0.5, -1.0, 1.0)
    psi = np.zeros((N, S, 2))
    psi[:, 0, 0] = np.sqrt(r1)
    psi[:, 1, 0] = s2 * np.sqrt(1 - r1) * np.sqrt(r2)
    psi[:, 1, 1] = s1 * np.sqrt(1 - r1) * np.sqrt(1 - r2)
    A = np.zeros((3, N))                                                                            # the reference starts from A = 0

    def mdstep(coll):
        nonlocal A
        old = A.copy()
        orc.vv_positions(R, V, A, L, DT)
        A = orc.forces_md(R, L, KAPPA, L / 2)
        if coll > 0:
            orc.vv_velocities(V, old, A, DT, coll, rng.uniform(size=N), np.ascontiguousarray(rng.normal(0, np.sqrt(1 / GAMMA), size=(N, 3))))
        else:
            orc.vv_velocities(V, old, A, DT)

    for _ in range(npre):
        mdstep(COLL)
    for _ in range(npump):
        for _ in range(ratio):
            vx = V[0].copy()
            sweep(psi, vx, p, rng.uniform(size=(N, 5)))
        mdstep(0.0)
    tagged, _ = orc.tag(psi, rng.uniform(size=(N, 2)))
    m = tagged.astype(bool)
    mom, temp = np.zeros((nrec, 5)), np.zeros(nrec)
    for k in range(nrec):
        vt = V[0][m]
        mom[k] = [k * DT, vt.mean(), (vt ** 2).mean(), (vt ** 3).mean(), (vt ** 4).mean()]              # MC408L:1088-1115
        temp[k] = (V ** 2).sum() / (3.0 * N)                                                            # MC408L:771-790
        mdstep(0.0)
    return m.sum() / N, mom, temp, (psi ** 2).sum(axis=2).mean(axis=0)


@pytest.mark.parametrize("program", ["mc408l", "mc422l"])
def test_oracle_program_stages_match_the_reference_runs(oracle, golden_dir, program):
    fx = mc_stats.fixture(golden_dir, program)
    p, ratio = po.mc408_params() if program == "mc408l" else po.mc422_params()
    assert ratio == fx["ratio"]                                                # plasmaToQuantumTimestepRatio (MC408L:116, MC422L:114)
    runs = [oracle_stages(oracle, program, s, fx["npre"], fx["npump"], fx["nrec"]) for s in SEEDS]
    obs, names = mc_stats.observables(np.array([r[0] for r in runs]), np.stack([r[1] for r in runs]), np.stack([r[2] for r in runs]))
    ok, rows = mc_stats.compare(obs, fx["obs"], names, tmax=6.0, rel=dict(default=0.10, m1=0.60, T=0.05))  # 8 runs of 512 ions: <v_x> of ~240 tagged ions scatters by 0.03 per run
    assert ok, rows
    # level populations after the pump (they do not change in the recording stage)
    pops = np.stack([r[3] for r in runs])
    assert np.abs(pops.mean(axis=0) - fx["pops"].mean(axis=0)).max() < 0.015, (pops.mean(axis=0), fx["pops"].mean(axis=0))

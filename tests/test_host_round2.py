"""CPU suite, round 2: host-side pieces added this round -- the re-entrant init() replica (jobs of an ensemble are initialised from
several threads), the coupled and MD-program fixtures' self-consistency against the live reference harness, and the reference
arm's JSON line."""
import gzip
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from mdqtplasmasims_b200 import hostio
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_init_su_is_reentrant():
    """mdqt_io_init_su draws from a local erand48 state (the same 48-bit LCG as the reference's global drand48 stream): the states
    of eight jobs initialised concurrently equal the ones initialised one after the other, bit for bit."""
    seeds = list(range(300, 308))
    serial = [hostio.init_su(s, N0=700) for s in seeds]
    out = [None] * len(seeds)

    def work(k):
        out[k] = hostio.init_su(seeds[k], N0=700)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(seeds))]
    [t.start() for t in th]
    [t.join() for t in th]
    for a, b in zip(serial, out):
        assert a["N"] == b["N"] and np.array_equal(a["R"], b["R"]) and np.array_equal(a["psi"], b["psi"])
    assert len({a["N"] for a in serial}) > 1  # the jobs drew different ion numbers


def test_coupled_fixture_is_sane(golden_dir):
    """tests/golden/su_coupled.npz (16 reference trajectories of the coupled loop): start states reproducible from the seeds, output
    times on the reference's schedule, disorder-induced heating, P population near its steady state, jump fraction consistent with it."""
    g = np.load(os.path.join(golden_dir, "su_coupled.npz"))
    N0 = int(g["N0"])
    assert g["energies"].shape[0] == 16 and g["energies"].shape[2] == 7 and g["pops"].shape[:2] == g["energies"].shape[:2]
    assert [hostio.init_su(int(s), N0=N0)["N"] for s in g["seeds"][:4]] == [int(n) for n in g["N"][:4]]
    t = g["energies"][0, :, 0]
    assert np.allclose(np.diff(t), 0.08, atol=1e-6) and abs(t[0] - 0.07808) < 1e-5      # output() every 40 MD steps, after the first substep
    ek = g["energies"][:, :, 1:4].mean(axis=0)
    assert ek[10].min() > 5 * ek[0].max()                                                # heating out of the frozen start
    popP = g["pops"][:, -1, 1].mean()
    assert 0.15 < popP < 0.22
    h, Gam = 0.002 / 25 * 174.07 / np.sqrt(2.0), 1.0617
    assert abs(g["recent_jump_frac"].mean() - (1 - np.exp(-25 * h * Gam * popP))) < 0.01  # jump rate = Gamma popP per unit time


@pytest.mark.skipif(not po.ref_available("md"), reason="oracle/_ref/libref_md.so not present")
def test_md_program_fixture_matches_the_reference_init_and_tags(golden_dir):
    """The MD-program fixture pins more than files: the reference's init() with std::mt19937 seeded 4321 and its tagParticles()
    are reproducible here, and row 0 of taggedV{One..Four}Moments.dat is the moment set of the initial velocities over those tags
    (MD:923-1003) -- the same numbers `mdqt_run --program md` must produce on the GPU."""
    gdir = os.path.join(golden_dir, "md_program")
    tags = np.load(os.path.join(gdir, "tags.npy"))
    md = po.RefMD()
    md.seed(4321)
    md.init()
    V = md.get_state()["V"]
    Gamma = md.consts["Gamma"]
    for k, name in enumerate(("One", "Two", "Three", "Four")):
        row = [float(x) for x in gzip.open(os.path.join(gdir, "taggedV%sMoments.dat.gz" % name), "rt").readline().split()]
        m = ((tags >> k) & 1).astype(bool)
        v = V[0][m]
        want = [0.0, v.mean(), (v ** 2).mean() - 1 / Gamma, (v ** 3).mean(), (v ** 4).mean() - 3 / Gamma ** 2]
        assert np.allclose(row, want, rtol=2e-5, atol=2e-6), (name, row, want)
    t0 = float(gzip.open(os.path.join(gdir, "temperature.dat.gz"), "rt").readline())
    assert abs(t0 - (V ** 2).mean()) < 2e-6 * t0 + 1e-9


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: one JSON line with the arm's keys; every step is whole MD steps run for real (here bounded to a
    few seconds through MDQT_REF_BUDGET_S)."""
    env = dict(os.environ, MDQT_REF_BUDGET_S="4", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], env=env,
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "ion-steps/s" and d["higher_is_better"] is True and d["value"] > 1e3
    assert d["config"]["md_steps_per_step"] >= 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"]["n_ions"] * 25.0 * d["config"]["md_steps_per_step"]) < 1e-6 * d["value"] * d["ms_per_step"]


@pytest.mark.parametrize("program,args,tree", [
    ("md", ["--N", "512"], "Gamma300Kappa50NumIons512/job3"),                                                        # MD:1037-1058
    ("mc408l", ["--N", "512"], "Gamma300Kappa50NumIons512PumpTime200Det250Om70Density20/job3"),                      # MC408L:1153
    ("fz408l", ["--N0", "700"], "PumpTime200PumpStart15Det250Om70Density20Ge100NumIons700/job3"),                    # FZ408L:990
    ("fz408q", ["--N0", "700"], "PumpTime100PumpStart15Det0Om200Density20Ge100NumIons700/job3"),                     # FZ408Q:58-60, 999
    ("fz422l", ["--N0", "700"], "PumpTime100PumpStart15Det100Om130Density20Ge100NumIons700/job3"),                   # FZ422L:55-57, 955
    ("ts", ["--N0", "100"], "Om50/Det-50NumIons100InitialTemp10000uK/job3"),                                         # TS:371-382
])
def test_program_drivers_build_the_reference_tree_and_fail_loudly_without_a_gpu(tmp_path, program, args, tree):
    """`mdqt_run --program <p>`: every host loop first makes the reference's directory tree from the reference's DEFAULT inputs
    (so the names below are what the reference's own main() would create; TS prints its negative detuning through an (unsigned)
    cast as -50 on x86-64) and then needs a CUDA device -- without one it must stop with the library's error, never compute on the CPU."""
    from mdqtplasmasims_b200 import load_library
    if load_library().mdqt_device_count() > 0:
        pytest.skip("a GPU is present")
    save = str(tmp_path) + "/"
    r = subprocess.run([os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run"), "--program", program, "3", "--seed", "1", "--saveDirectory", save,
                        "--quiet"] + args, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CUDA device" in r.stderr, (r.returncode, r.stderr[-500:])
    assert os.path.isdir(os.path.join(save, tree)), [os.path.join(dp, d) for dp, ds, _ in os.walk(save) for d in ds]
    assert not any(fs for _, _, fs in os.walk(save))          # nothing was written


def test_unknown_program_is_a_usage_error():
    r = subprocess.run([os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run"), "--program", "nosuch", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "unknown program" in r.stderr

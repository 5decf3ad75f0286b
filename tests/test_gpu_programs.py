"""GPU suite: the C++ host loops of the reference's other programs (`mdqt_run --program md|fz408l`, csrc/mdqt_programs.cpp) and
the files they write, against files written by the reference's own recorders (tests/golden/md_program) and against the
golden-tested Python loop on the same engine (FZ408L)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from mdqtplasmasims_b200 import Engine, SCHEME_CA5, SCHEME_NONE, SCHEME_SR7, drivers, hostio, md_params, su_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run")


def _table(text):
    return np.array([[float(x) for x in line.split()] for line in text.strip().splitlines()])


def test_md_program_files_match_the_reference_recorders(tmp_path, golden_dir):
    """`mdqt_run --program md` with the fixture's step counts (no collisional stages: deterministic) and std::mt19937 seed:
    same lattice + Maxwellian init(), same velocity-dependent tags, then taggedV*Moments.dat, temperature.dat, g(r) and the three
    TemperaturesAlongAxes files as the reference's recordTaggedParticleMoments / recordTemperature / recordPairPairCorr /
    recordTempForEachAxis wrote them (MD:525-582, 584-652, 923-1029)."""
    gdir = os.path.join(golden_dir, "md_program")
    save = str(tmp_path) + "/"
    r = subprocess.run([DRIVER, "--program", "md", "1", "--seed", "4321", "--preSteps", "0", "--recordSteps", "40", "--instSteps", "30",
                        "--reequilSteps", "0", "--establishSteps", "20", "--relaxSteps", "20", "--saveDirectory", save, "--quiet"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    d = os.path.join(save, "Gamma300Kappa50NumIons4096", "job1")
    gold = {f[:-3]: gzip.open(os.path.join(gdir, f), "rt").read() for f in os.listdir(gdir) if f.endswith(".gz")}
    assert len(gold) >= 9
    for f, g in gold.items():
        path = os.path.join(d, f)
        assert os.path.exists(path), f
        a, b = _table(open(path).read()), _table(g)
        assert a.shape == b.shape, f
        lo, lg = open(path).read().strip().splitlines(), g.strip().splitlines()
        same = sum(x == y for x, y in zip(lo, lg))
        if f.startswith("pairPairCorr"):
            assert np.array_equal(a, b), f                       # integer pair counts, the reference's normalisation: exact text
        elif f.startswith("tagged"):
            # means over ~2000 tagged ions minus an O(1) equilibrium constant: cancellation leaves ~1e-13 absolute noise
            assert np.allclose(a, b, rtol=2e-5, atol=2e-9) and same >= 0.85 * len(lg), (f, same, len(lg))
        else:
            assert np.allclose(a, b, rtol=2e-6, atol=1e-12) and same >= 0.95 * len(lg), (f, same, len(lg))
    # the autocorrelation files exist with one row per recorded step (their values: test_gpu_schemes autocorrelation goldens)
    for f in ("VAF.dat", "longViscAutoCorr.dat", "vCubeAutoCorr.dat", "vFourthAutoCorr.dat"):
        assert _table(open(os.path.join(d, f)).read()).shape == (40, 2)


def test_md_program_tags_and_moments_api(golden_dir):
    """mdqt_set_tags / mdqt_moments_record against numpy on the fixture's tag sets."""
    tags = np.load(os.path.join(golden_dir, "md_program", "tags.npy"))
    n = tags.shape[0]
    rng = np.random.default_rng(5)
    V = rng.normal(size=(3, n)) * 0.6
    e = Engine(md_params(scheme=SCHEME_NONE, n_ions=n))
    e.upload(R=rng.uniform(0, e.params.L, size=(3, n)), V=V)
    e.set_tags(tags)
    e.moments_begin(3)
    e.moments_record(1)
    e.scale_velocities(1.1, 0.9, 0.8)
    e.moments_record(2)
    rec = e.moments_download(3)
    assert np.all(rec[0] == 0)
    for slot, W in ((1, V), (2, V * np.array([1.1, 0.9, 0.8])[:, None])):
        assert np.allclose(rec[slot][:3], (W ** 2).sum(axis=1), rtol=1e-13)
        for k in range(4):
            m = ((tags >> k) & 1).astype(bool)
            want = [m.sum()] + [(W[0][m] ** p).sum() for p in (1, 2, 3, 4)]
            assert np.allclose(rec[slot][3 + 5 * k:8 + 5 * k], want, rtol=1e-12, atol=1e-12)
    e.close()


def test_fz408l_program_files_match_the_python_loop(tmp_path):
    """`mdqt_run --program fz408l` against drivers.fz_main_loop (itself pinned to the reference's loop by
    tests/golden/fz408l_loop.npz) on the same engine calls: same spin-up list, same VAF.dat rows, same final positions."""
    save = str(tmp_path) + "/"
    args = dict(N0=700, seed=99, tstartV0=0.0061, tpumpreal=2e-9, tmax=0.0201, sampleFreq=5)
    r = subprocess.run([DRIVER, "--program", "fz408l", "2", "--N0", "700", "--seed", "99", "--tstartV0", "0.0061", "--tpumpreal", "2e-9",
                        "--tmax", "0.0201", "--sampleFreq", "5", "--saveDirectory", save, "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    sub = [x for x in os.listdir(save)][0]
    d = os.path.join(save, sub, "job2")
    files = os.listdir(d)
    c0 = int([f for f in files if f.startswith("ions_")][0][len("ions_timestep"):-4])
    # the same run through the Python loop
    st = hostio.init_su(99, N0=700)
    n = st["N"]
    p = su_params(Ge=0.1, density=2.0, detuning=-2.5, detuningDP=0.0, Om=0.7, OmDP=0.0, N0=700, n_ions=n, scheme=SCHEME_SR7, traj0=2, seed=99)
    p.substeps_per_md = int(round(34.81 / np.sqrt(2.0)))
    p.dtq = 0.002 / p.substeps_per_md
    e = Engine(p)
    psi = np.zeros((n, 7, 2))
    psi[:, :2] = st["psi"][:, :2]
    e.upload(R=st["R"], V=st["V"], psi=psi, t=0.0, substep=0)
    tend = 0.0061 + 2e-9 * 813490 * np.sqrt(2.0)
    events = []
    out = drivers.fz_main_loop(e, 0.0201, 0.0061, tend, sampleFreq=5, on_measure=lambda t, tg, nu, v: events.append((t, v)),
                               on_sample=lambda t, c, v: events.append((t, v)))
    assert out["c0"] == c0
    spin = np.loadtxt(os.path.join(d, "spinUpIonsList_timestep%06d.dat" % c0), dtype=int)
    assert np.array_equal(spin, out["tagged"])
    vaf = _table(open(os.path.join(d, "VAF.dat")).read())
    assert vaf.shape == (len(events), 2)
    assert np.allclose(vaf, np.array(events), rtol=6e-6, atol=1e-300)
    s = e.download()
    cond = _table(open(os.path.join(d, "conditions_timestep%06d.dat" % c0)).read())
    assert np.allclose(cond[:, :3], s["R"].T, rtol=6e-6) and np.allclose(cond[:, 3:6], s["V"].T, rtol=6e-6, atol=1e-12)  # %lg: 6 digits
    en = _table(open(os.path.join(d, "energies.dat")).read())
    assert en.shape == (len(events), 6) and np.all(np.abs(en[:, 5]) < 1e-3)   # energy conserved from the frozen start
    tm = _table(open(os.path.join(d, "taggedMoments.dat")).read())
    assert tm.shape == (len(events), 5)
    assert any(f.startswith("vel_distX_timestep") for f in files)
    e.close()


def test_vel_dist_tagged_vs_numpy():
    """mdqt_vel_dist_tagged against the formula of recordTaggedParticleMoments (MC408L:1097-1125): Gaussian weights of width 0.002 on
    4001 bins of 0.0025 centred on 0, tagged ions only, / (6 sqrt(2 pi 0.002^2))."""
    n = 512
    rng = np.random.default_rng(8)
    V = rng.normal(size=(3, n)) * 0.4
    tags = (rng.uniform(size=n) < 0.4).astype(np.uint8) | ((rng.uniform(size=n) < 0.5).astype(np.uint8) << 1)
    e = Engine(md_params(scheme=SCHEME_NONE, n_ions=n))
    e.upload(R=rng.uniform(0, e.params.L, size=(3, n)), V=V)
    e.set_tags(tags)
    pv = e.vel_dist_tagged()
    vel = (np.arange(4001) - 2000) * 0.0025
    vt = V[0][(tags & 1).astype(bool)]
    want = np.exp(-(vel[:, None] - vt[None, :]) ** 2 / (2 * 0.002 ** 2)).sum(axis=1) / (6.0 * np.sqrt(2 * np.pi * 0.002 ** 2))
    big = want > 1e-250  # deep in the tails the sums are subnormal: exp() implementations differ there
    assert big.sum() > 500 and np.allclose(pv[big], want[big], rtol=1e-11) and np.all(pv[~big] < 1e-240)
    e.close()


@pytest.mark.parametrize("program,prefix", [("mc408l", "Gamma300Kappa50NumIons512PumpTime200Det250Om70Density20"),
                                            ("mc422l", "Gamma300Kappa50NumIons512PumpTime")])
def test_mc_tagging_programs_write_consistent_files(tmp_path, program, prefix):
    """`mdqt_run --program mc408l | mc422l` (collisional MD, pump with the 7- resp. 5-level scheme, spin measurement, recording
    stage): the files exist in the reference's shapes, and the two things it writes about the tagged ions agree with each other --
    the integral of vel_distX_timestep k is numTagged / 6, and its first and second moments are the taggedMoments.dat row k."""
    save = str(tmp_path) + "/"
    r = subprocess.run([DRIVER, "--program", program, "3", "--N", "512", "--seed", "17", "--preSteps", "20", "--pumpSteps", "6", "--recordSteps", "12",
                        "--saveDirectory", save, "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    runs = [x for x in os.listdir(save) if x.startswith(prefix)]
    assert len(runs) == 1, os.listdir(save)
    assert ("Date" in runs[0]) == (program == "mc422l")   # MC422L:1127-1134 appends the date to the run directory
    d = os.path.join(save, runs[0], "job3")
    assert os.path.isdir(d), os.listdir(save)
    tm = _table(open(os.path.join(d, "taggedMoments.dat")).read())
    assert tm.shape == (12, 5) and np.allclose(tm[:, 0], np.arange(12) * 0.005)
    assert _table(open(os.path.join(d, "temperature.dat")).read()).shape == (12, 1)
    assert os.path.exists(os.path.join(d, "pairPairCorrStepNum0.dat"))
    for f in ("VAF.dat", "longViscAutoCorr.dat", "vCubeAutoCorr.dat", "vFourthAutoCorr.dat"):
        assert _table(open(os.path.join(d, f)).read()).shape == (12, 2)
    for k in (0, 11):
        pv = _table(open(os.path.join(d, "vel_distX_timestep%06d.dat" % k)).read())
        assert pv.shape == (4001, 2)
        ntag = pv[:, 1].sum() * 0.0025 * 6
        assert abs(ntag - round(ntag)) < 1e-3 and 0.05 * 512 < ntag < 0.95 * 512   # an integer number of tagged ions, a sizeable fraction
        m1 = (pv[:, 0] * pv[:, 1]).sum() / pv[:, 1].sum()
        m2 = (pv[:, 0] ** 2 * pv[:, 1]).sum() / pv[:, 1].sum() - 0.002 ** 2         # minus the kernel's own variance
        assert abs(m1 - tm[k, 1]) < 2e-5 and abs(m2 - tm[k, 2]) < 2e-5 * max(1.0, tm[k, 2])


def test_vsq_autocorr_vs_numpy():
    """mdqt_vsq_autocorr against Zfunc() of the Quad program (FZ408Q:942-967): sum_j (1/N)(Vh_j^2 - a)(V_j^2 - a), a = <V_x^2> now."""
    n = 777
    rng = np.random.default_rng(11)
    V = rng.normal(size=(3, n)) * 0.3
    e = Engine(md_params(scheme=SCHEME_NONE, n_ions=n))
    e.upload(R=rng.uniform(0, e.params.L, size=(3, n)), V=V)
    a = (V[0] ** 2).mean()
    assert np.isclose(e.ZfuncLongKin(0), ((V[0] ** 2 - a) ** 2).mean(), rtol=1e-13)
    e.scale_velocities(1.3, 1.0, 1.0)
    W = V[0] * 1.3
    a = (W ** 2).mean()
    assert np.isclose(e.ZfuncLongKin(1), ((V[0] ** 2 - a) * (W ** 2 - a)).mean(), rtol=1e-13)
    assert np.isclose(e.Zfunc(1), (V[0] * W).mean(), rtol=1e-13)   # the stored velocities serve both correlators
    e.close()


@pytest.mark.parametrize("program", ["fz408q", "fz422l"])
def test_fz_sibling_programs_match_the_python_loop(tmp_path, program):
    """`mdqt_run --program fz408q | fz422l` (randomFrozenStartTag408Quad.cpp / ...422Linear.cpp: the FZ408L main() with the circular-pump
    mask + the v_x^2 correlator, resp. the 5-level 422 nm scheme with its constants) against drivers.fz_main_loop on an engine set
    up by hand from the reference's numbers (FZ408Q:58-60, 441, 969-979; FZ422L:55-74, 116-117, 1000-1005)."""
    save = str(tmp_path) + "/"
    quad = program == "fz408q"
    det, Om = (0.0, 2.0) if quad else (-1.0, 1.3)
    r = subprocess.run([DRIVER, "--program", program, "2", "--N0", "700", "--seed", "99", "--tstartV0", "0.0061", "--tpumpreal", "2e-9",
                        "--tmax", "0.0201", "--sampleFreq", "5", "--saveDirectory", save, "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    sub = os.listdir(save)
    assert len(sub) == 1 and sub[0] == "PumpTime2PumpStart0Det%dOm%dDensity20Ge100NumIons700" % (abs(det) * 100, Om * 100), sub
    d = os.path.join(save, sub[0], "job2")
    files = os.listdir(d)
    c0 = int([f for f in files if f.startswith("ions_")][0][len("ions_timestep"):-4])
    st = hostio.init_su(99, N0=700)
    n = st["N"]
    density = 2.0
    p = su_params(Ge=0.1, density=density, detuning=det, detuningDP=0.0, Om=Om, OmDP=0.0, N0=700, n_ions=n, scheme=SCHEME_SR7 if quad else SCHEME_CA5,
                  quad=1 if quad else 0, traj0=2, seed=99)
    p.substeps_per_md = int(round(34.81 / np.sqrt(density)))
    if not quad:
        p.g2E = 174.07 * .894 / np.sqrt(density)
        p.substeps_per_md = int(round(34.81 * .894 / np.sqrt(density)))
        p.pv2qv = 1.1821 * density ** (1. / 6) * .967
        p.dR = 0.0754
        p.vKick = 0.001257 / p.pv2qv
    p.dtq = 0.002 / p.substeps_per_md
    e = Engine(p)
    S = 7 if quad else 5
    psi = np.zeros((n, S, 2))
    psi[:, :2] = st["psi"][:, :2]
    e.upload(R=st["R"], V=st["V"], psi=psi, t=0.0, substep=0)
    tend = 0.0061 + 2e-9 * 813490 * np.sqrt(density)
    events = []
    out = drivers.fz_main_loop(e, 0.0201, 0.0061, tend, sampleFreq=5, on_measure=lambda t, tg, nu, v: events.append((t, v, 1)),
                               on_sample=lambda t, c, v: events.append((t, v, 0)), zfunc=e.ZfuncLongKin if quad else None)
    assert out["c0"] == c0 and 0 < out["n_up"] < n
    spin = np.loadtxt(os.path.join(d, "spinUpIonsList_timestep%06d.dat" % c0), dtype=int)
    assert np.array_equal(spin, out["tagged"])
    acf = "vSquareAutoCorr.dat" if quad else "VAF.dat"
    assert (("VAF.dat" in files) != quad) and acf in files
    ac = _table(open(os.path.join(d, acf)).read())
    assert ac.shape == (len(events), 2)
    assert np.allclose(ac, np.array(events)[:, :2], rtol=6e-6, atol=1e-300)
    s = e.download()
    cond = _table(open(os.path.join(d, "conditions_timestep%06d.dat" % c0)).read())
    assert np.allclose(cond[:, :3], s["R"].T, rtol=6e-6) and np.allclose(cond[:, 3:6], s["V"].T, rtol=6e-6, atol=1e-12)
    en = _table(open(os.path.join(d, "energies.dat")).read())
    # the 422 nm program calls no output() at the measurement (FZ422L:1000-1005): one row less than correlator rows
    assert en.shape == (len(events) - (0 if quad else 1), 6)
    e.close()


def test_three_state_program_follows_the_oracle_cooling_curve(tmp_path):
    """`mdqt_run --program ts` (laserCoolNoPlasmaThreeState.cpp main(), TS:352-409): energies.dat rows "t, <v_x^2>/2" at the
    reference's output times, and the Doppler cooling the program exists to show -- 1000 free ions at 10 mK (sigma_v = 0.105) in a
    J = 0 -> J = 1 molasses at detuning -0.5, Om 0.5. The curve is held against the oracle restatement of TS qstep() (itself pinned to
    the reference by tests/golden/ts_three_state.npz) run with 300 ions from the same Maxwellian: <v_x^2>/2 = 5.5e-3 at t = 0,
    3.55e-3 at 500, 2.62e-3 at 1000, 1.46e-3 at 2000, 8.8e-4 at 3000 (statistical error of either run ~ sqrt(2/N) = 5-8 %)."""
    save = str(tmp_path) + "/"
    r = subprocess.run([DRIVER, "--program", "ts", "4", "--N0", "1000", "--seed", "5", "--tmax", "3000", "--sampleFreq", "500",
                        "--saveDirectory", save, "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    top = os.path.join(save, "Om50")
    runs = os.listdir(top)
    assert len(runs) == 1 and runs[0].endswith("NumIons1000InitialTemp10000uK"), runs   # TS:375: Det%dNumIons%dInitialTemp%duK
    en = _table(open(os.path.join(top, runs[0], "job4", "energies.dat")).read())
    assert en.shape[0] in (600, 601) and en.shape[1] == 2 and np.allclose(en[:3, 0], [5.0, 10.0, 15.0], rtol=1e-6)
    assert 4.6e-3 < en[0, 1] < 6.4e-3                       # 499 sweeps after the Maxwellian start: (1.0508^2 x 0.01)/2 = 5.52e-3 (TS:83)
    curve = {500.0: 3.55e-3, 1000.0: 2.62e-3, 2000.0: 1.46e-3, 3000.0: 8.8e-4}
    for t, want in curve.items():
        k = int(np.argmin(np.abs(en[:, 0] - t)))
        got = en[max(0, k - 5):k + 1, 1].mean()            # the 6 rows up to t: +-12.5 time units, the curve moves < 1 % over them
        assert abs(got - want) < 0.22 * want, (t, got, want)
    assert np.all(np.diff(en[::100, 1]) < 0)                # monotone cooling on the scale of the run

"""GPU suite, part 2: the mdqt_run driver (the reference's main loop on top of the C ABI) and the output()/restart
files, against files written by the unmodified reference (tests/golden/su_mainloop, and oracle/_ref live when the
prebuilt harness library is present)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from mdqtplasmasims_b200 import Engine, hostio, su_params
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run")


def _numbers(text):
    return np.array([[float(x) for x in line.split()] for line in text.strip().splitlines()])


def _compare_text(ours, gold, rtol, atol, min_identical):
    lo, lg = ours.strip().splitlines(), gold.strip().splitlines()
    assert len(lo) == len(lg)
    same = sum(a == b for a, b in zip(lo, lg))
    assert same >= min_identical * len(lg), "only %d of %d lines byte-identical" % (same, len(lg))
    a, b = _numbers(ours), _numbers(gold)
    assert a.shape == b.shape
    assert np.allclose(a, b, rtol=rtol, atol=atol), np.abs(a - b).max()


def test_mdqt_run_reproduces_the_reference_main_loop(tmp_path, golden_dir):
    """newRun=1, srand48(777), lasers off, tmax=0.081: 41 MD steps, one output() at c0=39, writeConditions at the end.
    Every file the reference wrote is reproduced; text differs at most in the last printed digit of a few numbers."""
    gdir = os.path.join(golden_dir, "su_mainloop")
    save = str(tmp_path) + "/"
    r = subprocess.run([DRIVER, "1", "--seed", "777", "--Om", "0", "--OmDP", "0", "--tmax", "0.081", "--saveDirectory", save],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = hostio.dirname(save, Om=0.0, OmDP=0.0, job=1)
    assert os.path.isdir(d)
    gold = {f[:-3]: gzip.open(os.path.join(gdir, f), "rt").read() for f in os.listdir(gdir) if f.endswith(".gz")}
    c0 = [f for f in gold if f.startswith("ions_")][0][len("ions_timestep"):-4]
    for f, g in gold.items():
        path = os.path.join(d, f)
        assert os.path.exists(path), f
        ours = open(path).read()
        if f.startswith(("ions_", "wvFns_", "VZERO_")):
            assert ours == g, f                                   # integers / untouched S-state amplitudes / zeros: exact
        elif f.startswith("conditions_"):
            _compare_text(ours, g, rtol=2e-6, atol=1e-9, min_identical=0.999)
        elif f == "energies.dat":
            a, b = _numbers(ours), _numbers(g)
            assert a.shape == b.shape == (1, 7)
            assert np.allclose(a[0, :5], b[0, :5], rtol=2e-6) and abs(a[0, 5] - b[0, 5]) < 1e-9 and abs(a[0, 6] - b[0, 6]) < 1e-9
        else:  # vel_dist*, statePopulations*
            _compare_text(ours, g, rtol=2e-5, atol=1e-12, min_identical=0.99)
    # all 13 VZERO files exist (readConditions needs them, SU:898-913)
    assert all(os.path.exists(os.path.join(d, "VZERO_timestep%s_interval%d.dat" % (c0, k))) for k in range(13))


def test_mdqt_run_restart_continues(tmp_path):
    save = str(tmp_path) + "/"
    common = ["--seed", "5", "--N0", "600", "--saveDirectory", save, "--quiet"]
    r = subprocess.run([DRIVER, "3", "--tmax", "0.02"] + common, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = hostio.dirname(save, N0=600, job=3)
    labels = sorted(int(f[len("ions_timestep"):-4]) for f in os.listdir(d) if f.startswith("ions_"))
    assert len(labels) == 1
    first = hostio.read_conditions(d, labels[0], ld=1600)
    # resume from the label (newRun=0, c0=label), as README.md:51-53 describes
    r = subprocess.run([DRIVER, "3", "--newRun", "0", "--c0", str(labels[0]), "--tmax", "0.06"] + common, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    labels2 = sorted(int(f[len("ions_timestep"):-4]) for f in os.listdir(d) if f.startswith("ions_"))
    assert len(labels2) == 2 and labels2[-1] > labels[0]
    second = hostio.read_conditions(d, labels2[-1], ld=1600)
    assert second["N"] == first["N"]
    assert np.all((second["R"] >= 0) & (second["R"] <= (600 * 4 * np.pi / 3) ** 0.333333333))
    norm = (second["psi"] ** 2).sum(axis=(1, 2))
    # the reference's propagator lets the norm drift between jumps (reNormalizewvFns=false, SU:74; ~2e-4 per 200
    # substeps) and the restart files keep 6 significant digits
    assert np.abs(norm - 1).max() < 2e-2
    assert not np.array_equal(second["R"], first["R"])


@pytest.mark.skipif(not po.ref_available("su"), reason="oracle/_ref/libref_su.so not present")
def test_output_files_match_reference_output(tmp_path):
    """output() (SU:917-1032): energies.dat, vel_dist{X,Y,Z}, statePopulationsVsV written from the GPU observables vs
    the reference's own output() on the same state."""
    ref = po.RefSU()
    n = ref.init(4242)
    s = ref.get_state()
    rng = np.random.default_rng(2)
    V = rng.normal(size=(3, n)) * 0.03
    V[0] += 0.004
    psi = rng.normal(size=(n, 12, 2))
    psi /= np.sqrt((psi ** 2).sum(axis=(1, 2)))[:, None, None]
    ref.set_state(R=s["R"], V=V, psi=psi, tPart=np.zeros(n), t=0.75)
    da, db = str(tmp_path / "a") + "/", str(tmp_path / "b") + "/"
    os.mkdir(da); os.mkdir(db)
    ref.set_savedir(da)
    ref.set_counters(39, 7)
    ref.lib.ref_su_set_Epot0(4.5)
    ref.output()
    eng = Engine(su_params(n_ions=n))
    eng.upload(R=s["R"], V=V, psi=psi, tPart=np.zeros(n), t=0.75)
    dg = eng.diagnostics()
    hostio.append_energies(db, dg["t"], dg["ekin_x"], dg["ekin_y"], dg["ekin_z"], dg["epot"], 4.5, dg["vx_avg"])
    hostio.write_vel_dist(db, 7, eng.vel_dist(), dg["vx_avg"])
    hostio.write_populations(db, 7, V[0], eng.populations())
    for f in ("energies.dat", "vel_distX_time000007.dat", "vel_distY_time000007.dat", "vel_distZ_time000007.dat",
              "statePopulationsVsVTime000007.dat"):
        _compare_text(open(db + f).read(), open(da + f).read(), rtol=2e-6, atol=1e-300, min_identical=0.995)


DROPIN = os.path.join(ROOT, "oracle", "_ref", "su_dropin")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/su_dropin not built (needs /root/reference at build time)")
def test_reference_main_runs_on_the_abi(tmp_path, golden_dir):
    """The drop-in boundary executed: the UNMODIFIED reference program's own main() (directory setup, operator tables, init(),
    the time loop, output(), writeConditions()) with forces()/step()/qstep()/Epotential() routed through the C ABI exactly as
    INTEGRATION.md section 2 shows (examples/su_dropin.cpp). srand48(777), lasers off, tmax = 0.081: the files it writes are
    the files the all-CPU reference wrote (tests/golden/su_mainloop)."""
    gdir = os.path.join(golden_dir, "su_mainloop")
    r = subprocess.run([DROPIN, "1", "--seed", "777", "--Om", "0", "--OmDP", "0", "--tmax", "0.081"], capture_output=True, text=True,
                       timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    d = hostio.dirname(str(tmp_path) + "/dataLaserCool/", Om=0.0, OmDP=0.0, job=1)
    assert os.path.isdir(d), os.listdir(str(tmp_path))
    gold = {f[:-3]: gzip.open(os.path.join(gdir, f), "rt").read() for f in os.listdir(gdir) if f.endswith(".gz")}
    for f, g in gold.items():
        path = os.path.join(d, f)
        assert os.path.exists(path), f
        ours = open(path).read()
        if f.startswith(("ions_", "wvFns_", "VZERO_")):
            assert ours == g, f
        elif f.startswith("conditions_"):
            _compare_text(ours, g, rtol=2e-6, atol=1e-9, min_identical=0.999)
        elif f == "energies.dat":
            a, b = _numbers(ours), _numbers(g)
            assert a.shape == b.shape == (1, 7)
            assert np.allclose(a[0, :5], b[0, :5], rtol=2e-6) and abs(a[0, 5] - b[0, 5]) < 1e-9 and abs(a[0, 6] - b[0, 6]) < 1e-9
        else:  # vel_dist* and statePopulations*: computed by the reference's own output() from the downloaded state
            _compare_text(ours, g, rtol=2e-5, atol=1e-12, min_identical=0.99)

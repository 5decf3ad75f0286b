"""GPU suite: ensembles whose jobs drew different ion numbers (every reference job draws N ~ Binomial around N0,
SU:299-337) batched in one handle, and the mdqt_run array driver that replaces the reference's SLURM array
(exampleSlurmFile.slurm:3,16). The bar: a job gives the SAME BITS alone and inside any batch."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from mdqtplasmasims_b200 import Engine, hostio, su_params, synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "mdqtplasmasims_b200", "mdqt_run")


def _job_state(n, L, seed):
    return (synthetic.random_positions(n, L, seed=seed), synthetic.maxwellian(n, 0.05, seed=seed + 1),
            synthetic.random_full_state(n, 12, seed=seed + 2), np.zeros(n))


@pytest.mark.parametrize("N0,counts,nominal", [(600, (571, 640, 600, 613, 588), False), (3500, (3487, 3561, 3440), False),
                                              (3500, (3487, 3653, 3440, 3529), False), (600, (571, 640, 600, 613, 588), True),
                                              (3500, (3487, 3561, 3440), True)])
def test_variable_n_batch_equals_single_trajectories_bitwise(N0, counts, nominal):
    """B jobs with different N in one handle (mdqt_set_ion_counts + per-job Philox keys) against B single-trajectory
    handles: forces, potential energy, observables and the state after 3 MD steps (quantum jumps included) are bitwise
    identical. plan_n = 0 (what mdqt_run uses): every trajectory's summation order follows from its own ion count, so the
    lone job is simply a handle of its own size; plan_n = N0 (`nominal`): one order for all, fixed by the nominal N0."""
    B, cap, traj0 = len(counts), max(counts), 11
    plan_n = N0 if nominal else 0
    seeds = np.array([1000 + 7 * b for b in range(B)], dtype=np.uint64)
    pb = su_params(n_ions=cap, N0=N0, n_traj=B, traj0=traj0, plan_n=plan_n, seed=int(seeds[0]))
    L = pb.L
    R, V, psi, tp = (np.zeros((B, 3, cap)), np.zeros((B, 3, cap)), np.zeros((B, cap, 12, 2)), np.zeros((B, cap)))
    psi[:, :, 0, 0] = 1.0
    jobs = []
    for b, n in enumerate(counts):
        s = _job_state(n, L, 50 + b)
        jobs.append(s)
        R[b, :, :n], V[b, :, :n], psi[b, :n], tp[b, :n] = s
    eb = Engine(pb)
    eb.set_ion_counts(counts)
    eb.set_traj_seeds(seeds)
    eb.upload(R=R, V=V, psi=psi, tPart=tp, t=0.0, substep=0)
    eb.forces()
    Fb = eb.download_forces()
    Eb = eb.Epotential()
    db = eb.diagnostics()
    eb.md_steps(3)
    sb = eb.download()
    pops_b = eb.populations()
    pv_b = eb.vel_dist()
    jumped = 0
    for b, n in enumerate(counts):
        e1 = Engine(su_params(n_ions=n, N0=N0, traj0=traj0 + b, plan_n=plan_n, seed=int(seeds[b])))
        e1.upload(R=jobs[b][0], V=jobs[b][1], psi=jobs[b][2], tPart=jobs[b][3], t=0.0, substep=0)
        e1.forces()
        assert np.array_equal(e1.download_forces(), Fb[b][:, :n])
        assert e1.Epotential() == Eb[b]
        d1 = e1.diagnostics()
        assert all(d1[k] == db[b][k] for k in d1)
        e1.md_steps(3)
        s1 = e1.download()
        for k in ("R", "V"):
            assert np.array_equal(s1[k], sb[k][b][:, :n]), k
        assert np.array_equal(s1["psi"], sb["psi"][b][:n])
        assert np.array_equal(s1["tPart"], sb["tPart"][b][:n])
        assert np.array_equal(e1.populations(), pops_b[b][:n])
        assert np.array_equal(e1.vel_dist(), pv_b[b])
        jumped += int((s1["tPart"] < 75 * pb.dtq * 0.999).sum())
        e1.close()
    assert jumped > 0  # the comparison covered quantum jumps
    eb.close()


def test_items_kernel_matches_tile_kernel_and_plan_is_batch_independent():
    """The item-walking force kernel against the CTA-tile kernel (MDQT_K1_ITEMS=0) on the same positions: the two sum in
    different orders, so agreement is to rounding (1e-13 relative to the largest force); and the plan depends on plan_n only."""
    n = 3000
    p = su_params(n_ions=n, N0=n)
    R = synthetic.random_positions(n, p.L, seed=5)
    a = Engine(p)
    a.upload(R=R)
    a.forces()
    Fa = a.download_forces()
    os.environ["MDQT_K1_ITEMS"] = "0"
    try:
        b = Engine(su_params(n_ions=n, N0=n))
    finally:
        del os.environ["MDQT_K1_ITEMS"]
    b.upload(R=R)
    b.forces()
    Fb = b.download_forces()
    assert np.abs(Fa - Fb).max() <= 1e-13 * np.abs(Fb).max()
    assert abs(a.Epotential() - b.Epotential()) <= 1e-13 * abs(b.Epotential())
    plans = set()
    for B in (1, 3, 16):
        e = Engine(su_params(n_ions=n, N0=n, n_traj=B, plan_n=3500))
        plans.add(e.force_plan()[1])
        e.close()
    assert len(plans) == 1


def test_batch_rejects_md_family_calls_with_ion_counts():
    e = Engine(su_params(n_ions=256, N0=256, n_traj=2))
    e.set_ion_counts([250, 256])
    with pytest.raises(Exception):
        e.MDStep()
    with pytest.raises(Exception):
        e.set_ion_counts([0, 256])
    with pytest.raises(Exception):
        e.set_ion_counts([257, 256])
    e.set_ion_counts(None)
    e.close()


def test_mdqt_run_array_reproduces_single_jobs(tmp_path):
    """mdqt_run --jobs 3-10 --batch 3 (three batches of 3, 3, 2 jobs with different N, lasers on) writes, for every job,
    byte-identical files to eight single-job runs `mdqt_run j --seed S+j`."""
    sa, sb = str(tmp_path / "array") + "/", str(tmp_path / "single") + "/"
    os.mkdir(sa); os.mkdir(sb)
    common = ["--N0", "400", "--tmax", "0.17", "--quiet"]
    r = subprocess.run([DRIVER, "--jobs", "3-10", "--batch", "3", "--seed", "100", "--saveDirectory", sa] + common, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ns = set()
    for j in range(3, 11):
        r = subprocess.run([DRIVER, str(j), "--seed", str(100 + j), "--saveDirectory", sb] + common, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr
        da, db = hostio.dirname(sa, N0=400, job=j), hostio.dirname(sb, N0=400, job=j)
        fa, fb = sorted(os.listdir(da)), sorted(os.listdir(db))
        assert fa == fb and len(fa) >= 1 + 2 * 4 + 16
        match, mismatch, err = filecmp.cmpfiles(da, db, fa, shallow=False)
        assert not mismatch and not err, (j, mismatch[:3], err[:3])
        ns.add(int(open(os.path.join(da, [f for f in fa if f.startswith("ions_")][0])).read().split()[0]))
        pop = np.loadtxt(os.path.join(da, "statePopulationsVsVTime000001.dat"))
        assert pop[:, 2].mean() > 0.05  # lasers on: P population present, so jumps happened
    assert len(ns) > 1  # the jobs really had different ion numbers


def test_ion_counts_can_be_changed_and_reset():
    """mdqt_set_ion_counts re-plans the item kernel (per-trajectory chunk lengths, partial-sum buffers): setting counts, changing
    them and clearing them again leaves a handle that gives the bits of a fresh one."""
    n, B = 1400, 3
    p = su_params(n_ions=n, N0=n, n_traj=B)
    R = np.stack([synthetic.random_positions(n, p.L, seed=70 + b) for b in range(B)])
    fresh = Engine(p)
    fresh.upload(R=R)
    fresh.forces()
    F0, E0 = fresh.download_forces(), fresh.Epotential()
    e = Engine(p)
    e.upload(R=R)
    e.set_ion_counts((700, 1400, 333))   # small trajectories need MORE chunk slots than the uniform plan holds
    e.forces()
    F1 = e.download_forces()
    for b, nb in enumerate((700, 1400, 333)):
        one = Engine(su_params(n_ions=nb, N0=n))
        one.upload(R=np.ascontiguousarray(R[b][:, :nb]))
        one.forces()
        assert np.array_equal(one.download_forces(), F1[b][:, :nb]), b
        one.close()
    e.set_ion_counts((1399, 5, 1400))
    e.forces()
    assert np.all(np.isfinite(e.download_forces()))
    e.set_ion_counts(None)
    e.forces()
    assert np.array_equal(e.download_forces(), F0) and np.array_equal(e.Epotential(), E0)
    e.close(); fresh.close()

"""GPU suite, multi-GPU part (needs >= 2 devices; skipped on a single-GPU box): the row-decomposed large-N path INSIDE the
library -- mdqt_comm_init + mdqt_md_steps with one in-place ncclAllGather of fixed-point positions per MD step, overlapped with
the local j chunks of the next force call -- against the same run on one GPU. One Python thread per GPU (ctypes releases the
GIL, so the collective calls run concurrently), one handle and one communicator each."""
import filecmp
import os
import subprocess
import threading

import numpy as np
import pytest

from mdqtplasmasims_b200 import Engine, load_library, su_params, synthetic

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        return load_library().mdqt_device_count()
    except Exception:
        return 0


def _run_ranks(world, n, state, nsteps, seed):
    uid = Engine.comm_unique_id()
    out, err = [None] * world, [None] * world

    def work(rank):
        try:
            rows = n // world
            e = Engine(su_params(n_ions=n, N0=n, row0=rank * rows, n_rows=rows, device=rank, seed=seed))
            e.comm_init(uid, rank, world)
            e.upload(R=state[0], V=state[1], psi=state[2], tPart=state[3], t=0.0, substep=0)
            e.md_steps(1)
            e.md_steps(nsteps - 1)
            d = e.diagnostics()
            pv = e.vel_dist()
            s = e.download()
            out[rank] = (s, e.download_forces(), d, pv, e.populations_rows(), e.force_plan())
            e.close()
        except Exception as ex:  # pragma: no cover
            err[rank] = ex

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not any(err), err
    return out


@pytest.mark.skipif(_ndev() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("n", [2048, 16384])  # item-walking force kernel (no chunk split) and CTA-tile kernel (overlapped exchange)
def test_comm_md_steps_equal_the_single_gpu_run_bitwise(n):
    nsteps, seed = 4, 31
    p = su_params(n_ions=n, N0=n, seed=seed)
    state = (synthetic.random_positions(n, p.L, seed=1), synthetic.maxwellian(n, 0.05, seed=2), synthetic.random_full_state(n, 12, seed=3),
             np.zeros(n))
    full = Engine(p)
    full.upload(R=state[0], V=state[1], psi=state[2], tPart=state[3], t=0.0, substep=0)
    full.md_steps(1)
    full.md_steps(nsteps - 1)
    s0, F0, d0, pv0, pop0 = full.download(), full.download_forces(), full.diagnostics(), full.vel_dist(), full.populations()
    assert (s0["tPart"] < nsteps * 25 * p.dtq * 0.999).sum() > 0  # jumps happened
    for world in [w for w in (2, 4, 8) if w <= _ndev()]:
        res = _run_ranks(world, n, state, nsteps, seed)
        rows = n // world
        for rank, (s, F, d, pv, pops, plan) in enumerate(res):
            sl = slice(rank * rows, (rank + 1) * rows)
            assert plan == full.force_plan()
            assert np.array_equal(F[:, sl], F0[:, sl])
            for k in ("R", "V"):
                assert np.array_equal(s[k][:, sl], s0[k][:, sl]), (world, rank, k)
            assert np.array_equal(s["psi"][sl], s0["psi"][sl]) and np.array_equal(s["tPart"][sl], s0["tPart"][sl])
            assert s["t"] == s0["t"]
            assert np.array_equal(pops, pop0[sl])
            # observables: partial sums in a different order than the one-GPU reductions -> rounding-level agreement, and the
            # same value on every rank
            for k in ("ekin_x", "ekin_y", "ekin_z", "epot"):
                assert abs(d[k] - d0[k]) <= 1e-12 * abs(d0[k]), (k, d[k], d0[k])
                assert d[k] == res[0][2][k]
            assert abs(d["vx_avg"] - d0["vx_avg"]) <= 1e-15
            assert np.abs(pv - pv0).max() <= 1e-11 * pv0.max() and np.array_equal(pv, res[0][3])
    full.close()


@pytest.mark.skipif(_ndev() < 2, reason="needs at least 2 GPUs")
def test_mdqt_run_row_decomposed_writes_the_single_gpu_files(tmp_path):
    """`mdqt_run <job> --gpus G`: one job row-decomposed over G GPUs inside the library (unequal row blocks when G does not divide
    N) writes byte-identical files to the one-GPU run of the same job, lasers on."""
    from mdqtplasmasims_b200 import hostio
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    drv = os.path.join(root, "mdqtplasmasims_b200", "mdqt_run")
    common = ["4", "--seed", "21", "--N0", "1200", "--tmax", "0.17", "--quiet"]
    s1 = str(tmp_path / "one") + "/"
    os.mkdir(s1)
    r = subprocess.run([drv] + common + ["--saveDirectory", s1], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    d1 = hostio.dirname(s1, N0=1200, job=4)
    files = sorted(os.listdir(d1))
    for G in [g for g in (2, 3, 8) if g <= _ndev()]:
        sg = str(tmp_path / ("g%d" % G)) + "/"
        os.mkdir(sg)
        r = subprocess.run([drv] + common + ["--gpus", str(G), "--saveDirectory", sg], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        dg = hostio.dirname(sg, N0=1200, job=4)
        assert sorted(os.listdir(dg)) == files
        # per-ion files (restart files, populations): byte-identical. energies.dat and the velocity distributions are sums over the
        # ranks' partial sums (a different order than the one-GPU reduction): equal to the printed precision up to its last digit
        exact = [f for f in files if not (f == "energies.dat" or f.startswith("vel_dist"))]
        match, mismatch, err = filecmp.cmpfiles(d1, dg, exact, shallow=False)
        assert not mismatch and not err, (G, mismatch[:4], err[:4])
        for f in files:
            if f not in exact:
                a, b = np.loadtxt(os.path.join(d1, f), ndmin=2), np.loadtxt(os.path.join(dg, f), ndmin=2)
                assert a.shape == b.shape and np.allclose(a, b, rtol=3e-6, atol=1e-300), f

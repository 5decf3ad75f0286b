"""CPU suite, part 2: the oracle restatement against the LIVE unmodified reference (oracle/_ref/libref_*.so: the
reference sources compiled through the hijack harnesses). Needs oracle/_ref (built in the container where
/root/reference exists; the prebuilt .so files travel with the snapshot). Skipped when absent."""
import numpy as np
import pytest

from oracle import pyoracle as po

needs_su = pytest.mark.skipif(not po.ref_available("su"), reason="oracle/_ref/libref_su.so not built")
needs_md = pytest.mark.skipif(not po.ref_available("md"), reason="oracle/_ref/libref_md.so not built")
needs_mc = pytest.mark.skipif(not po.ref_available("mc408l"), reason="oracle/_ref/libref_mc408l.so not built")


def full_psi(rng, n, S):
    psi = rng.normal(size=(n, S, 2))
    psi /= np.sqrt((psi ** 2).sum(axis=(1, 2)))[:, None, None]
    return psi


@needs_su
def test_operator_tables_reduce_to_the_sparse_form():
    """decayMatrix is diagonal (Gamma = 1 + dR on the P states), hamDecayTerm = -i/2 decayMatrix, and the static
    coupling matrix has exactly the 10 real entries the sparse restatement uses."""
    ref = po.RefSU()
    C, D, HD, gs = ref.tables()
    assert np.count_nonzero(D - np.diag(np.diag(D))) == 0
    assert np.allclose(np.diag(D).real, [0, 0, 1.0617, 1.0617, 1.0617, 1.0617] + [0] * 6, atol=1e-15)
    assert np.allclose(HD, -0.5j * D, atol=0)
    nz = {(r, c): C[r, c] for r in range(12) for c in range(12) if C[r, c] != 0}
    assert set(nz) == {(2, 1), (3, 0), (4, 1), (5, 0), (5, 6), (4, 7), (3, 8), (3, 10), (2, 9), (2, 11)}
    assert all(abs(v.imag) == 0 for v in nz.values())
    assert abs(nz[(2, 1)] + 0.5) < 1e-16 and abs(nz[(3, 0)] + 0.5 / np.sqrt(3)) < 1e-16
    assert abs(nz[(5, 6)] + 0.5 * np.sqrt(2. / 3)) < 1e-15 and abs(nz[(2, 9)] + 0.5 * np.sqrt(1. / 15)) < 1e-15


@needs_su
def test_forces_epot_step_bitwise(oracle):
    ref = po.RefSU()
    c = ref.consts
    rng = np.random.default_rng(0)
    n = 700
    R = rng.uniform(0, c["L"], size=(3, n)); V = rng.normal(size=(3, n)) * 0.2
    ref.set_state(R=R, V=V, psi=full_psi(rng, n, 12), tPart=np.zeros(n), t=0.0)
    ref.forces()
    F = ref.get_state()["F"]
    assert np.array_equal(oracle.forces_su(R, c["L"], c["lDeb"]), F)
    assert oracle.epot_su(R, c["L"], c["lDeb"]) == ref.epot()
    for t in (0.0, 0.5):  # 2nd-order start branch (SU:370-379) and plain leap-frog
        ref.set_state(R=R, V=V, t=t)
        ref.step()
        s = ref.get_state()
        R2, V2 = R.copy(), V.copy()
        oracle.step_su(R2, V2, F, c["L"], c["dtq"], t)
        assert np.array_equal(R2, s["R"]) and np.array_equal(V2, s["V"])


@needs_su
@pytest.mark.parametrize("frac,dens", [(0.0, 2.0), (0.5, 2.0), (0.3, 0.7)])
def test_qstep12_with_jumps_per_ion_streams(oracle, frac, dens):
    """30 sweeps with per-ion uniform streams, half the ions forced to jump now and then: identical branch
    decisions (draws consumed) and amplitudes to ~1e-15."""
    ref = po.RefSU(fracOfSig=frac, density=dens)
    p, ratio = po.su_params(fracOfSig=frac, density=dens)
    assert ratio == int(ref.consts["ratio"]) and p.dtq == ref.consts["dtq"] and p.vKick == ref.consts["vKick"]
    rng = np.random.default_rng(1)
    n = 120
    psi = full_psi(rng, n, 12); V = rng.normal(size=(3, n)) * 0.3; tp = rng.uniform(0, 0.5, size=n)
    ref.set_state(R=np.zeros((3, n)), V=V, psi=psi, tPart=tp, t=0.37)
    psi_o, vx, tpo, t = psi.copy(), V[0].copy(), tp.copy(), 0.37
    njump = 0
    for step in range(30):
        u5 = rng.uniform(size=(n, 5))
        u5[:, 0] = np.where(rng.uniform(size=n) < 0.15, 1e-12, 0.9999)
        used_r = ref.qstep_stream(u5)
        t, used_o = oracle.qstep12(psi_o, vx, tpo, t, p, u5)
        assert np.array_equal(used_r, used_o)
        njump += int((used_o > 1).sum())
        s = ref.get_state()
        assert np.abs(s["psi"] - psi_o).max() <= 5e-15
        assert np.abs(s["V"][0] - vx).max() <= 1e-17
        assert np.array_equal(s["tPart"], tpo) and s["t"] == t
    assert njump > 100


@needs_su
def test_qstep12_sequential_stream_and_renormalise(oracle):
    ref = po.RefSU(renorm=1)
    p, _ = po.su_params(renorm=1)
    rng = np.random.default_rng(2)
    n = 64
    psi = full_psi(rng, n, 12) * 1.01; V = rng.normal(size=(3, n)) * 0.1; tp = np.zeros(n)
    ref.set_state(R=np.zeros((3, n)), V=V, psi=psi, tPart=tp, t=0.1)
    u = rng.uniform(size=5 * n)
    used = ref.qstep(u)
    psi_o, vx, tpo = psi.copy(), V[0].copy(), tp.copy()
    _, used_o = oracle.qstep12(psi_o, vx, tpo, 0.1, p, u, sequential=True)
    assert used == used_o.sum()
    s = ref.get_state()
    assert np.abs(s["psi"] - psi_o).max() <= 5e-15
    assert np.abs((psi_o ** 2).sum(axis=(1, 2)) - 1).max() <= 1e-15


@needs_su
def test_main_loop_schedule_with_lasers_off(oracle):
    """Two MD steps of the reference main loop (forces every 25 substeps, SU:1369-1378) with Om = OmDP = 0
    (no P population, hence no jumps): the restatement follows bitwise in R and V."""
    ref = po.RefSU(Om=0.0, OmDP=0.0)
    p, ratio = po.su_params(Om=0.0, OmDP=0.0)
    c = ref.consts
    rng = np.random.default_rng(3)
    n = 300
    L = (n * 4 * np.pi / 3) ** (1. / 3)
    ref.set_box(L, c["lDeb"])
    R = rng.uniform(0, L, size=(3, n)); V = np.zeros((3, n)); psi = np.zeros((n, 12, 2)); psi[:, 0, 0] = 1.0
    ref.set_state(R=R, V=V, psi=psi, tPart=np.zeros(n), t=0.0)
    ref.lib.ref_su_set_counters(-1, 0)
    tsc = ref.lib.ref_su_run_loop(50, ratio, 0)
    assert tsc == ratio and ref.lib.ref_su_get_c0() == 1
    s = ref.get_state()
    Ro, Vo, t = R.copy(), V.copy(), 0.0
    tp = np.zeros(n)
    for k in range(50):
        if k % 25 == 0:
            F = oracle.forces_su(Ro, L, c["lDeb"])
        oracle.step_su(Ro, Vo, F, L, p.dtq, t)
        vx = Vo[0].copy()
        t, _ = oracle.qstep12(psi, vx, tp, t, p, np.full((n, 5), 0.5))
        Vo[0] = vx
    assert np.array_equal(Ro, s["R"]) and np.array_equal(Vo, s["V"]) and t == s["t"]


@needs_md
def test_md_family_bitwise(oracle):
    md = po.RefMD()
    c = md.consts
    md.seed(7); md.init(); md.set_controls(0.0, 0, 0)
    rng = np.random.default_rng(4)
    R = rng.uniform(0, c["L"], size=(3, md.N))
    md.set_state(R=R)
    V = md.get_state()["V"].copy()
    md.accelerations()
    A = md.get_state()["A"].copy()
    assert np.array_equal(oracle.forces_md(R, c["L"], c["kappa"], c["rCut"]), A)
    md.set_controls(0.0, 1, 1)  # addLaserForce=1, applyForceAlongOneAxisOnly=true (MD:490-491)
    md.mdstep()
    s = md.get_state()
    R2, V2 = R.copy(), V.copy()
    oracle.vv_positions(R2, V2, A, c["L"], c["timeStep"])
    A2 = oracle.forces_md(R2, c["L"], c["kappa"], c["rCut"])
    oracle.vv_velocities(V2, A, A2, c["timeStep"], laser=2, beta=c["beta"], dens=c["n"])
    assert np.array_equal(R2, s["R"]) and np.array_equal(V2, s["V"]) and np.array_equal(A2, s["A"])


@needs_mc
def test_qstep7_full_sweeps_with_jumps(oracle):
    """MC408L's serial sweep over 4096 ions consuming ONE sequential uniform stream (jumps included)."""
    mc = po.RefMC408L()
    c = mc.consts
    p, ratio = po.mc408_params(n=c["n"])
    assert ratio == int(c["ratio"]) and p.dtq == c["dtq"]
    rng = np.random.default_rng(5)
    n = mc.N
    psi = full_psi(rng, n, 7); V = rng.normal(size=(3, n)) * 0.5
    mc.set_state(V=V, psi=psi)
    psi_o = psi.copy()
    for step in range(4):
        u = rng.uniform(size=5 * n)
        if step % 2 == 0:
            u[::5] = 1e-9
        used = mc.qstep(u)
        used_o = oracle.qstep7(psi_o, V[0].copy(), p, u, sequential=True)
        assert used == used_o.sum()
        assert np.abs(mc.get_state()["psi"] - psi_o).max() <= 2e-15

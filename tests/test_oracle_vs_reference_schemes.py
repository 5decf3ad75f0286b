"""CPU suite, part 3: the oracle restatements of the OTHER level schemes and drivers against the
LIVE unmodified reference programs compiled through the hijack harnesses:
  MC422L  5-level 422 nm pump qstep() + tagParticles()          (oracle/_ref/libref_mc422l.so)
  TS      3-level test system qstep() with kick                 (oracle/_ref/libref_ts.so)
  FZ408L  leap-frog step(), 7-level qstep(), measureSpinUps(), Zfunc()   (oracle/_ref/libref_fz408l.so)
  MC408L  tagParticles()
Skipped when oracle/_ref is absent."""
import numpy as np
import pytest

from oracle import pyoracle as po

needs_m422 = pytest.mark.skipif(not po.ref_available("mc422l"), reason="oracle/_ref/libref_mc422l.so not built")
needs_ts = pytest.mark.skipif(not po.ref_available("ts"), reason="oracle/_ref/libref_ts.so not built")
needs_fz = pytest.mark.skipif(not po.ref_available("fz408l"), reason="oracle/_ref/libref_fz408l.so not built")
needs_mc = pytest.mark.skipif(not po.ref_available("mc408l"), reason="oracle/_ref/libref_mc408l.so not built")


def full_psi(rng, n, S):
    psi = rng.normal(size=(n, S, 2))
    psi /= np.sqrt((psi ** 2).sum(axis=(1, 2)))[:, None, None]
    return psi


@needs_m422
def test_qstep5_full_sweeps_with_jumps(oracle):
    """MC422L's serial sweep over 4096 ions consuming ONE sequential uniform stream (jumps included)."""
    mc = po.RefMC422L()
    c = mc.consts
    p, ratio = po.mc422_params(n=c["n"])
    assert ratio == int(c["ratio"]) and p.dtq == c["dtq"] and p.g2E == c["g2E"] and p.pv2qv == c["pv2qv"] and p.dR == c["dR"]
    rng = np.random.default_rng(7)
    n = mc.N
    psi = full_psi(rng, n, 5)
    V = rng.normal(size=(3, n)) * 0.5
    mc.set_state(V=V, psi=psi)
    psi_o = psi.copy()
    for step in range(4):
        u = rng.uniform(size=5 * n)
        if step % 2 == 0:
            u[::3] = 1e-9  # many forced jumps, at shifting stream positions
        used = mc.qstep(u)
        used_o = oracle.qstep5(psi_o, V[0].copy(), p, u, sequential=True)
        assert used == used_o.sum()
        assert np.abs(mc.get_state()["psi"] - psi_o).max() <= 2e-15
    assert np.array_equal(mc.get_state()["V"], V)  # no kick in the pump stage


@needs_m422
def test_tag422_matches_reference(oracle):
    mc = po.RefMC422L()
    rng = np.random.default_rng(8)
    psi = full_psi(rng, mc.N, 5)
    mc.set_state(psi=psi)
    u = rng.uniform(size=2 * mc.N)
    tagged, used = mc.tag(u)
    tagged_o, used_o = oracle.tag(psi, u, sequential=True)
    assert used == used_o and np.array_equal(tagged, tagged_o)
    assert 0.3 < tagged.mean() < 0.7


@needs_mc
def test_tag408_matches_reference(oracle):
    mc = po.RefMC408L()
    rng = np.random.default_rng(9)
    psi = full_psi(rng, mc.N, 7)
    mc.set_state(psi=psi)
    u = rng.uniform(size=2 * mc.N)
    tagged, used = mc.tag(u)
    tagged_o, used_o = oracle.tag(psi, u, sequential=True)
    assert used == used_o and np.array_equal(tagged, tagged_o)


@needs_ts
def test_qstep3_sweeps_with_jumps_and_kicks(oracle):
    ts = po.RefTS(detuning=-0.5, Om=0.5)
    p = po.ts_params(detuning=-0.5, Om=0.5)
    assert p.vKick == ts.vKick
    rng = np.random.default_rng(10)
    n = ts.N
    psi = full_psi(rng, n, 3)
    Vx = rng.normal(size=n) * 0.1
    tp = rng.uniform(size=n)
    ts.set_state(Vx=Vx, psi=psi, tPart=tp)
    psi_o, Vx_o, tp_o = psi.copy(), Vx.copy(), tp.copy()
    for step in range(6):
        u = rng.uniform(size=2 * n)
        if step % 2 == 1:
            u[::3] = 1e-9
        used = ts.qstep(u)
        used_o = oracle.qstep3(psi_o, Vx_o, tp_o, p, u, sequential=True)
        assert used == used_o.sum()
        s = ts.get_state()
        assert np.abs(s["psi"] - psi_o).max() <= 2e-15
        assert np.abs(s["Vx"] - Vx_o).max() <= 1e-17 and np.array_equal(s["tPart"], tp_o)


@needs_fz
def test_fz_leapfrog_step_bitwise(oracle):
    """FZ408L step(): first call (t = 0: 2nd-order start, forces() three times) and a later call, bitwise."""
    fz = po.RefFZ408L()
    c = fz.consts
    n = fz.init(4242)
    s0 = fz.get_state()
    R, V = s0["R"].copy(), s0["V"].copy()
    dt = c["dtq"] * c["ratio"]
    fz.step()
    s1 = fz.get_state()
    oracle.lf_step(R, V, c["L"], c["lDeb"], dt, first=True)
    assert np.array_equal(R, s1["R"]) and np.array_equal(V, s1["V"])
    fz.set_state(R=R, V=V, t=0.002)
    fz.step()
    s2 = fz.get_state()
    F = oracle.lf_step(R, V, c["L"], c["lDeb"], dt, first=False)
    assert np.array_equal(R, s2["R"]) and np.array_equal(V, s2["V"]) and np.array_equal(F, s2["F"])
    assert n == R.shape[1]


@needs_fz
def test_fz_qstep_measure_and_vaf(oracle):
    fz = po.RefFZ408L()
    c = fz.consts
    n = fz.init(99)
    rng = np.random.default_rng(12)
    psi = full_psi(rng, n, 7)
    V = rng.normal(size=(3, n)) * 0.3
    fz.set_state(V=V, psi=psi, t=15.01, n=n)
    # FZ408L's qstep body is MC408L's with h = (0.002/25) g2E; it also advances t (FZ408L:597)
    import math
    p, _ = po.mc408_params(n=2.0)
    p.dtq = c["dtq"]
    assert p.g2E == c["g2E"] and p.pv2qv == c["pv2qv"] and int(c["ratio"]) == int(round(34.81 / math.sqrt(2.0)))
    psi_o = psi.copy()
    t0 = fz.get_state()["t"]
    for step in range(3):
        u = rng.uniform(size=5 * n)
        if step == 1:
            u[::4] = 1e-9
        used = fz.qstep(u)
        used_o = oracle.qstep7(psi_o, V[0].copy(), p, u, sequential=True)
        assert used == used_o.sum()
        assert np.abs(fz.get_state()["psi"] - psi_o).max() <= 2e-15
    tt = t0
    for _ in range(3):
        tt += c["dtq"]
    assert fz.get_state()["t"] == tt
    u = rng.uniform(size=2 * n)
    tagged, cnt, used = fz.measure(u)
    tagged_o, used_o = oracle.tag(psi_o, u, sequential=True)
    assert used == used_o and cnt == tagged_o.sum() and np.array_equal(tagged, tagged_o)
    v0 = fz.zfunc(0)
    assert v0 == oracle.vaf(V[0], V[0])
    V2 = V * 0.9 + 0.01
    fz.set_state(V=V2, n=n)
    assert fz.zfunc(1) == oracle.vaf(V[0], V2[0])


@pytest.mark.skipif(not po.ref_available("md"), reason="oracle/_ref/libref_md.so not built")
def test_pair_correlation_matches_reference_file(oracle):
    """recordPairPairCorr() (MD:584-652) through its output file: every bin to the file's 6 digits, and the integer
    pair counts recovered from the file equal the restatement's."""
    md = po.RefMD()
    c = md.consts
    rng = np.random.default_rng(77)
    R = rng.uniform(0, c["L"], size=(3, md.N))
    md.set_state(R=R)
    r, g, step, rmax = md.pair_correlation()
    counts, gr = oracle.pair_correlation(R, c["L"], step, rmax)
    assert len(g) == len(gr)
    assert np.allclose(gr, g, rtol=6e-6, atol=0)
    norm = np.where(np.arange(len(g)) == 0, md.N * 4 // 3 * np.pi * step ** 3, md.N * 3 * step ** 3 * np.arange(len(g)) ** 2.0)
    rec = g * norm
    big = counts > 0
    assert np.abs(rec[big] - counts[big]).max() <= 6e-6 * counts.max()

"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates the golden input/output vectors under tests/golden/ by running the UNMODIFIED reference programs
(through oracle/_ref/libref_*.so, i.e. /root/reference compiled by oracle/Makefile) on seeded inputs.
Run in the build container (where /root/reference exists):

    OMP_NUM_THREADS=1 python -m oracle.gen_golden

The reference has no golden vectors of its own; these files pin the oracle restatement and
the CUDA path to the reference's actual outputs and travel to the GPU box, where /root/reference does not exist.
"""
import ctypes
import os
import sys

import numpy as np

os.environ["OMP_NUM_THREADS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NOJUMP = 0.99999  # injected `rand`: always > dp, so no jump


def full_psi(rng, n, S):
    psi = rng.normal(size=(n, S, 2))
    psi /= np.sqrt((psi ** 2).sum(axis=(1, 2)))[:, None, None]
    return psi


def gen_su_forces():
    ref = po.RefSU()
    N = ref.init(12345)  # random frozen start, drand48 seed 12345 (SU:289-348)
    s = ref.get_state()
    ref.forces()
    F = ref.get_state()["F"]
    np.savez(os.path.join(OUT, "su_forces_N%d.npz" % N), R=s["R"], F=F, Epot=ref.epot(), L=ref.consts["L"],
             lDeb=ref.consts["lDeb"], psi0=s["psi"][:64])
    print("su_forces: N=%d F[:,0]=%s Epot=%.17g" % (N, F[:, 0], ref.epot()))


def gen_su_nojump():
    """{step(); qstep();} x nsub with injected no-jump uniforms, F frozen: the deterministic evolution."""
    rng = np.random.default_rng(2024)
    n = 96
    cases = {}
    for name, frac, t0, nsub in (("a", 0.0, 0.0, 1), ("b", 0.0, 0.0, 25), ("c", 0.5, 0.4321, 25), ("d", 0.5, 0.0, 3),
                                 ("e", 0.0, 7.5, 50)):
        ref = po.RefSU(fracOfSig=frac)
        c = ref.consts
        R = rng.uniform(0, c["L"], size=(3, n))
        R[0, :4] = [1e-7, c["L"] - 1e-7, 0.0, c["L"]]  # exercise the wrap (SU:381-389)
        V = rng.normal(size=(3, n)) * 0.3
        V[0, :2] = [-0.5, 0.5]
        F = rng.normal(size=(3, n)) * 2.0
        psi = full_psi(rng, n, 12)
        psi[:8] = 0.0
        psi[:8, 0, 0] = 1.0  # pure S ions
        tp = rng.uniform(0, 0.5, size=n)
        ref.set_state(R=R, V=V, psi=psi, tPart=tp, t=t0)
        ref.set_F(F)
        for _ in range(nsub):
            ref.step()
            ref.qstep(np.full(n, NOJUMP))
        o = ref.get_state()
        for k, v in (("R", R), ("V", V), ("F", F), ("psi", psi), ("tPart", tp), ("t0", t0), ("nsub", nsub), ("frac", frac),
                     ("R_out", o["R"]), ("V_out", o["V"]), ("psi_out", o["psi"]), ("tPart_out", o["tPart"]), ("t_out", o["t"])):
            cases[name + "_" + k] = v
    np.savez(os.path.join(OUT, "su_nojump.npz"), **cases)
    print("su_nojump: cases a-e written")


def gen_su_jumps():
    """Every branch of the jump table (SU:573-703) by forced uniforms, through the reference's own qstep()."""
    ref = po.RefSU()
    rows = []
    # P source state (0-indexed 2..5) x {S decay, D decay} x kick sign x destination selector
    for src in (2, 3, 4, 5):
        for u2 in (0.01, 0.5):          # randDOrS: < dR/(1+dR)=0.0581 -> D decay
            for u3 in (0.25, 0.75):     # randDir
                for u4 in (0.02, 0.3, 0.5, 0.7, 0.97):  # rand3
                    rows.append((src, u2, u3, u4))
    n = len(rows)
    rng = np.random.default_rng(7)
    psi = np.zeros((n, 12, 2))
    u5 = np.zeros((n, 5))
    for i, (src, u2, u3, u4) in enumerate(rows):
        # P populations 0.1 each: cumulative thresholds 0.25/0.5/0.75 select the source sublevel
        amp = rng.normal(size=(12, 2))
        amp[2:6] *= 0.0
        for m in range(2, 6):
            ph = rng.uniform(0, 2 * np.pi)
            amp[m] = np.sqrt(0.1) * np.array([np.cos(ph), np.sin(ph)])
        rest = np.sqrt((amp[:2] ** 2).sum() + (amp[6:] ** 2).sum())
        amp[:2] *= np.sqrt(0.6) / rest
        amp[6:] *= np.sqrt(0.6) / rest
        psi[i] = amp
        u5[i] = [1e-12, (src - 2) * 0.25 + 0.125, u2, u3, u4]
    V = np.zeros((3, n))
    V[0] = rng.normal(size=n) * 0.2
    tp = rng.uniform(0.1, 0.3, size=n)
    ref.set_state(R=np.zeros((3, n)), V=V, psi=psi, tPart=tp, t=1.0)
    used = ref.qstep_stream(u5)
    o = ref.get_state()
    np.savez(os.path.join(OUT, "su_jumps.npz"), psi=psi, V=V, tPart=tp, u5=u5, used=used, psi_out=o["psi"], V_out=o["V"],
             tPart_out=o["tPart"])
    dests = np.argmax((o["psi"] ** 2).sum(axis=2), axis=1)
    print("su_jumps: %d cases, used hist %s, %d distinct (src,dest) pairs" %
          (n, np.bincount(used), len(set(zip([r[0] for r in rows], dests.tolist())))))


def gen_su_stream():
    """A short coupled run with the engine's own Philox uniforms (orc_uniforms5), jumps included: F frozen per MD
    step, forces() recomputed by the reference every 25 substeps (the main-loop schedule SU:1369-1378)."""
    orc = po.Oracle()
    ref = po.RefSU()
    c = ref.consts
    rng = np.random.default_rng(99)
    n, seed, traj, nsub_total = 128, 20260101, 3, 50
    # a small dense box so that forces matter: override L for n ions
    L = (n * 4 * np.pi / 3) ** (1. / 3)
    ref.set_box(L, c["lDeb"])
    R = rng.uniform(0, L, size=(3, n))
    V = rng.normal(size=(3, n)) * 0.1
    psi = full_psi(rng, n, 12)  # substantial P population -> jumps within 50 substeps
    tp = np.zeros(n)
    ref.set_state(R=R, V=V, psi=psi, tPart=tp, t=0.0)
    used_all = []
    for s in range(nsub_total):
        if s % 25 == 0:
            ref.forces()
        ref.step()
        used_all.append(ref.qstep_stream(orc.uniforms5(seed, traj, n, s)))
    o = ref.get_state()
    used_all = np.array(used_all)
    np.savez(os.path.join(OUT, "su_stream.npz"), R=R, V=V, psi=psi, tPart=tp, L=L, lDeb=c["lDeb"], seed=seed, traj=traj,
             nsub=nsub_total, used=used_all, R_out=o["R"], V_out=o["V"], psi_out=o["psi"], tPart_out=o["tPart"], t_out=o["t"])
    print("su_stream: %d substeps, %d jumps" % (nsub_total, int((used_all > 1).sum())))


def gen_md():
    md = po.RefMD()
    c = md.consts
    md.seed(12345)
    md.init()
    rng = np.random.default_rng(5)
    R = rng.uniform(0, c["L"], size=(3, md.N))
    md.set_state(R=R)
    md.set_controls(0.0, 0, 0)
    V = md.get_state()["V"].copy()
    md.accelerations()
    A = md.get_state()["A"].copy()
    md.mdstep()
    s1 = md.get_state()
    md.set_controls(0.0, 1, 0)
    md.mdstep()
    s2 = md.get_state()
    np.savez(os.path.join(OUT, "md_N4096.npz"), R=R, V=V, A=A, R1=s1["R"], V1=s1["V"], A1=s1["A"], R2=s2["R"], V2=s2["V"],
             A2=s2["A"], **{k: c[k] for k in c})
    print("md: A[:,0]=%s" % A[:, 0])


def gen_mc408():
    mc = po.RefMC408L()
    c = mc.consts
    rng = np.random.default_rng(11)
    n = mc.N
    psi = full_psi(rng, n, 7)
    psi[:16] = 0.0
    psi[:16, 1, 0] = 1.0
    V = rng.normal(size=(3, n)) * 0.5
    mc.set_state(V=V, psi=psi)
    nsub = 62
    for _ in range(nsub):
        used = mc.qstep(np.full(n, NOJUMP))
        assert used == n
    o = mc.get_state()
    keep = 512
    np.savez(os.path.join(OUT, "mc408l_nojump.npz"), psi=psi[:keep], Vx=V[0, :keep], psi_out=o["psi"][:keep], nsub=nsub,
             **{k: c[k] for k in c})
    print("mc408l_nojump: %d substeps" % nsub)


def gen_su_mainloop():
    """The reference's whole main loop, files included: init() with srand48(777), lasers off (Om = OmDP = 0: no P
    population, hence no jumps and no dependence on the random stream), run to tmax = 0.081 (41 MD steps, one output()
    at c0 = 39), writeConditions. The produced files are the fixture for the mdqt_run driver (gzip'd text)."""
    import gzip
    import shutil
    import tempfile
    ref = po.RefSU(Om=0.0, OmDP=0.0)
    d = tempfile.mkdtemp() + "/"
    ref.set_savedir(d)
    n = ref.init(777)
    nsub = ref.run_until(0.081, do_output=1)
    c0, counter = ref.counters()
    out = os.path.join(OUT, "su_mainloop")
    os.makedirs(out, exist_ok=True)
    keep = ["energies.dat", "ions_timestep%06d.dat" % c0, "conditions_timestep%06d.dat" % c0, "wvFns_timestep%06d.dat" % c0,
            "vel_distX_time000000.dat", "vel_distY_time000000.dat", "vel_distZ_time000000.dat", "statePopulationsVsVTime000000.dat",
            "VZERO_timestep%06d_interval0.dat" % c0]
    for f in keep:
        with open(os.path.join(d, f), "rb") as src, gzip.GzipFile(os.path.join(out, f + ".gz"), "wb", mtime=0) as dst:
            shutil.copyfileobj(src, dst)
    with open(os.path.join(out, "README"), "w") as f:
        f.write("reference main loop (SU:1139-1383), seed 777, Om=OmDP=0, tmax=0.081: N=%d, %d substeps, c0=%d, counter=%d\n"
                "files written by the unmodified reference (oracle/gen_golden.py: gen_su_mainloop)\n" % (n, nsub, c0, counter))
    print("su_mainloop: N=%d substeps=%d c0=%d counter=%d files=%s" % (n, nsub, c0, counter, sorted(os.listdir(d))[:4]))


def gen_su_ensemble_stats():
    """Ensemble statistics of the reference's own stochastic evolution (its own drand48 stream, srand48(2468)): 2048
    ions, no plasma forces, 1500 qstep() sweeps from random S states and a thermal v_x spread. Every 100 substeps:
    mean S/P/D populations, mean and variance of v_x. The GPU engine (Philox streams) must agree statistically."""
    ref = po.RefSU()
    rng = np.random.default_rng(31)
    n, nsub, every = 2048, 1500, 100
    psi = np.zeros((n, 12, 2))
    r1, r2 = rng.uniform(size=n), rng.uniform(size=n)
    psi[:, 0, 0] = np.sqrt(r1); psi[:, 1, 0] = np.sqrt(1 - r1) * np.sqrt(r2); psi[:, 1, 1] = np.sqrt(1 - r1) * np.sqrt(1 - r2)
    V = np.zeros((3, n)); V[0] = rng.normal(size=n) * 0.05
    ref.set_state(R=np.zeros((3, n)), V=V, psi=psi, tPart=np.zeros(n), t=1.0)
    seed_stream(ref, 2468)
    rows = []
    for s in range(1, nsub + 1):
        ref.qstep()
        if s % every == 0:
            st = ref.get_state()
            p = (st["psi"] ** 2).sum(axis=2)
            rows.append([s, p[:, :2].sum(axis=1).mean(), p[:, 2:6].sum(axis=1).mean(), p[:, 6:].sum(axis=1).mean(),
                         st["V"][0].mean(), st["V"][0].var(), p[:, 2:6].sum(axis=1).std(), (st["tPart"] < 50 * ref.consts["dtq"]).mean()])
    np.savez(os.path.join(OUT, "su_ensemble_stats.npz"), psi=psi, Vx=V[0], rows=np.array(rows), n=n, nsub=nsub, every=every)
    print("su_ensemble_stats: final popS,P,D = %s" % rows[-1][1:4])


def gen_schemes():
    """the 5-level 422 nm pump (MC422L), the 3-level test system (TS) and the FZ408L driver pieces
    (leap-frog step(), measureSpinUps(), Zfunc()), all produced by the unmodified reference programs."""
    # ---- MC422L: one pump MD step = ratio no-jump qstep() sweeps; then tagParticles() with an injected stream ----
    mc = po.RefMC422L()
    c = mc.consts
    rng = np.random.default_rng(21)
    n = mc.N
    psi = full_psi(rng, n, 5)
    psi[:16] = 0.0
    psi[:16, 0, 0] = 1.0
    V = rng.normal(size=(3, n)) * 0.5
    mc.set_state(V=V, psi=psi)
    nsub = int(c["ratio"])
    for _ in range(nsub):
        assert mc.qstep(np.full(n, NOJUMP)) == n
    o = mc.get_state()
    keep = 512
    # the tagger consumes 1 or 2 uniforms per ion from one stream; fixtures keep the per-ion pairs actually consumed
    u_seq = rng.uniform(size=2 * n)
    tagged, used = mc.tag(u_seq)
    u_tab = np.zeros((n, 2))
    cur = 0
    nr = (o["psi"] ** 2).sum(axis=2)
    for i in range(n):
        u_tab[i, 0] = u_seq[cur]; cur += 1
        if not (u_tab[i, 0] < nr[i, 0]) and (u_tab[i, 0] < nr[i, 0] + nr[i, 2] + nr[i, 3]):
            u_tab[i, 1] = u_seq[cur]; cur += 1
    assert cur == used
    np.savez(os.path.join(OUT, "mc422l_pump.npz"), psi=psi[:keep], Vx=V[0, :keep], psi_out=o["psi"][:keep], nsub=nsub,
             tag_u=u_tab[:keep], tagged=tagged[:keep], **{k: c[k] for k in c})
    print("mc422l_pump: %d substeps, %d of %d tagged" % (nsub, tagged[:keep].sum(), keep))

    # ---- TS: 40 no-jump sweeps with the optical-force kick, then one sweep where every 3rd ion jumps ----
    ts = po.RefTS(detuning=-0.5, Om=0.5)
    n = ts.N
    psi = full_psi(rng, n, 3)
    psi[:8] = 0.0
    psi[:8, 0, 0] = 1.0
    Vx = rng.normal(size=n) * 0.1
    tp = rng.uniform(size=n)
    ts.set_state(Vx=Vx, psi=psi, tPart=tp)
    for _ in range(40):
        assert ts.qstep(np.full(n, NOJUMP)) == n
    s1 = ts.get_state()
    # per-ion table (slot 0 rand, slot 3 randDir): as a stream, a jumping ion consumes 2 numbers, the others 1
    u5 = rng.uniform(size=(n, 5))
    u5[:, 0] = NOJUMP
    u5[::3, 0] = 1e-9
    seq = []
    for i in range(n):
        seq.append(u5[i, 0])
        if u5[i, 0] < 1e-6:
            seq.append(u5[i, 3])
    used = ts.qstep(np.array(seq))
    assert used == len(seq)
    s2 = ts.get_state()
    np.savez(os.path.join(OUT, "ts_three_state.npz"), psi=psi, Vx=Vx, tPart=tp, psi1=s1["psi"], Vx1=s1["Vx"], tPart1=s1["tPart"],
             u5=u5, psi2=s2["psi"], Vx2=s2["Vx"], tPart2=s2["tPart"], detuning=-0.5, Om=0.5, nsub=40)
    print("ts_three_state: %d jumps in the last sweep" % int((u5[:, 0] < 1e-6).sum()))

    # ---- FZ408L: init(), the first step() (t = 0, 2nd-order start), a later step(), measureSpinUps, Zfunc ----
    fz = po.RefFZ408L()
    c = fz.consts
    n = fz.init(4242)
    s0 = fz.get_state()
    fz.step()
    s1 = fz.get_state()
    fz.set_state(R=s1["R"], V=s1["V"], t=0.002)
    fz.step()
    s2 = fz.get_state()
    psi = full_psi(rng, n, 7)
    fz.set_state(psi=psi, n=n)
    u_seq = rng.uniform(size=2 * n)
    tagged, cnt, used = fz.measure(u_seq)
    u_tab = np.zeros((n, 2))
    cur = 0
    nr = (psi ** 2).sum(axis=2)
    for i in range(n):
        u_tab[i, 0] = u_seq[cur]; cur += 1
        c1 = nr[i, 0] + nr[i, 2]
        if not (u_tab[i, 0] < c1) and (u_tab[i, 0] < c1 + nr[i, 3] + nr[i, 4]):
            u_tab[i, 1] = u_seq[cur]; cur += 1
    assert cur == used
    vaf0 = fz.zfunc(0)
    Vb = s2["V"] * 0.5 + 1e-3
    fz.set_state(V=Vb, n=n)
    vaf1 = fz.zfunc(1)
    np.savez(os.path.join(OUT, "fz408l_driver.npz"), R0=s0["R"], V0=s0["V"], R1=s1["R"], V1=s1["V"], R2=s2["R"], V2=s2["V"],
             F2=s2["F"], psi=psi, tag_u=u_tab, tagged=tagged, n_tagged=cnt, vaf0=vaf0, vaf1=vaf1, Vb=Vb,
             **{k: c[k] for k in c})
    print("fz408l_driver: N=%d, %d spin-up, VAF %g -> %g" % (n, cnt, vaf0, vaf1))


def gen_recorders():
    """g(r) and the four power autocorrelations of the MD program, by the reference itself.
    The reference's loops always run over its compile-time N = 4096 and T = 2500 (about a minute per recorder): only
    the first 48 ions carry non-zero series, the rest are zeros, so the committed fixture stays small."""
    md = po.RefMD()
    c = md.consts
    md.seed(2024)
    md.init()
    rng = np.random.default_rng(31)
    R = rng.uniform(0, c["L"], size=(3, md.N))
    md.set_state(R=R)
    r, g, step, rmax = md.pair_correlation()
    T, n = md.T, 48
    v = rng.normal(size=(3, n, T)) / np.sqrt(c["Gamma"])
    # give the series some memory so that the lag dependence is not flat
    for k in range(1, T):
        v[:, :, k] = 0.9 * v[:, :, k - 1] + 0.436 * v[:, :, k]
    md.set_vstore(v)
    ac = [md.autocorr(w) for w in (1, 2, 3, 4)]
    np.savez_compressed(os.path.join(OUT, "md_recorders.npz"), R=R, gr_r=r, gr_g=g, pairPairStep=step, pairPairMax=rmax,
                        v_seed=31, n_series=n, T=T, vaf=ac[0], longvisc=ac[1], vcube=ac[2], vfourth=ac[3],
                        **{k: c[k] for k in c})
    print("md_recorders: %d g(r) bins, g max %.3f; VAF[0]=%.6g VAF[100]=%.6g" % (len(r), g.max(), ac[0][0], ac[0][100]))


def recorder_series(seed, n, T, Gamma):
    """The velocity series of the md_recorders fixture (regenerated from the seed instead of storing 2.9 MB)."""
    rng = np.random.default_rng(seed)
    rng.uniform(0, 1.0, size=(3, 4096))  # the positions drawn first in gen_recorders (same stream position)
    v = rng.normal(size=(3, n, T)) / np.sqrt(Gamma)
    for k in range(1, T):
        v[:, :, k] = 0.9 * v[:, :, k - 1] + 0.436 * v[:, :, k]
    return v


def gen_fz_loop():
    """The FZ408L time loop itself (FZ408L:1040-1072, through ref_fz_run_loop: the loop body re-typed around the reference's
    own step()/qstep()/measureSpinUps()/Zfunc() and globals): a new run from init(4242), tmax = 0.0201, pump window
    (0.0061, 0.0143), sampleFreq = 5. No jumps during the pump (injected rand = NOJUMP); the measurement draws follow."""
    fz = po.RefFZ408L()
    c = fz.consts
    n = fz.init(4242)
    s0 = fz.get_state()
    tmax, tstart, tend, sf = 0.0201, 0.0061, 0.0143, 5
    # count the pump sweeps with the loop's own time arithmetic
    t, k = 0.0, 0
    while t <= tmax + 0.0009:
        if t < tend and t > tstart:
            k += 1
        t += c["dtq"]
    rng = np.random.default_rng(41)
    meas = rng.uniform(size=2 * n)
    u = np.concatenate([np.full(k * n, NOJUMP), meas])
    fz.set_c0(-1)
    r = fz.run_loop(tmax, tstart, tend, sf, u)
    s1 = fz.get_state()
    nr = (s1["psi"] ** 2).sum(axis=2)  # psi is not touched after the pump window, so these are the measured norms
    u_tab = np.zeros((n, 2))
    cur = 0
    for i in range(n):
        u_tab[i, 0] = meas[cur]; cur += 1
        c1 = nr[i, 0] + nr[i, 2]
        if not (u_tab[i, 0] < c1) and (u_tab[i, 0] < c1 + nr[i, 3] + nr[i, 4]):
            u_tab[i, 1] = meas[cur]; cur += 1
    assert r["used"] == k * n + cur, (r["used"], k * n + cur)
    np.savez_compressed(os.path.join(OUT, "fz408l_loop.npz"), R0=s0["R"], V0=s0["V"], psi0=s0["psi"], R1=s1["R"], V1=s1["V"],
                        psi1=s1["psi"], t1=s1["t"], c0=r["c0"], iters=r["iters"], pump_sweeps=k, tag_u=u_tab, spin=r["spin"],
                        nspin=r["nspin"], vaf=r["vaf"], tmax=tmax, tstart=tstart, tend=tend, sampleFreq=sf,
                        **{kk: c[kk] for kk in c})
    print("fz408l_loop: N=%d, %d iterations, %d pump sweeps, c0=%d, %d spin-up, VAF %s" % (n, r["iters"], k, r["c0"], r["nspin"], r["vaf"]))


def gen_md_program():
    """The MD program's recording, instantaneous-anisotropy and laser-force stages (main() stages 5, 7, 8; MD:1090-1165) run by
    the reference's own functions in its own order through ref_md_run_stages, with run-time step counts (40, 30, 20, 20 instead of
    the compile-time 2500, 2500, 1012, 2000), from init() with std::mt19937 seeded 4321 (lattice + Maxwellian; no Monte-Carlo, no
    collisions: fully deterministic). The files its recorders write are the fixture of `mdqt_run --program md`."""
    import gzip
    import shutil
    import tempfile
    md = po.RefMD()
    md.seed(4321)
    md.init()
    d = tempfile.mkdtemp() + "/"
    md.lib.ref_md_run_stages.argtypes = [ctypes.c_char_p] + [ctypes.c_int] * 4
    md.lib.ref_md_run_stages(d.encode(), 40, 30, 20, 20)
    tags = np.zeros(md.N, dtype=np.uint8)
    md.lib.ref_md_get_tags(tags.ctypes.data_as(ctypes.c_void_p))
    out = os.path.join(OUT, "md_program")
    os.makedirs(out, exist_ok=True)
    files = sorted(os.listdir(d))
    for f in files:
        with open(os.path.join(d, f), "rb") as src, gzip.GzipFile(os.path.join(out, f + ".gz"), "wb", mtime=0) as dst:
            shutil.copyfileobj(src, dst)
    np.save(os.path.join(out, "tags.npy"), tags)
    with open(os.path.join(out, "README"), "w") as f:
        f.write("reference MD program stages 5, 7, 8 (MD:1090-1165) via oracle/ref_md_harness.cpp: ref_md_run_stages(40, 30, 20, 20), "
                "mt19937 seed 4321, N = 4096; files written by the reference's own recorders (oracle/gen_golden.py: gen_md_program)\n")
    print("md_program:", files, "tag counts", [(int((tags >> k) & 1).sum()) if False else int(((tags >> k) & 1).sum()) for k in range(4)])


def seed_stream(ref, seed):
    """srand48(seed) inside the harness process (the reference's drand48 stream is then its own, un-injected)."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.srand48(ctypes.c_long(seed))


# ---- coupled statistical fixture (north-star: "temperatures and state populations must agree statistically") ------------
COUPLED = dict(N0=500, tmax=2.4, seeds=list(range(101, 117)))


def coupled_worker(seed):
    """One reference trajectory of the COUPLED main loop (SU:1248-1381: forces() every 25 substeps, step(), qstep() with the
    lasers ON, output() every 40 MD steps), 1 thread, the reference's own drand48 stream (srand48(seed)). The reference
    hard-codes N0 = 3500 (#define, SU:69); N0 enters only init() and the directory name, so the start state for N0 = 500 is
    drawn by mdqt_io_init_su -- the replica of init() that tests/test_hostio.py pins bitwise to the reference -- and the
    box is set through the harness. Returns the parsed energies.dat rows and the mean S/P/D populations per output()."""
    import shutil
    import tempfile
    from mdqtplasmasims_b200 import hostio
    N0, tmax = COUPLED["N0"], COUPLED["tmax"]
    st = hostio.init_su(seed, N0=N0, Ge=0.1)
    ref = po.RefSU()
    ref.set_box(st["L"], st["lDeb"])
    ref.set_state(R=st["R"], V=st["V"], psi=st["psi"], tPart=st["tPart"], t=0.0)
    seed_stream(ref, seed + 100000)
    d = tempfile.mkdtemp() + "/"
    ref.set_savedir(d)
    ref.set_counters(-1, 0)                      # c0 = -1 after init() (SU:347)
    ref.lib.ref_su_set_Epot0(ref.epot())         # Epot0 = Epot (SU:345-346)
    nsub = ref.run_until(tmax, do_output=1)
    en = np.loadtxt(os.path.join(d, "energies.dat"), ndmin=2)
    pops = []
    for k in range(en.shape[0]):
        a = np.loadtxt(os.path.join(d, "statePopulationsVsVTime%06d.dat" % k), ndmin=2)
        pops.append([a[:, 1].mean(), a[:, 2].mean(), a[:, 3].mean(), a[:, 0].std()])
    fin = ref.get_state()
    recent = float((fin["tPart"] < 25 * ref.consts["dtq"] * 0.999).mean())  # ions that jumped within the last 25 substeps
    shutil.rmtree(d, ignore_errors=True)
    return dict(seed=seed, N=st["N"], nsub=nsub, energies=en, pops=np.array(pops), recent_jump_frac=recent,
                norm_final=float((fin["psi"] ** 2).sum(axis=(1, 2)).mean()))


def gen_su_coupled(procs=7):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(coupled_worker, COUPLED["seeds"], chunksize=1)
    nrow = min(r["energies"].shape[0] for r in res)
    np.savez(os.path.join(OUT, "su_coupled.npz"), N0=COUPLED["N0"], tmax=COUPLED["tmax"], seeds=np.array([r["seed"] for r in res]),
             N=np.array([r["N"] for r in res]), nsub=np.array([r["nsub"] for r in res]),
             energies=np.stack([r["energies"][:nrow] for r in res]), pops=np.stack([r["pops"][:nrow] for r in res]),
             recent_jump_frac=np.array([r["recent_jump_frac"] for r in res]), norm_final=np.array([r["norm_final"] for r in res]))
    e = np.stack([r["energies"][:nrow] for r in res])
    print("su_coupled: %d seeds, N=%s, %d outputs; <EkinX>(t_end)=%.5g <EkinY>=%.5g popP=%.4f" %
          (len(res), [r["N"] for r in res], nrow, e[:, -1, 1].mean(), e[:, -1, 2].mean(),
           np.mean([r["pops"][-1, 1] for r in res])))


# ---- the MC-tagging programs' own stages 4-6, statistically -------------------------------------------
MCPROG = dict(npre=200, nrec=200, seeds=list(range(301, 309)))


def mc_program_worker(job):
    """One run of main()'s stages 4-6 of MC408L (MC408L:1211-1244) or MC422L (MC422L:1178-1211) by the reference's own functions
    (ref_mc_run_stages / ref_m422_run_stages: collisional MD, the pump stage with the reference's pumpMDTimeSteps and ratio,
    tagParticles(), the recording stage), after init() under `seed`, 1 thread, its own mt19937 + drand48 streams. The Metropolis
    stage is skipped exactly as `mdqt_run --program mc408l|mc422l` skips it. Returns the parsed
    taggedMoments.dat and temperature.dat, the number of tagged ions and the mean level populations after the pump."""
    import shutil
    import tempfile
    which, seed = job
    ref = po.RefMC408L() if which == "mc408l" else po.RefMC422L()
    pre = "ref_mc" if which == "mc408l" else "ref_m422"
    getattr(ref.lib, pre + "_init")(ctypes.c_uint(seed))
    run = getattr(ref.lib, pre + "_run_stages")
    run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    d = tempfile.mkdtemp() + "/"
    npump = int(ref.consts["pumpMDTimeSteps"])
    ntag = run(d.encode(), MCPROG["npre"], npump, MCPROG["nrec"], 0.25)
    tm = np.loadtxt(os.path.join(d, "taggedMoments.dat"), ndmin=2)
    temp = np.loadtxt(os.path.join(d, "temperature.dat"), ndmin=1)
    vd0 = np.loadtxt(os.path.join(d, "vel_distX_timestep%06d.dat" % 0), ndmin=2)[:, 1]
    psi = ref.get_state()["psi"]
    shutil.rmtree(d, ignore_errors=True)
    return dict(which=which, seed=seed, ntag=ntag, npump=npump, ratio=int(ref.consts["ratio"]), moments=tm, temperature=temp,
                vel_dist0=vd0, pops=(psi ** 2).sum(axis=2).mean(axis=0))


def gen_mc_programs(procs=8):
    import multiprocessing as mp
    jobs = [(w, s) for s in MCPROG["seeds"] for w in ("mc408l", "mc422l")]
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(mc_program_worker, jobs, chunksize=1)
    out = dict(npre=MCPROG["npre"], nrec=MCPROG["nrec"], seeds=np.array(MCPROG["seeds"]))
    for w in ("mc408l", "mc422l"):
        rs = [r for r in res if r["which"] == w]
        out[w + "_ntag"] = np.array([r["ntag"] for r in rs])
        out[w + "_npump"] = rs[0]["npump"]
        out[w + "_ratio"] = rs[0]["ratio"]
        out[w + "_moments"] = np.stack([r["moments"] for r in rs])
        out[w + "_temperature"] = np.stack([r["temperature"] for r in rs])
        out[w + "_vel_dist0"] = np.stack([r["vel_dist0"] for r in rs]).astype(np.float32)
        out[w + "_pops"] = np.stack([r["pops"] for r in rs])
        print(w, "ntag", out[w + "_ntag"], "npump", rs[0]["npump"], "m1(0)", out[w + "_moments"][:, 0, 1].mean(),
              "m1(end)", out[w + "_moments"][:, -1, 1].mean(), "T(0)", out[w + "_temperature"][:, 0].mean(), "pops", out[w + "_pops"].mean(axis=0))
    np.savez_compressed(os.path.join(OUT, "mc_programs.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    po.build()
    if "--mainloop" in sys.argv:
        gen_su_mainloop()
        sys.exit(0)
    if "--fzloop" in sys.argv:
        gen_fz_loop()
        sys.exit(0)
    if "--recorders" in sys.argv:
        gen_recorders()
        sys.exit(0)
    if "--schemes" in sys.argv:
        gen_schemes()
        sys.exit(0)
    if "--ensemble" in sys.argv:
        gen_su_ensemble_stats()
        sys.exit(0)
    if "--mdprogram" in sys.argv:
        gen_md_program()
        sys.exit(0)
    if "--coupled" in sys.argv:
        gen_su_coupled()
        sys.exit(0)
    if "--mcprograms" in sys.argv:
        gen_mc_programs()
        sys.exit(0)
    gen_su_forces()
    gen_su_nojump()
    gen_su_jumps()
    gen_su_stream()
    gen_md()
    gen_mc408()
    gen_schemes()

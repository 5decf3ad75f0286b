"""Quick device timings, round 2 (development aid): in-graph per-kernel times of the MD step at the thesis shape with the
item-walking vs the CTA-tile force kernel and both lane mappings; batched ensembles; large N."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic


def state(N, L, B=1, seed=0):
    R = np.stack([synthetic.random_positions(N, L, seed=seed + b) for b in range(B)])
    psi = np.stack([synthetic.random_s_state(N, seed=seed + b) for b in range(B)])
    V, tp = np.zeros((B, 3, N)), np.zeros((B, N))
    return (R, V, psi, tp) if B > 1 else (R[0], V[0], psi[0], tp[0])


def md_step_times(eng, nmd, reps=3):
    eng.md_steps(nmd); eng.sync()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.4:
        eng.md_steps(nmd); eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.md_steps(nmd)
    eng.sync()
    wall = (time.perf_counter() - t0) / reps / nmd
    eng.enable_timing(2)
    eng.md_steps(nmd)
    out = [eng.kernel_time_ms(k)[0] * 1e3 for k in range(4)]
    eng.enable_timing(0)
    return wall * 1e6, out


def run(label, N, B=1, nmd=40, env=None, **over):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        p = su_params(n_ions=N, N0=N, n_traj=B, **over)
        eng = Engine(p)
    finally:
        for k in (env or {}):
            del os.environ[k]
    R, V, psi, tp = state(N, p.L, B)
    eng.upload(R=R, V=V, psi=psi, tPart=tp, t=0.0, substep=0)
    wall, (k1, k2, g12, g21) = md_step_times(eng, nmd)
    print("%-34s N=%d B=%d plan=%s: MD step %.2f us | in graph: K1 %.2f K2 %.2f gaps %.2f %.2f us | %.3e pairs/s in K1, %.3e ion-steps/s"
          % (label, N, B, eng.force_plan(), wall, k1, k2, g12, g21, float(N) * N * B / (k1 * 1e-6), 25.0 * N * B / (wall * 1e-6)), flush=True)
    eng.close()


e = Engine(su_params(n_ions=256, N0=256))
print("fp64 peak TFLOP/s:", e.fp64_peak_tflops(), flush=True)
e.close()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "pdl":  # run once with MDQT_PDL=0 and once without (the mode is read once per process)
    for N in (1000, 2048, 3000, 3500, 4096):
        run("items", N)
    run("items", 3653, plan_n=3500)
    run("items", 3653)
    run("items", 3500, B=8, nmd=4)
if which == "k1ab":
    run("items", 3500)
    run("items", 3653, plan_n=3500)
    run("items", 3500, B=64, nmd=4)
if which in ("all", "small"):
    run("items, auto lanes", 3500)
    run("items, plan_n (2 lanes)", 3500, plan_n=3500)
    run("tiles, auto lanes", 3500, env={"MDQT_K1_ITEMS": "0"})
    for N in (1000, 2048, 3000, 4096, 7000):
        run("items", N)
        run("tiles", N, env={"MDQT_K1_ITEMS": "0"})
if which in ("all", "batch"):
    for B in (8, 64):
        run("items", 3500, B=B, nmd=4)
        run("tiles", 3500, B=B, nmd=4, env={"MDQT_K1_ITEMS": "0"})
if which in ("all", "large"):
    for N in (20000, 100000):
        run("tiles", N, nmd=2)

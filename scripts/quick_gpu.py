"""Quick device timings (development aid): FP64 peak probe, force kernel at several N, MD step at N=3500."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic

def timeit(fn, eng, reps):
    fn(); eng.sync()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.4:  # let the clocks ramp (sync every call: launches are asynchronous)
        fn(); eng.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    eng.sync()
    return (time.perf_counter() - t0) / reps

e = Engine(su_params(n_ions=3500))
print("fp64 peak TFLOP/s:", e.fp64_peak_tflops(), flush=True)
for N, reps in ((3500, 200), (4096, 200), (20000, 20), (100000, 3), (300000, 1)):
    p = su_params(n_ions=N, N0=N)
    eng = Engine(p)
    eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N))
    dt = timeit(eng.forces, eng, reps)
    print("forces N=%d plan=%s: %.3f us  %.3e pairs/s  (%.2f TFLOP/s @34 flop/pair)" % (N, eng.force_plan(), dt * 1e6, N * N / dt, 34 * N * N / dt / 1e12), flush=True)
    if N <= 100000:
        dt2 = timeit(lambda: eng.step_qstep(25), eng, max(3, reps // 4))
        print("  25 substeps: %.3f us  %.3e ion-steps/s" % (dt2 * 1e6, 25 * N / dt2), flush=True)
        dt3 = timeit(lambda: eng.md_steps(10), eng, max(2, reps // 20)) / 10
        print("  md step: %.3f us  %.3e ion-steps/s" % (dt3 * 1e6, 25 * N / dt3), flush=True)
# ensemble batch
for B in (8, 64):
    N = 3500
    p = su_params(n_ions=N, n_traj=B)
    eng = Engine(p)
    eng.upload(R=np.stack([synthetic.random_positions(N, p.L, seed=b) for b in range(B)]), V=np.zeros((B, 3, N)),
               psi=np.stack([synthetic.random_s_state(N, seed=b) for b in range(B)]), tPart=np.zeros((B, N)))
    dt = timeit(eng.forces, eng, 10)
    print("ensemble B=%d forces plan=%s: %.3f us %.3e pairs/s" % (B, eng.force_plan(), dt * 1e6, B * N * N / dt), flush=True)
    dt3 = timeit(lambda: eng.md_steps(4), eng, 3) / 4
    print("  md step: %.3f us  %.3e ion-steps/s" % (dt3 * 1e6, 25 * N * B / dt3), flush=True)

"""Per-launch device time of the force / substep kernels vs N, from the engine's own CUDA-event timing (developer aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
for N in (128, 256, 512, 1024, 2048, 3500, 7000):
    p = su_params(n_ions=N, N0=N)
    eng = Engine(p)
    eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N), t=0.0, substep=0)
    eng.md_steps(200); eng.sync()
    eng.enable_timing(True)
    eng.md_steps(100)
    k1, _ = eng.kernel_time_ms(0); k2, _ = eng.kernel_time_ms(1)
    print("N=%5d plan=%s  K1 %.2f us  K2 %.2f us" % (N, eng.force_plan(), k1 * 1e3, k2 * 1e3), flush=True)
    eng.close()

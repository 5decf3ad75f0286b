"""MD-step timing at the thesis shape (developer aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
N = 3500
p = su_params(n_ions=N)
eng = Engine(p)
eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N), t=0.0, substep=0)
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.5:
    eng.md_steps(40); eng.sync()
for rep in range(3):
    t0 = time.perf_counter(); eng.md_steps(400); eng.sync(); dt = (time.perf_counter() - t0) / 400
    print("md step %.2f us  %.3e ion-steps/s  (PDL=%s)" % (dt * 1e6, 25 * N / dt, os.environ.get("MDQT_PDL", "1")), flush=True)

"""Small fixed workload for ncu: a few MD steps at N=3500 (thesis shape) and one force call at N=1e5."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "small"
if which == "small":
    N = 3500
    p = su_params(n_ions=N)
    eng = Engine(p)
    eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N), t=0.0, substep=0)
    eng.md_steps(60)
    eng.sync()
else:
    N = 100000
    p = su_params(n_ions=N, N0=N)
    eng = Engine(p)
    eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N), t=0.0, substep=0)
    eng.md_steps(2)
    eng.sync()
print("done", which)

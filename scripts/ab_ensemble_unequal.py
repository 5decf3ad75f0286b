"""A/B aid: the bench's unequal-N ensemble block alone (64 trajectories, N_b ~ Binomial(729*3500, 1/729)) -- MD step and in-graph
kernel times, plus a checksum of the state (the item list must not change a bit). Usage: MDQT_K1_ILIST=0|1 python scripts/ab_ensemble_unequal.py"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic

B, N0 = 64, 3500
counts = np.random.default_rng(4321).binomial(729 * N0, 1.0 / 729, size=B).astype(np.int32)
cap = int(counts.max())
p = su_params(n_ions=cap, N0=N0, n_traj=B, traj0=1000, seed=12345, plan_n=0)
e = Engine(p)
e.set_ion_counts(counts)
e.set_traj_seeds(np.arange(B, dtype=np.uint64) + 777)
R = np.stack([synthetic.random_positions(cap, p.L, seed=b) for b in range(B)])
psi = np.stack([synthetic.random_s_state(cap, seed=b) for b in range(B)])
e.upload(R=R, V=np.zeros((B, 3, cap)), psi=psi, tPart=np.zeros((B, cap)), t=0.0, substep=0)
e.md_steps(4); e.md_steps(4); e.sync()
s = e.download()
md5 = hashlib.md5(b"".join(s[k][b][..., :counts[b]].tobytes() for b in range(B) for k in ("R", "V"))).hexdigest()[:12]
t0 = time.perf_counter()
for _ in range(3): e.md_steps(4)
e.sync()
wall = (time.perf_counter() - t0) / 12 * 1e6
e.enable_timing(2); e.md_steps(4)
k = [e.kernel_time_ms(j)[0] * 1e3 for j in range(4)]
pairs = float((counts.astype(np.float64) ** 2).sum())
print("MDQT_K1_ILIST=%s: MD step %.1f us | K1 %.1f K2 %.1f us | %.3e pairs/s in K1, %.3e ion-steps/s | state md5 %s"
      % (os.environ.get("MDQT_K1_ILIST", "(default 1)"), wall, k[0], k[1], pairs / (k[0] * 1e-6), 25.0 * counts.sum() / (wall * 1e-6), md5), flush=True)
e.close()

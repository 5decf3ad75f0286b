"""Row-decomposed large-N MD step on G GPUs of one box, one Python thread per GPU (development aid; bench.py times the same
path under torchrun). Usage: python scripts/large_n_threads.py N G [nsteps]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic

N, G = int(sys.argv[1]), int(sys.argv[2])
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
N = (N // G) * G
p0 = su_params(n_ions=N, N0=N)
R = synthetic.random_positions(N, p0.L, seed=777)
psi = synthetic.random_s_state(N, 12, seed=777)
V, tp = np.zeros((3, N)), np.zeros(N)
uid = Engine.comm_unique_id() if G > 1 else None
res = [None] * G
bar = threading.Barrier(G)

def work(r):
    rows = N // G
    e = Engine(su_params(n_ions=N, N0=N, row0=r * rows, n_rows=rows, device=r, seed=777))
    if G > 1:
        e.comm_init(uid, r, G)
    e.upload(R=R, V=V, psi=psi, tPart=tp, t=0.0, substep=0)
    e.md_steps(2); e.sync(); bar.wait()
    t0 = time.perf_counter()
    e.md_steps(nsteps); e.sync(); bar.wait()
    dt = (time.perf_counter() - t0) / nsteps
    d = e.diagnostics()
    res[r] = (dt, d["epot"], e.force_plan())
    e.close()

th = [threading.Thread(target=work, args=(r,)) for r in range(G)]
[t.start() for t in th]; [t.join() for t in th]
dt = max(x[0] for x in res)
print("N=%d G=%d plan=%s: %.3f ms per MD step, %.3e pairs/s, epot %r" % (N, G, res[0][2], dt * 1e3, float(N) * N / dt, res[0][1]), flush=True)

"""Developer aid: time the force kernel under explicit plans (env knobs of plan_force) at one N.
Usage: python scripts/ab_plan.py N "RG,JSUB,NSPLIT[,IPT]" ...   (0 / auto = leave to the planner)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
N = int(sys.argv[1])
B = int(os.environ.get("AB_B", "1"))
p = su_params(n_ions=N, N0=N, n_traj=B)
eng = Engine(p)
eng.upload(R=synthetic.random_positions(N, p.L) if B == 1 else np.stack([synthetic.random_positions(N, p.L, seed=b) for b in range(B)]))
eng.forces(); eng.sync()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.3: eng.forces()
eng.sync()
reps = max(5, int(0.3 / max(1e-5, (B * N * N / 3.5e11))))
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(reps): eng.forces()
    eng.sync()
    best = min(best, (time.perf_counter() - t0) / reps)
print("  N=%%d plan=%%s %%.2f us %%.3e pairs/s" %% (N, eng.force_plan(), best * 1e6, B * N * N / best), flush=True)
''' % ROOT
N = sys.argv[1]
for spec in sys.argv[2:]:
    env = dict(os.environ)
    for k in ("MDQT_FORCE_RG", "MDQT_FORCE_JSUB", "MDQT_FORCE_NSPLIT", "MDQT_FORCE_IPT"):
        env.pop(k, None)
    if spec != "auto":
        f = spec.split(",")
        env["MDQT_FORCE_RG"], env["MDQT_FORCE_JSUB"], env["MDQT_FORCE_NSPLIT"] = f[0], f[1], f[2]
        if len(f) > 3:
            env["MDQT_FORCE_IPT"] = f[3]
    out = subprocess.run([sys.executable, "-c", CHILD, N], env=env, capture_output=True, text=True)
    print(spec, (out.stdout.rstrip() or out.stderr[-400:]), flush=True)

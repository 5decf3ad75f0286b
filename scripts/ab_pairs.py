"""A/B timing of the force kernel for a few plans (developer aid). Usage: python scripts/ab_pairs.py"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
N = int(sys.argv[1])
p = su_params(n_ions=N, N0=N)
eng = Engine(p)
eng.upload(R=synthetic.random_positions(N, p.L))
eng.forces(); eng.sync()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.5: eng.forces()
eng.sync()
reps = max(2, int(0.5 / max(1e-5, (N * N / 3.5e11))))
t0 = time.perf_counter()
for _ in range(reps): eng.forces()
eng.sync()
dt = (time.perf_counter() - t0) / reps
print("N=%%d plan=%%s %%.1f us %%.3e pairs/s" %% (N, eng.force_plan(), dt * 1e6, N * N / dt))
''' % ROOT
for N in (3500, 100000):
    for lib in ("libmdqt_b200.so", "libmdqt_b200_v0.so"):
        for ipt in ("1", "2"):
            for ns in ((None, "16", "24", "32", "48") if N == 3500 else (None, "8", "24")):
                env = dict(os.environ, MDQT_LIB_PATH=os.path.join(ROOT, "mdqtplasmasims_b200", lib), MDQT_FORCE_IPT=ipt)
                if ns: env["MDQT_FORCE_NSPLIT"] = ns
                out = subprocess.run([sys.executable, "-c", CHILD, str(N)], env=env, capture_output=True, text=True)
                print(lib, "ipt", ipt, "ns", ns, "|", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)

"""A/B timing of the force kernel (developer aid). One process per library variant; plans via env knobs.
Usage: python scripts/ab_pairs.py [lib.so ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time
sys.path.insert(0, %r)
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
from mdqtplasmasims_b200 import engine as E
cases = eval(sys.argv[1])
for (N, ipt, ns) in cases:
    os.environ["MDQT_FORCE_IPT"] = str(ipt)
    if ns: os.environ["MDQT_FORCE_NSPLIT"] = str(ns)
    else: os.environ.pop("MDQT_FORCE_NSPLIT", None)
    p = su_params(n_ions=N, N0=N)
    eng = Engine(p)
    eng.upload(R=synthetic.random_positions(N, p.L))
    eng.forces(); eng.sync()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.3: eng.forces()
    eng.sync()
    reps = max(2, int(0.3 / max(1e-5, (N * N / 3.5e11))))
    t0 = time.perf_counter()
    for _ in range(reps): eng.forces()
    eng.sync()
    dt = (time.perf_counter() - t0) / reps
    print("  N=%%d ipt=%%d plan=%%s %%.1f us %%.3e pairs/s" %% (N, ipt, eng.force_plan(), dt * 1e6, N * N / dt), flush=True)
    eng.close()
''' % ROOT
cases = eval(os.environ.get('AB_CASES', '[(3500, 1, None)]'))
libs = sys.argv[1:] or ["libmdqt_b200.so"]
for lib in libs:
    env = dict(os.environ, MDQT_LIB_PATH=os.path.join(ROOT, "mdqtplasmasims_b200", lib))
    print(lib, flush=True)
    out = subprocess.run([sys.executable, "-c", CHILD, repr(cases)], env=env, capture_output=True, text=True)
    print(out.stdout.rstrip() or out.stderr[-500:], flush=True)

"""A/B of substep-kernel build variants (developer aid): one subprocess per library; prints the in-graph kernel times at the
thesis shape and a checksum of the state after 80 MD steps (identical checksums = identical bits, jumps included).
Usage: python scripts/ab_k2.py lib1.so lib2.so ...   (names inside mdqtplasmasims_b200/)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, hashlib
sys.path.insert(0, %r)
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic
N = 3500
for label, over in (("auto lanes", {}), ("two lanes ", {"plan_n": 3500})):
    p = su_params(n_ions=N, N0=N, seed=99, **over)
    e = Engine(p)
    e.upload(R=synthetic.random_positions(N, p.L, seed=1), V=np.zeros((3, N)), psi=synthetic.random_s_state(N, seed=1), tPart=np.zeros(N), t=0.0, substep=0)
    e.md_steps(40); e.md_steps(40); e.sync()
    s = e.download()
    h = hashlib.md5(s["psi"].tobytes() + s["V"].tobytes() + s["R"].tobytes() + s["tPart"].tobytes()).hexdigest()[:12]
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        e.md_steps(40); e.sync()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps): e.md_steps(40)
    e.sync()
    wall = (time.perf_counter() - t0) / reps / 40 * 1e6
    e.enable_timing(2); e.md_steps(40)
    k = [e.kernel_time_ms(j)[0] * 1e3 for j in range(4)]
    e.enable_timing(0)
    print("  %%s: MD step %%.2f us | K1 %%.2f K2 %%.2f gaps %%.2f %%.2f | state md5 %%s jumps-in-last-step %%d" %% (label, wall, k[0], k[1], k[2], k[3], h, int((s["tPart"] < 25 * p.dtq * 0.999).sum())), flush=True)
    e.close()
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, MDQT_LIB_PATH=os.path.join(ROOT, "mdqtplasmasims_b200", lib))
    print(lib, flush=True)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(out.stdout.rstrip() or out.stderr[-800:], flush=True)

"""Developer aid: per-CTA phase timeline of the force kernel at small N (needs a library built with -DMDQT_K1_TRACE).
Usage: MDQT_LIB_PATH=.../libv_trace.so python scripts/k1_trace.py [N]"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic, load_library

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3500
p = su_params(n_ions=N, N0=N)
eng = Engine(p)
eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N))
for _ in range(200):
    eng.forces()
eng.sync()
eng.forces(); eng.sync()
ns, jlen = eng.force_plan()
buf = np.zeros(8 * 8192, dtype=np.int64)
load_library().mdqt_debug_read_trace(ctypes.c_void_p(buf.ctypes.data), buf.size)
tr = buf.reshape(8192, 8)
tr = tr[tr[:, 0] > 0]
t0 = tr[:, 0].min()
rel = (tr[:, :5] - t0) / 1e3
last = tr[:, 4] >= tr[:, 3]
print("N=%d plan=(%d,%d) CTAs=%d" % (N, ns, jlen, len(tr)))
print("CTA start   us: min %.2f med %.2f max %.2f" % (rel[:, 0].min(), np.median(rel[:, 0]), rel[:, 0].max()))
print("prologue    us: med %.2f max %.2f" % (np.median(rel[:, 1] - rel[:, 0]), (rel[:, 1] - rel[:, 0]).max()))
print("main loop   us: med %.2f max %.2f" % (np.median(rel[:, 2] - rel[:, 1]), (rel[:, 2] - rel[:, 1]).max()))
print("partials    us: med %.2f max %.2f" % (np.median(rel[:, 3] - rel[:, 2]), (rel[:, 3] - rel[:, 2]).max()))
lr = rel[last]
print("last-CTA reduce us (%d CTAs): med %.2f max %.2f" % (len(lr), np.median(lr[:, 4] - lr[:, 3]), (lr[:, 4] - lr[:, 3]).max()))
print("main-loop end us: min %.2f med %.2f max %.2f" % (rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max()))
print("kernel end   us: %.2f (last stamp)" % max(rel[:, 3].max(), lr[:, 4].max()))
sm = tr[:, 7]
cnt = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM: min %d max %d ; SMs used %d" % (cnt[cnt > 0].min(), cnt.max(), (cnt > 0).sum()))

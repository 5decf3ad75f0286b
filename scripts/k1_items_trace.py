"""Developer aid: per-warp phase timeline of the item-walking force kernel (library built with -DMDQT_K1_TRACE).
Usage: MDQT_LIB_PATH=.../libv_trace.so python scripts/k1_items_trace.py [N]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mdqtplasmasims_b200 import Engine, su_params, synthetic, load_library

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3500
p = su_params(n_ions=N, N0=N)
eng = Engine(p)
eng.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N))
eng.md_steps(40); eng.md_steps(40); eng.sync()   # inside the replayed graph: the last force launch leaves its stamps
buf = np.zeros(8 * 8192, dtype=np.int64)
load_library().mdqt_debug_read_trace(ctypes.c_void_p(buf.ctypes.data), buf.size)
tr = buf.reshape(8192, 8)
tr = tr[tr[:, 0] > 0]
t0 = tr[:, 0].min()
rel = (tr[:, :5] - t0) / 1e3
busy = tr[:, 3] > tr[:, 0]  # warps that had an item
print("N=%d plan=%s warps=%d with an item=%d" % (N, eng.force_plan(), len(tr), busy.sum()))
print("warp start      us: min %.2f med %.2f max %.2f" % (rel[:, 0].min(), np.median(rel[:, 0]), rel[:, 0].max()))
print("prologue        us: med %.2f max %.2f   (entry -> table + first tile landed, barrier)" % (np.median(rel[:, 1] - rel[:, 0]), (rel[:, 1] - rel[:, 0]).max()))
b = rel[busy]
print("loop entry      us: med %.2f max %.2f   (decode next, prefetch, wait)" % (np.median(b[:, 2] - b[:, 1]), (b[:, 2] - b[:, 1]).max()))
d = b[:, 3] - b[:, 2]
print("pair loop       us: min %.2f med %.2f max %.2f" % (d.min(), np.median(d), d.max()))
print("loop end        us: min %.2f med %.2f max %.2f" % (b[:, 3].min(), np.median(b[:, 3]), b[:, 3].max()))
print("warp exit       us: med %.2f max %.2f" % (np.median(rel[:, 4]), rel[:, 4].max()))
sm = tr[:, 7].astype(int)
cnt = np.bincount(sm, minlength=148)
print("warps per SM: min %d max %d ; SMs used %d" % (cnt[cnt > 0].min(), cnt.max(), (cnt > 0).sum()))
per_sm_end = np.array([rel[sm == k, 3].max() if (sm == k).any() else 0 for k in range(148)])
print("per-SM last loop end us: min %.2f med %.2f max %.2f" % (per_sm_end[per_sm_end > 0].min(), np.median(per_sm_end[per_sm_end > 0]), per_sm_end.max()))
# structure of the unfairness: mean pair-loop time by warp index within the CTA and by CTA residency order on its SM
w_idx = np.arange(len(tr)) % 8            # trace slot = blockIdx.x * 8 + warp
cta = np.arange(len(tr)) // 8
dur = rel[:, 3] - rel[:, 2]
print("pair loop us by warp index:", " ".join("%.1f" % dur[(w_idx == k) & busy].mean() for k in range(8)))
print("loop END us by warp index :", " ".join("%.1f" % rel[(w_idx == k) & busy, 3].mean() for k in range(8)))
first = np.zeros(len(tr), dtype=bool)
for s_ in range(148):
    ids = np.unique(cta[sm == s_])
    if len(ids) >= 1:
        first[cta == ids.min()] = True
print("pair loop us: lower-index CTA on its SM %.2f, the other %.2f" % (dur[first & busy].mean(), dur[~first & busy].mean()))
for k in range(4):
    sel = busy & ((w_idx % 4) == k)
    print("scheduler slot %d (warps %d, %d): loop end us mean %.2f max %.2f" % (k, k, k + 4, rel[sel, 3].mean(), rel[sel, 3].max()))
# one SM in detail
s0 = int(np.argmax(per_sm_end))
print("SM %d: (cta, warp, loop start, loop end)" % s0)
for i in np.where(sm == s0)[0]:
    print("   cta %4d warp %d  %.2f -> %.2f" % (cta[i], w_idx[i], rel[i, 2], rel[i, 3]))

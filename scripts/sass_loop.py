"""Developer aid: per-pair instruction mix of the hottest (longest) loop of a kernel, from cuobjdump -sass.
Usage: python scripts/sass_loop.py <lib.so> <mangled-kernel-substring> <pairs-per-iteration>"""
import collections
import re
import subprocess
import sys

lib, pat, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    loops = []
    for k, (addr, text) in enumerate(ins):
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    for (tgt, addr) in loops:  # innermost loops only: no other backward branch inside
        if any(t2 >= tgt and a2 < addr for (t2, a2) in loops if (t2, a2) != (tgt, addr)):
            continue
        body = [t for a, t in ins if tgt <= a <= addr]
        if best is None or len(body) > len(best):
            best = body
    cnt = collections.Counter()
    for t in best:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        cnt[t.split()[0].split(".")[0]] += 1
    fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(name)
    print("  loop body %d instr, %.2f per pair; FP64-pipe %.2f per pair, other %.2f" % (len(best), len(best) / per, fp64 / per, (len(best) - fp64) / per))
    print("  " + ", ".join("%s %.2f" % (k, v / per) for k, v in cnt.most_common()))

"""Developer aid: the SASS of the hottest (longest innermost) loop of a kernel, with the per-pair instruction mix on top.
Usage: python scripts/sass_listing.py <lib.so> <mangled-kernel-substring> <pairs-per-iteration>"""
import collections, re, subprocess, sys
lib, pat, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in re.split(r"\n\s*Function : ", txt):
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", b)]
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    best = None
    for tg, ad in loops:
        if any(t2 >= tg and a2 < ad for t2, a2 in loops if (t2, a2) != (tg, ad)):
            continue
        body = [(a, t) for a, t in ins if tg <= a <= ad]
        if best is None or len(body) > len(best):
            best = body
    cnt = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in best)
    fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print("kernel %s" % name)
    print("innermost loop: %d instructions = %.2f per ordered pair (%g pairs per trip); FP64-pipe %.2f, other %.2f per pair" %
          (len(best), len(best) / per, per, fp64 / per, (len(best) - fp64) / per))
    print("mix per pair: " + ", ".join("%s %.2f" % (k, v / per) for k, v in cnt.most_common()))
    print("issue-port cycles per pair (FP64 instructions hold the port 2 cycles): %.1f" % ((len(best) + fp64) / per))
    for a, t in best:
        print("  /*%04x*/  %s" % (a, t))
    print()

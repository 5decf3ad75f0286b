"""A/B of force-kernel build variants (developer aid): in-graph kernel times at the thesis shape, one trajectory and 64.
Usage: python scripts/ab_k1.py lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    env = dict(os.environ, MDQT_LIB_PATH=os.path.join(ROOT, "mdqtplasmasims_b200", lib))
    print(lib, flush=True)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "quick2.py"), "k1ab"], env=env, capture_output=True, text=True)
    print(out.stdout.rstrip() or out.stderr[-800:], flush=True)

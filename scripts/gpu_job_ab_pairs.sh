# A/B of the pair arithmetic (profiles/r02o_ab_pairs.log). Build the two libraries first:
#   python -c "from mdqtplasmasims_b200 import build as b; b.build(force=True, out='mdqtplasmasims_b200/lib_o3.so', defines=['MDQT_RSQRT_ORDER=3']); b.build(force=True, out='mdqtplasmasims_b200/lib_o2.so', defines=['MDQT_RSQRT_ORDER=2'])"
set -x
for lib in lib_o3.so lib_o2.so; do
  echo "=== $lib"
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/force_accuracy.py
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/quick2.py k1ab
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/quick2.py large
done

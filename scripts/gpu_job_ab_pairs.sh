set -x
for lib in lib_o3.so lib_o2.so; do
  echo "=== $lib"
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/force_accuracy.py
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/quick2.py k1ab
  MDQT_LIB_PATH=$PWD/mdqtplasmasims_b200/$lib python scripts/quick2.py large
done

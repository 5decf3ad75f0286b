MDQT_QT_LANES=8 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
MDQT_QT_LANES=8 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "nojump_golden or jump_table or trajectory_golden or renorm" 2>&1 | tail -3
for L in 4 8; do echo "lanes $L"; MDQT_QT_LANES=$L python scripts/ab_k2.py libmdqt_b200.so 2>&1 | tail -1; done

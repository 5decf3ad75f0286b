set -x
timeout 900 python -m pytest tests/test_gpu_mc_programs.py -q -x -s > gpurun_out/r02t_mcprog.log 2>&1; tail -40 gpurun_out/r02t_mcprog.log
cat > /tmp/prof_small.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from mdqtplasmasims_b200 import Engine, su_params, synthetic
N = int(sys.argv[1]); nmd = int(sys.argv[2])
p = su_params(n_ions=N, N0=N)
e = Engine(p)
e.upload(R=synthetic.random_positions(N, p.L), V=np.zeros((3, N)), psi=synthetic.random_s_state(N), tPart=np.zeros(N), t=0.0, substep=0)
for _ in range(3): e.md_steps(nmd)
e.sync()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_substeps4 -s 20 -c 1 -o gpurun_out/r02t_sub_small -f python /tmp/prof_small.py 3500 10 > gpurun_out/r02t_ncu_sub_small.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs_items -s 20 -c 1 -o gpurun_out/r02t_pairs_small -f python /tmp/prof_small.py 3500 10 > gpurun_out/r02t_ncu_pairs_small.log 2>&1
ls -la gpurun_out | tail -8

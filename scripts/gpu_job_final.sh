set -x
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02z_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02z_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_gpu1.json 2> gpurun_out/r02z_bench_gpu1.err
tail -c 200 gpurun_out/r02z_bench_gpu1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_mdstep_N3500.csv python bench.py --steps 2 --warmup 3 --ensemble 0 --large-n 0 --large-n2 0 --no-md-family --no-cpu-baseline > gpurun_out/r02z_ncu_launches.log 2>&1
D=/tmp/mdqt_one; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --saveDirectory $D/a/ ) > gpurun_out/r02z_thesis_run.log 2>&1
grep "mdqt_run:" gpurun_out/r02z_thesis_run.log | cut -c1-130
rm -rf $D
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
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pairs_items -s 20 -c 1 -o gpurun_out/r02z_pairs_small -f python /tmp/prof_small.py 3500 10 > gpurun_out/r02z_ncu_pairs_small.log 2>&1
ls gpurun_out | tail -12

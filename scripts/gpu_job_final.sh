set -x
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02s_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02s_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02s_smoke.log 2>&1; tail -2 gpurun_out/r02s_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_bench_gpu1.json 2> gpurun_out/r02s_bench_gpu1.err
tail -c 200 gpurun_out/r02s_bench_gpu1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches_mdstep_N3500.csv python bench.py --steps 2 --warmup 3 --ensemble 0 --large-n 0 --large-n2 0 --no-md-family --no-cpu-baseline > gpurun_out/r02s_ncu_launches.log 2>&1
D=/tmp/mdqt_one; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --saveDirectory $D/a/ ) > gpurun_out/r02s_thesis_run.log 2>&1
grep "mdqt_run:" gpurun_out/r02s_thesis_run.log | cut -c1-130
rm -rf $D
python scripts/quick2.py small > gpurun_out/r02s_quick.log 2>&1; python scripts/quick2.py batch >> gpurun_out/r02s_quick.log 2>&1
grep -E "^items|^tiles" gpurun_out/r02s_quick.log | cut -c1-170

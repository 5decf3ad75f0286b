set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02n_bench_gpu8.json 2> gpurun_out/r02n_bench_gpu8.err
tail -c 300 gpurun_out/r02n_bench_gpu8.err
for G in 1 2 4 8; do python scripts/large_n_threads.py 200000 $G 4; done > gpurun_out/r02n_large_n_threads.log 2>&1
python scripts/large_n_threads.py 1000000 8 3 >> gpurun_out/r02n_large_n_threads.log 2>&1
grep "^N=" gpurun_out/r02n_large_n_threads.log
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2

set -x
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02r_bench_gpu2.json 2> gpurun_out/r02r_bench_gpu2.err
tail -c 300 gpurun_out/r02r_bench_gpu2.err
for G in 1 2; do python scripts/large_n_threads.py 200000 $G 4; done 2>&1 | grep "^N="

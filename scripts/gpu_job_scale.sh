# Round-2 multi-GPU measurements on one 8 x B200 box (run through gpurun --gpus 8); results land in gpurun_out/r02h_*.
set -x
nvidia-smi -L | head -8; nproc; df -h /tmp | tail -1
python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02h_pytest_multi_gpu8.log 2>&1; tail -3 gpurun_out/r02h_pytest_multi_gpu8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_gpu8.json 2> gpurun_out/r02h_bench_gpu8.err
tail -c 400 gpurun_out/r02h_bench_gpu8.err
# BASELINE configs[3] as a product: the SLURM array replaced by one command -- 512 jobs (each draws its own N), 64 per handle, 8 GPUs,
# the thesis parameters to tmax = 30, every job's own directory with all its files
D=/tmp/mdqt_ens; rm -rf $D; mkdir -p $D
FREE=$(df --output=avail -BG /tmp | tail -1 | tr -dc 0-9)
SF=40; if [ "$FREE" -lt 60 ]; then SF=400; fi
( time ./mdqtplasmasims_b200/mdqt_run --jobs 1-512 --batch 64 --gpus 8 --tmax 30 --seed 1000 --sampleFreq $SF --saveDirectory $D/ ) > gpurun_out/r02h_ensemble512.log 2>&1
echo "sampleFreq=$SF free_GB=$FREE files=$(find $D -type f | wc -l) bytes=$(du -sb $D | cut -f1)" >> gpurun_out/r02h_ensemble512.log
grep -v "^[0-9]*$" gpurun_out/r02h_ensemble512.log | tail -14
J=$(ls -d $D/*/job77); ls $J | head -5; tail -2 $J/energies.dat
rm -rf $D
# the thesis run itself (one job, tmax = 30) on one GPU: batch-reproducible default and --fast-single
D=/tmp/mdqt_one; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --saveDirectory $D/a/ ) > gpurun_out/r02h_thesis_run.log 2>&1
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --fast-single --saveDirectory $D/b/ ) >> gpurun_out/r02h_thesis_run.log 2>&1
grep -v "^[0-9]*$" gpurun_out/r02h_thesis_run.log | tail -12
rm -rf $D
# BASELINE configs[4] through the driver: one large job row-decomposed over 8 GPUs inside the library
D=/tmp/mdqt_big; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run 1 --N0 200000 --gpus 8 --tmax 0.0199 --seed 3 --saveDirectory $D/ ) > gpurun_out/r02h_large_run_gpu8.log 2>&1
grep -v "^[0-9]*$" gpurun_out/r02h_large_run_gpu8.log | tail -6
rm -rf $D

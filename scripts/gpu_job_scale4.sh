set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02r_bench_gpu8.json 2> gpurun_out/r02r_bench_gpu8.err
tail -c 300 gpurun_out/r02r_bench_gpu8.err
D=/tmp/mdqt_ens; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run --jobs 1-512 --batch 64 --gpus 8 --tmax 30 --seed 1000 --saveDirectory $D/ ) > gpurun_out/r02r_ensemble512.log 2>&1
echo "files=$(find $D -type f | wc -l) bytes=$(du -sb $D | cut -f1)" >> gpurun_out/r02r_ensemble512.log
# one job of the array alone: byte-identical files
J=137; mkdir -p /tmp/mdqt_one
./mdqtplasmasims_b200/mdqt_run $J --tmax 30 --seed $((1000 + J)) --saveDirectory /tmp/mdqt_one/ >> gpurun_out/r02r_ensemble512.log 2>&1
A=$(find $D -type d -name "job$J"); B=$(find /tmp/mdqt_one -type d -name "job$J")
diff -rq $A $B >> gpurun_out/r02r_ensemble512.log 2>&1 && echo "job $J alone == job $J inside the 512-job array: $(ls $A | wc -l) files byte-identical" >> gpurun_out/r02r_ensemble512.log
grep -v "^[0-9]*$" gpurun_out/r02r_ensemble512.log | tail -12
rm -rf $D /tmp/mdqt_one

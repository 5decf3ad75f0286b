set -x
python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02i_pytest_multi_gpu8.log 2>&1; tail -3 gpurun_out/r02i_pytest_multi_gpu8.log
for G in 1 2 4 8; do python scripts/large_n_threads.py 200000 $G 4; done > gpurun_out/r02i_large_n_threads.log 2>&1
python scripts/large_n_threads.py 1000000 8 3 >> gpurun_out/r02i_large_n_threads.log 2>&1
cat gpurun_out/r02i_large_n_threads.log
D=/tmp/mdqt_ens; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run --jobs 1-512 --batch 64 --gpus 8 --tmax 30 --seed 1000 --saveDirectory $D/ ) > gpurun_out/r02i_ensemble512.log 2>&1
echo "files=$(find $D -type f | wc -l) bytes=$(du -sb $D | cut -f1)" >> gpurun_out/r02i_ensemble512.log
grep -v "^[0-9]*$" gpurun_out/r02i_ensemble512.log | tail -14
J=$(ls -d $D/*/job77); ls $J | wc -l; tail -2 $J/energies.dat
rm -rf $D
D=/tmp/mdqt_one; rm -rf $D; mkdir -p $D
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --saveDirectory $D/a/ ) > gpurun_out/r02i_thesis_run.log 2>&1
( time ./mdqtplasmasims_b200/mdqt_run 1 --tmax 30 --seed 7 --fast-single --saveDirectory $D/b/ ) >> gpurun_out/r02i_thesis_run.log 2>&1
grep -v "^[0-9]*$" gpurun_out/r02i_thesis_run.log | tail -12
rm -rf $D

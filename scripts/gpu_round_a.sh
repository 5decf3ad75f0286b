#!/bin/bash
# one GPU call: parity suite on the default build, then A/B of the pair-kernel arithmetic variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_a.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
tail -5 gpurun_out/pytest_gpu_a.log
AB_CASES='[(3500,1,None),(100000,2,None)]' timeout 600 python scripts/ab_pairs.py libv_base.so libv_scaled.so libv_int.so libv_t1024.so libmdqt_b200.so > gpurun_out/ab_a.log 2>&1
cat gpurun_out/ab_a.log
timeout 300 python scripts/quick_gpu.py > gpurun_out/quick_a.log 2>&1
cat gpurun_out/quick_a.log

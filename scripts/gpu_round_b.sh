#!/bin/bash
# Round evidence run (one GPU): parity suite, both bench arms, the full thesis run through mdqt_run, ncu launch list +
# full captures of the two hot kernels. Every ncu command runs only after the same command exited 0 without ncu.
set -u
O=gpurun_out; T=${TAG:-r01c}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log; tail -3 $O/${T}_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/${T}_bench_gpu1.json 2> $O/${T}_bench_gpu1.err; echo "bench rc=$?"
tail -c 600 $O/${T}_bench_gpu1.json | head -c 300; echo
# the thesis run end to end: N0=3500, tmax=30 (15000 MD steps, 375 output() calls), files included
rm -rf /tmp/mdqt_thesis && mkdir -p /tmp/mdqt_thesis
( time ./mdqtplasmasims_b200/mdqt_run 1 --saveDirectory /tmp/mdqt_thesis/ --tmax 30 --seed 4242 ) > $O/${T}_thesis_run.log 2>&1; echo "thesis rc=$?"
tail -5 $O/${T}_thesis_run.log
ls /tmp/mdqt_thesis/*/job1 | wc -l >> $O/${T}_thesis_run.log
tail -3 /tmp/mdqt_thesis/*/job1/energies.dat >> $O/${T}_thesis_run.log
# ncu: launch list of the MD loop, then full captures
python scripts/prof_step.py small > $O/${T}_plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file $O/${T}_launches_mdstep_N3500.csv python scripts/prof_step.py small > /dev/null 2>&1
python scripts/prof_step.py small > $O/${T}_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 30 -c 1 -f -o $O/${T}_pairs_small python scripts/prof_step.py small > /dev/null 2>&1
python scripts/prof_step.py small > $O/${T}_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_substeps -s 30 -c 1 -f -o $O/${T}_sub_small python scripts/prof_step.py small > /dev/null 2>&1
python scripts/prof_step.py large > $O/${T}_plain_large.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 1 -c 1 -f -o $O/${T}_pairs_large python scripts/prof_step.py large > /dev/null 2>&1
ls -la $O | tail -12

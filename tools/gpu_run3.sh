#!/usr/bin/env bash
# gpurun call 3: GPU tests (incl. plugin), bench on three workloads, sort-pass configuration sweep, ncu full captures at bench size
mkdir -p gpurun_out/r3
O=gpurun_out/r3
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_dna30.json 2> $O/bench_dna30.err
cat $O/bench_dna30.json
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
for b in build/sb/sb_*; do timeout 120 $b 28 48 >> $O/sortbench.txt 2>&1; done
cat $O/sortbench.txt
# ncu --set full at the bench size: the big radix pass (3rd onesweep launch of the step) and the LPF kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 2 -c 1 -o $O/ncu_onesweep_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_onesweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpf_tile_kernel -s 0 -c 1 -o $O/ncu_lpf_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_lpf.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_dna30.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
ls -la $O

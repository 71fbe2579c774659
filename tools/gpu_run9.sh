#!/usr/bin/env bash
# gpurun call 9 (1 GPU): windowed look-back in the sort pass (window 1 / 4 / 8 / 16), parity tests, bench
mkdir -p gpurun_out/r9
O=gpurun_out/r9
for w in 1 4 8 16; do echo "look-back window $w" >> $O/sortbench.txt; timeout 120 build/sb/sb_lb$w 28 48 >> $O/sortbench.txt 2>&1; timeout 120 build/sb/sb_lb$w 30 48 >> $O/sortbench.txt 2>&1; done
cat $O/sortbench.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_dna30.json 2> $O/bench_dna30.err
head -c 300 $O/bench_dna30.json
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 2 -c 1 -o $O/ncu_onesweep_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_onesweep.log 2>&1

#!/usr/bin/env bash
# gpurun call 8 (1 GPU): sort-pass ranking variants (plain counter update vs leader atomic), re-measure after reverting the value prefetch
mkdir -p gpurun_out/r8
O=gpurun_out/r8
for b in build/sb/sb_base build/sb/sb_atomic; do echo $b >> $O/sortbench.txt; timeout 120 $b 28 48 >> $O/sortbench.txt 2>&1; timeout 120 $b 30 48 >> $O/sortbench.txt 2>&1; done
cat $O/sortbench.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_dna30.json 2> $O/bench_dna30.err
head -c 300 $O/bench_dna30.json

#!/usr/bin/env bash
# gpurun call 5 (1 GPU): merge-tree LPF kernel — parity tests, bench, ncu
mkdir -p gpurun_out/r5
O=gpurun_out/r5
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_dna30.json 2> $O/bench_dna30.err
cat $O/bench_dna30.json
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpf_tile_kernel -s 0 -c 1 -o $O/ncu_lpf_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_lpf.log 2>&1
ls -la $O

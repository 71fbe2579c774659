#!/usr/bin/env bash
# one gpurun call: GPU tests, bench on the three workloads, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_dna30.json 2> gpurun_out/bench_dna30.err
timeout 600 python bench.py --steps 3 --warmup 3 --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > gpurun_out/bench_markov27.json 2> gpurun_out/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > gpurun_out/bench_rep28.json 2> gpurun_out/bench_rep28.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_dna30.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_dna30.json

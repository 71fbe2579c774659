#!/usr/bin/env bash
# gpurun call 13 (1 GPU): state check after re-entry — GPU tests, default bench + reference arm, launch list, ncu full of the sort pass and the LPF kernel
mkdir -p gpurun_out/r13
O=gpurun_out/r13
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
cat $O/bench_default.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
cat $O/bench_reference.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 300 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_dna30.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 2 -c 1 -o $O/ncu_onesweep_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_onesweep.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lpf_tile_kernel -s 0 -c 1 -o $O/ncu_lpf_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_lpf.log 2>&1
for f in $O/bench_*.json; do echo "$f: $(grep '^{' $f | head -c 300)"; done
for f in $O/bench_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -3; done
ls -la $O

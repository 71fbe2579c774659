#!/usr/bin/env bash
# gpurun call 2: GPU tests, bench on three workloads, symbols-per-key sweep, ncu full captures of the top kernels
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_dna30.json 2> $O/bench_dna30.err
cat $O/bench_dna30.json
for k in 15 19 21 25; do
  TDCGPU_SA_SYMBOLS=$k timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_dna30_k$k.json 2> $O/bench_dna30_k$k.err
done
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
for k in 6 8 10 12; do
  TDCGPU_SA_SYMBOLS=$k timeout 300 python bench.py --steps 2 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27_k$k.json 2> $O/bench_markov27_k$k.err
done
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
# ncu --set full of the top kernels at 2^28 (one launch each, after the warm-up launches)
for kn in rs_onesweep_kernel lpf_tile_kernel rerank_apply_kernel scatter_pairs_kernel lcp_fix_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 6 -c 1 -o $O/ncu_$kn -f python bench.py --steps 1 --warmup 1 --log2-bytes 28 --no-cpu-baseline > $O/ncu_$kn.log 2>&1
done
ls -la $O

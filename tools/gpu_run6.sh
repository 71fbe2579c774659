#!/usr/bin/env bash
# gpurun call 6 (1 GPU): predicated merge-tree LPF (64- vs 128-chunk tiles), sampled key-length model
mkdir -p gpurun_out/r6
O=gpurun_out/r6
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_dna30.json 2> $O/bench_dna30.err
cat $O/bench_dna30.json
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
TDCGPU_SA_SYMBOLS=8 timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28_k8.json 2> $O/bench_rep28_k8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpf_tile_kernel -s 0 -c 1 -o $O/ncu_lpf_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_lpf.log 2>&1
# variant: 128 chunks per tile
touch tudocomp_b200/csrc/lzss_kernels.cuh
make -s -C tudocomp_b200/csrc NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v -DLPF_THREADS_CFG=128" > $O/rebuild.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_dna30_lpf128.json 2> $O/bench_dna30_lpf128.err
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27_lpf128.json 2> $O/bench_markov27_lpf128.err
ls -la $O

#!/usr/bin/env bash
# gpurun call 7 (1 GPU): LPF with 16-rank chunks, value prefetch in the sort pass
mkdir -p gpurun_out/r7
O=gpurun_out/r7
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
for b in build/sb/sb_*; do timeout 120 $b 28 48 >> $O/sortbench.txt 2>&1; done
cat $O/sortbench.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_dna30.json 2> $O/bench_dna30.err
cat $O/bench_dna30.json
timeout 400 python bench.py --steps 3 --warmup 3 --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
timeout 400 python bench.py --steps 2 --warmup 3 --workload repetitive --log2-bytes 28 --no-cpu-baseline > $O/bench_rep28.json 2> $O/bench_rep28.err
timeout 600 python bench.py --gpus 1 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_dist1_dna30.json 2> $O/bench_dist1_dna30.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpf_tile_kernel -s 0 -c 1 -o $O/ncu_lpf_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_lpf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 2 -c 1 -o $O/ncu_onesweep_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_onesweep.log 2>&1
# variant: 32-rank chunks, 64 chunks per tile (previous configuration) for an A/B on the same box
touch tudocomp_b200/csrc/lzss_kernels.cuh
make -s -C tudocomp_b200/csrc NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v -DLPF_THREADS_CFG=64 -DLPF_CHUNK_CFG=32" > $O/rebuild.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_dna30_chunk32.json 2> $O/bench_dna30_chunk32.err
ls -la $O

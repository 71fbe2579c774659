#!/usr/bin/env bash
# gpurun call 20 (2 GPUs): bucket partition + push staged through shared memory (whole runs over NVLink)
mkdir -p gpurun_out/r20
O=gpurun_out/r20
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_gpu_dist.log 2>&1
tail -3 $O/pytest_gpu_dist.log
timeout 300 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist2_dna30.json 2> $O/bench_dist2_dna30.err
grep '^{' $O/bench_dist2_dna30.json | head -c 300; echo
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $O/bench_dist2_dna30.err | tail -3

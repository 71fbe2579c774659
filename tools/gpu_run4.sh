#!/usr/bin/env bash
# gpurun call 4 (2 GPUs): full GPU test suite incl. the sharded text index over NCCL, dist-mode bench, block-mode bench
mkdir -p gpurun_out/r4
O=gpurun_out/r4
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 600 $TR bench.py --gpus 2 --mode dist --workload markov --log2-bytes 27 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_dist2_markov27.json 2> $O/bench_dist2_markov27.err
tail -c 600 $O/bench_dist2_markov27.err
timeout 900 $TR bench.py --gpus 2 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_dist2_dna30.json 2> $O/bench_dist2_dna30.err
tail -c 600 $O/bench_dist2_dna30.err
timeout 600 python bench.py --gpus 1 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_dist1_dna30.json 2> $O/bench_dist1_dna30.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_dna30.json 2> $O/bench_dna30.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_block2_dna30.json 2> $O/bench_block2_dna30.err
for f in $O/bench_*.json; do echo "$f: $(head -c 400 $f)"; done

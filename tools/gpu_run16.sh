#!/usr/bin/env bash
# gpurun call 16 (8 GPUs): sharded text index with the fused partition + all-to-all (peer-memory pushes) at 8 ranks
mkdir -p gpurun_out/r16
O=gpurun_out/r16
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_gpu_dist.log 2>&1
tail -4 $O/pytest_gpu_dist.log
timeout 400 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --mode dist --workload dna --bytes 4000000000 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist8_dna4e9.json 2> $O/bench_dist8_dna4e9.err
timeout 300 $TR --nproc-per-node 8 --master-port 29582 bench.py --gpus 8 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist8_dna30.json 2> $O/bench_dist8_dna30.err
timeout 300 $TR --nproc-per-node 4 --master-port 29583 bench.py --gpus 4 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist4_dna30.json 2> $O/bench_dist4_dna30.err
for f in $O/bench_*.json; do echo "$f: $(grep '^{' $f | head -c 260)"; done
for f in $O/bench_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -5; done

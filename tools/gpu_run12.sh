#!/usr/bin/env bash
# gpurun call 12 (2 GPUs): partition kernel pushes straight into the peers' receive buffers (fused partition + all-to-all) vs NCCL
mkdir -p gpurun_out/r12
O=gpurun_out/r12
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_gpu_dist.log 2>&1
tail -4 $O/pytest_gpu_dist.log
timeout 400 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist2_dna30_p2p.json 2> $O/bench_dist2_dna30_p2p.err
TDCGPU_DIST_NO_P2P=1 timeout 400 $TR --nproc-per-node 2 --master-port 29572 bench.py --gpus 2 --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist2_dna30_nccl.json 2> $O/bench_dist2_dna30_nccl.err
timeout 400 $TR --nproc-per-node 2 --master-port 29573 bench.py --gpus 2 --mode dist --workload markov --log2-bytes 27 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist2_markov27_p2p.json 2> $O/bench_dist2_markov27_p2p.err
for f in $O/bench_*.json; do echo "$f: $(grep '^{' $f | head -c 230)"; done
for f in $O/bench_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -5; done

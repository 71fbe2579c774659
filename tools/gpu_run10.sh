#!/usr/bin/env bash
# gpurun call 10 (8 GPUs): sharded text index over NCCL at 2/4/8 ranks (1 GiB and 4 GB texts, sampled verification),
# block mode at 8 ranks
mkdir -p gpurun_out/r10
O=gpurun_out/r10
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_gpu_dist.log 2>&1
tail -4 $O/pytest_gpu_dist.log
timeout 400 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_block8_dna30.json 2> $O/bench_block8_dna30.err
for N in 2 4 8; do
  timeout 400 $TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --mode dist --workload dna --log2-bytes 30 --steps 2 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist${N}_dna30.json 2> $O/bench_dist${N}_dna30.err
done
timeout 600 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --mode dist --workload dna --bytes 4000000000 --steps 1 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist8_dna4e9.json 2> $O/bench_dist8_dna4e9.err
timeout 400 $TR --nproc-per-node 8 --master-port 29582 bench.py --gpus 8 --mode dist --workload repetitive --log2-bytes 28 --steps 1 --warmup 1 --verify --no-cpu-baseline > $O/bench_dist8_rep28.json 2> $O/bench_dist8_rep28.err
for f in $O/bench_*.json; do echo "$f: $(grep '^{' $f | head -c 260)"; done
for f in $O/bench_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -5; done

#!/usr/bin/env bash
# gpurun call 21 (1 GPU, last ~3 GPU-minutes of the round): what changed since calls 17-19 — MTF with per-warp folds, staged push (1 rank), smoke
mkdir -p gpurun_out/r21
O=gpurun_out/r21
( time timeout 90 python -m pytest tests/test_stream_stages.py tests/test_gpu_dist.py -m gpu -x -q ) > $O/pytest_gpu_changed.log 2>&1
tail -4 $O/pytest_gpu_changed.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-160
timeout 60 python tools/chain_bench.py 28 > $O/chain_bench_rep28.txt 2>&1; head -4 $O/chain_bench_rep28.txt

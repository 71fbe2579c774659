#!/usr/bin/env bash
# gpurun call 17 (1 GPU): the whole GPU suite as the driver runs it (incl. lcpcomp + packed arrays), smoke, default bench
mkdir -p gpurun_out/r17
O=gpurun_out/r17
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-200
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
cut -c1-1200 $O/bench_default.json
tail -3 $O/bench_default.err

#!/usr/bin/env bash
# gpurun call 22 (1 GPU, the round's last GPU minutes): the complete GPU suite exactly as the driver runs it
mkdir -p gpurun_out/r22
( time timeout 160 python -m pytest tests -x -q -m gpu ) > gpurun_out/r22/pytest_gpu.log 2>&1
tail -5 gpurun_out/r22/pytest_gpu.log

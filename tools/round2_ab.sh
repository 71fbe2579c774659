#!/usr/bin/env bash
# First GPU call of the next round: A/B of experiments that are compiled out by default.
#   RS_VALS_ASYNC               radix pass: values prefetched to shared memory with cp.async (radix_sort.cuh)
#   TDCGPU_DIST_SLICE_UPLOAD=1  sharded path: every rank uploads n/P bytes, the peers exchange the rest (dist_textds.cu)
# Build here (nvcc cross-compiles):   bash tools/round2_ab.sh build
# Run on the GPU box:                 gpurun -- 'bash tools/round2_ab.sh run'
#                                     gpurun --gpus 2 -- 'bash tools/round2_ab.sh run2'   (e2e of the sharded path, both ways)
#                                     gpurun -- 'bash tools/round2_ab.sh sanitize'        (compute-sanitizer on the new kernels)
#                                     gpurun --gpus 8 -- 'bash tools/round2_ab.sh block 8 8'  (config 5 via tdc_block, 8 GiB)
#                                     gpurun -- 'bash tools/round2_ab.sh ncu'             (ncu --set full of the sort pass)
set -e
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
mkdir -p build/sb
if [ "${1:-build}" = build ]; then
  $NV tools/sortbench.cu -o build/sb/sb_base
  $NV -DRS_VALS_ASYNC tools/sortbench.cu -o build/sb/sb_vals_async
  $NV -DRS_VALS_ASYNC -DRS_IPT64_CFG=12 -DRS_MIN_CTAS_CFG=4 tools/sortbench.cu -o build/sb/sb_vals_async_12x4 || true
  ls -la build/sb
  # lpf_tile_kernel configurations as whole-library variants (tools/lpf_variants.py loads build/variants/libtdcgpu_*.so)
  make -s -C tudocomp_b200/csrc
  mkdir -p build/variants
  for cfg in "t32c16 -DLPF_THREADS_CFG=32" "t128c16 -DLPF_THREADS_CFG=128"; do
    set -- $cfg; name=$1; shift
    $NV -Xcompiler -fPIC "$@" -c tudocomp_b200/csrc/lzss_factorize.cu -o build/variants/lzss_factorize_$name.o
    nvcc -shared -o build/variants/libtdcgpu_$name.so build/tdcgpu_api.o build/suffix_array.o build/lcp.o build/variants/lzss_factorize_$name.o \
      build/lzss_encode.o build/stream_codecs.o build/dist_comm.o build/dist_textds.o -lcudart -ldl
  done
  ls -la build/variants/*.so
elif [ "$1" = sanitize ]; then
  # the kernels added late in round 1 (lzss encoder, packed arrays, mtf / rle / literal encoder, warp-local LPF merges) have
  # not been under compute-sanitizer yet: memcheck + racecheck on small inputs (minutes, not seconds, under the tool)
  mkdir -p gpurun_out/ab
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ab/sanitize_${tool}_smoke.log 2>&1
    tail -3 gpurun_out/ab/sanitize_${tool}_smoke.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_stream_stages.py -m gpu -k golden -x -q > gpurun_out/ab/sanitize_${tool}_stream.log 2>&1
    tail -3 gpurun_out/ab/sanitize_${tool}_stream.log
  done
elif [ "$1" = ncu ]; then
  # fresh `ncu --set full` of the sort pass in its current configuration (profiles/traffic.json still holds the 256 x 12
  # capture).  The micro-benchmark only launches full-size passes, so "-s 3 -c 1" is the 4th pass of the first sort — inside
  # bench.py the first rs_onesweep launches belong to the 2^16-element sample sort (that is what call 13 captured).
  mkdir -p gpurun_out/ab
  for b in sb_base sb_vals_async; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 3 -c 1 -o gpurun_out/ab/ncu_$b -f build/sb/$b 30 48 > gpurun_out/ab/ncu_$b.log 2>&1
    tail -2 gpurun_out/ab/ncu_$b.log
  done
elif [ "$1" = block ]; then
  # config 5 through the real plugin: tdc_block over the GPU registry, 256 MiB blocks, one worker per GPU, with and without the
  # context cache.  gpurun --gpus N -- 'bash tools/round2_ab.sh block N GIB'
  mkdir -p gpurun_out/ab
  N=${2:-1}; GIB=${3:-2}
  python - <<PY
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
with open("/dev/shm/block_in.txt", "wb") as f:
    for i in range($GIB):
        f.write(synth.markov_text(1 << 30, 500 + i)[:-1].tobytes())
PY
  for flag in "" "-c"; do
    ./build/tdc_block_gpu -a "lzss_lcp(coder=huff)" -b 268435456 -g $N $flag /dev/shm/block_in.txt -o /dev/shm/block_out.tdcb 2>&1 | sed "s#^#[-g $N $flag] #" | tee -a gpurun_out/ab/block_mode.txt
  done
  rm -f /dev/shm/block_in.txt /dev/shm/block_out.tdcb
elif [ "$1" = run2 ]; then
  mkdir -p gpurun_out/ab
  TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
  timeout 300 $TR --master-port 29571 bench.py --gpus 2 --mode dist --steps 2 --warmup 1 --verify --no-cpu-baseline > gpurun_out/ab/dist2_full_upload.json 2> gpurun_out/ab/dist2_full_upload.err
  TDCGPU_DIST_SLICE_UPLOAD=1 timeout 300 $TR --master-port 29572 bench.py --gpus 2 --mode dist --steps 2 --warmup 1 --verify --no-cpu-baseline > gpurun_out/ab/dist2_slice_upload.json 2> gpurun_out/ab/dist2_slice_upload.err
  python - <<'PY'
import json
for f in ("full", "slice"):
    d = json.loads([l for l in open(f"gpurun_out/ab/dist2_{f}_upload.json") if l.startswith("{")][0])
    print(f, "resident", round(d["ms_per_step"], 1), "ms  e2e", round(d["e2e"]["ms_per_step"], 1), "ms  verify", d["verify"]["ok"])
PY
else
  mkdir -p gpurun_out/ab
  for b in build/sb/sb_*; do for lg in 28 30; do timeout 120 $b $lg 48 | sed "s#^#$(basename $b): #" | tee -a gpurun_out/ab/sortbench.txt; done; done
  timeout 300 python tools/lpf_variants.py 30 | tee gpurun_out/ab/lpf_variants.txt
fi

#!/usr/bin/env bash
# Round-2 GPU job (one parameterised script instead of round 1's gpu_run1..22.sh): tools/gpu_r2.sh <tag> <stage>...
#   tests     pytest -m gpu (TDC_SIZE_CASES restricts tests/test_gpu_sizes.py)
#   bench     default bench.py line (N=1) -> gpurun_out/<tag>_bench.json
#   launches  ncu launch list of a 2-step bench
# Everything the job wants to keep goes to gpurun_out/.
set -u
cd "$(dirname "$0")/.."
tag="$1"; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
for stage in "$@"; do
  case "$stage" in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_tests.txt ;;
    quicktests)
      timeout 900 python -m pytest tests/test_check.py tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_sizes.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_tests.txt ;;
    bench)
      timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; head -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-dist --no-block-driver --no-pipeline --no-cpu-baseline --no-verify > gpurun_out/${tag}_launches_bench.log 2>&1
      echo "launches rc=$?" ;;
    sortbench)  # A/B of radix-pass variants built by: nvcc ... -D<flag> tools/sortbench.cu -o build/sortbench_<name>
      { for b in build/sortbench_*; do echo "== $b"; timeout 120 $b ${SORT_LG:-28} 48; timeout 120 $b ${SORT_LG:-28} 33 keys; done; } 2>&1 | tee gpurun_out/${tag}_sortbench.txt ;;
    ncusort)  # `ncu --set full` of ONE full-size pair pass (deterministic: third onesweep launch of the micro-benchmark = second plain pass)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 2 -c 1 -o gpurun_out/${tag}_ncu_sortpass -f \
        ${NCU_SORT_BIN:-build/sortbench_b_new} ${NCU_SORT_LG:-30} 48 > gpurun_out/${tag}_ncu_sortpass.log 2>&1
      echo "ncusort rc=$?"; tail -2 gpurun_out/${tag}_ncu_sortpass.log ;;
    benchmodes)  # the initial-sort layouts side by side (resident timing only)
      for mode in wide packed; do
        TDCGPU_SA_MODE=$mode timeout 600 python bench.py --steps 5 --warmup 2 --no-dist --no-block-driver --no-pipeline --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${tag}_bench_$mode.json 2> gpurun_out/${tag}_bench_$mode.err
        echo "$mode rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_$mode.json"))
print("$mode", "ms/step", round(d["ms_per_step"],2), "verified", d["verified"], d["sa_stats"])
print({k:(v["launches"],v["ms"]) for k,v in list(d["kernels"].items())[:14]})
PY
      done ;;
    plugintests)
      timeout 900 python -m pytest tests/test_plugin.py tests/test_encode.py tests/test_check.py tests/test_stream_stages.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_plugintests.txt ;;
    workloads)  # the other single-GPU configs of BASELINE.json: 100 MB Markov (1), 2^30 B repetitive (3); resident + e2e, verified
      timeout 600 python bench.py --steps 5 --warmup 2 --workload markov --bytes 100000000 --no-dist --no-block-driver --no-cpu-baseline > gpurun_out/${tag}_bench_markov1e8.json 2> gpurun_out/${tag}_bench_markov1e8.err
      echo "markov rc=$?"
      timeout 900 python bench.py --steps 3 --warmup 1 --workload repetitive --log2-bytes 30 --no-dist --no-block-driver --no-cpu-baseline --no-pipeline > gpurun_out/${tag}_bench_rep30.json 2> gpurun_out/${tag}_bench_rep30.err
      echo "repetitive rc=$?"
      python - <<PY
import json
for f in ("markov1e8", "rep30"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % f))
        print(f, "ms/step", round(d["ms_per_step"], 2), "MB/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "verified", d["verified"], d["sa_stats"], d["last_step_phases_ms"])
        print({k: (v["launches"], v["ms"]) for k, v in list(d["kernels"].items())[:12]})
    except Exception as e:
        print(f, "failed", e)
PY
      ;;
    sanitize)  # compute-sanitizer memcheck + racecheck over the smoke path, the checkers and the stream stages (small inputs)
      for tool in memcheck racecheck; do
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitize_${tool}_smoke.log 2>&1
        tail -3 gpurun_out/${tag}_sanitize_${tool}_smoke.log
        TDCGPU_SA_MODE=packed timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/${tag}_sanitize_${tool}_cases.log 2>&1
        tail -3 gpurun_out/${tag}_sanitize_${tool}_cases.log
      done ;;
    block)  # BASELINE config 5 through the real plugin: tdc_block over the GPU registry, 256 MiB blocks, one worker per GPU
      N=${BLOCK_GPUS:-1}; GIB=${BLOCK_GIB:-2}
      python - <<PY
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
with open("/dev/shm/block_in.txt", "wb") as f:  # Markov text generated on the GPU, 1 GiB at a time (numpy: 2 min per GiB)
    for i in range($GIB):
        f.write(synth.markov_text_device(1 << 30, 500 + i).cpu().numpy().tobytes())
PY
      ./build/tdc_block_gpu -a "lzss_lcp(coder=huff)" -b 268435456 -g $N /dev/shm/block_in.txt -o /dev/shm/block_out.tdcb 2>&1 | sed "s#^#[-g $N] #" | tee -a gpurun_out/${tag}_block_mode.txt
      ls -l /dev/shm/block_in.txt /dev/shm/block_out.tdcb | tee -a gpurun_out/${tag}_block_mode.txt
      # the container is decoded by the REFERENCE registry (tdc_block_ref -d) and compared with the input
      if [ -z "${BLOCK_SKIP_RT:-}" ]; then
        ( time ./build/tdc_block_ref -d /dev/shm/block_out.tdcb -o /dev/shm/block_back.txt ) 2>&1 | tail -4 | tee -a gpurun_out/${tag}_block_mode.txt
        cmp /dev/shm/block_in.txt /dev/shm/block_back.txt && echo "round trip through the reference decoder: identical" | tee -a gpurun_out/${tag}_block_mode.txt
      fi
      rm -f /dev/shm/block_in.txt /dev/shm/block_out.tdcb /dev/shm/block_back.txt ;;
    chain)  # BASELINE config 3 through the stock driver + GPU-only registry: bwt:mtf:rle:encode(huff), device-resident chain vs host chain
      python - <<PY
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
for lg in (25, 28, 30):
    open("/dev/shm/rep%d.txt" % lg, "wb").write(synth.repetitive(1 << lg, 3)[:-1].tobytes())
PY
      {
        A="bwt:mtf:rle:encode(huff)"
        ./build/tdc_ref -a "$A" /dev/shm/rep25.txt -o /dev/shm/rep25.ref --force
        ./build/tdc_gpu_only -a "$A" /dev/shm/rep25.txt -o /dev/shm/rep25.gpu --force
        TDCGPU_HOST_CHAIN=1 ./build/tdc_gpu_only -a "$A" /dev/shm/rep25.txt -o /dev/shm/rep25.gpuh --force
        cmp /dev/shm/rep25.ref /dev/shm/rep25.gpu && cmp /dev/shm/rep25.ref /dev/shm/rep25.gpuh && echo "2^25 B chain archives identical (reference driver, device-resident chain, host chain): $(stat -c %s /dev/shm/rep25.ref) bytes"
        for lg in 28 30; do
          for mode in 0 1; do
            t0=$(date +%s%N)
            TDCGPU_HOST_CHAIN=$mode ./build/tdc_gpu_only -a "$A" /dev/shm/rep$lg.txt -o /dev/shm/rep$lg.gpu$mode --force --stats > gpurun_out/${tag}_chain_stats_${lg}_$mode.json
            echo "$(( ($(date +%s%N) - t0) / 1000000 )) ms wall  <- tdc_gpu_only $A, 2^$lg B repetitive, TDCGPU_HOST_CHAIN=$mode"
          done
          cmp /dev/shm/rep$lg.gpu0 /dev/shm/rep$lg.gpu1 && echo "2^$lg B: device-resident and host chain archives identical: $(stat -c %s /dev/shm/rep$lg.gpu0) bytes"
        done
        rm -f /dev/shm/rep2*.txt /dev/shm/rep3*.txt /dev/shm/rep*.gpu* /dev/shm/rep*.ref
      } 2>&1 | tee gpurun_out/${tag}_chain.txt ;;
    ncu)  # one `ncu --set full` capture of the dominant kernel inside the real bench (full-size launch: skip the sample sort)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-rs_onesweep} -s ${NCU_SKIP:-3} -c 1 -o gpurun_out/${tag}_ncu -f \
        python bench.py --steps 1 --warmup 1 --no-dist --no-block-driver --no-pipeline --no-cpu-baseline --no-verify > gpurun_out/${tag}_ncu.log 2>&1
      echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log ;;
    distab)  # sharded path A/B on all GPUs of the box: every rank uploads the whole text vs its own slice (+ peer exchange)
      N=$(nvidia-smi -L | wc -l)
      for su in ${SLICE_MODES:-0 1}; do
        TDCGPU_DIST_SLICE_UPLOAD=$su timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$su bench.py --gpus $N --mode dist --bytes ${DIST_BYTES:-2000000000} --steps 3 --warmup 2 > gpurun_out/${tag}_dist_n${N}_slice$su.json 2> gpurun_out/${tag}_dist_n${N}_slice$su.err
        python - <<PY
import json
for ln in open("gpurun_out/${tag}_dist_n${N}_slice$su.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("slice_upload=$su N=$N", d["config"]["workload"], "ms/step", round(d["ms_per_step"], 2), "MB/s", round(d["value"]), "e2e MB/s", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "verified", (d.get("verify") or {}).get("ok"))
PY
      done 2>&1 | tee gpurun_out/${tag}_distab.txt ;;
    dist)  # the default bench line on all GPUs of the box (block-mode headline + sharded sub-records)
      N=$(nvidia-smi -L | wc -l)
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
      echo "dist bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_n$N.json; tail -3 gpurun_out/${tag}_bench_n$N.err ;;
    *) echo "unknown stage $stage" ;;
  esac
done

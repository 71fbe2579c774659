#!/usr/bin/env bash
# gpurun call 14 (1 GPU): device-side lzss::encode_text — GPU parity tests, smoke, bench with the archive leg, tdc driver A/B
mkdir -p gpurun_out/r14
O=gpurun_out/r14
( time timeout 900 python -m pytest tests/test_encode.py tests/test_plugin.py -m gpu -x -q ) > $O/pytest_gpu_encode.log 2>&1
tail -5 $O/pytest_gpu_encode.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_dna30.json 2> $O/bench_dna30.err
timeout 300 python bench.py --workload markov --log2-bytes 27 --no-cpu-baseline > $O/bench_markov27.json 2> $O/bench_markov27.err
python - <<'PY' > gpurun_out/r14/archive_summary.txt 2>&1
import json
for f in ("dna30", "markov27"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r14/bench_{f}.json") if l.startswith("{")][0])
        print(f, "step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "archive", json.dumps(d["archive"]))
    except Exception as e:
        print(f, "failed", e)
PY
cat $O/archive_summary.txt
# the stock tdc driver, 256 MiB of markov text: device encode vs host encode_text, --stats phases
python - <<'PY'
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
open("/tmp/markov256.txt", "wb").write(synth.markov_text(1 << 28, 77)[:-1].tobytes())
open("/tmp/markov32.txt", "wb").write(synth.markov_text(1 << 25, 78)[:-1].tobytes())
PY
for mode in dev host; do
  if [ $mode = host ]; then export TDCGPU_HOST_ENCODE=1; else unset TDCGPU_HOST_ENCODE; fi
  /usr/bin/time -v ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.$mode.tdc --force --stats > $O/tdc_markov256_$mode.stats 2> $O/tdc_markov256_$mode.time
  grep -E "Elapsed|Maximum resident" $O/tdc_markov256_$mode.time
done
unset TDCGPU_HOST_ENCODE
cmp /tmp/m256.dev.tdc /tmp/m256.host.tdc && echo "256 MiB archives identical: $(stat -c %s /tmp/m256.dev.tdc) bytes"
/usr/bin/time -v ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.gpu.tdc --force > /dev/null 2> $O/tdc_markov32_gpu.time
/usr/bin/time -v ./build/tdc_ref -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.ref.tdc --force --stats > $O/tdc_markov32_ref.stats 2> $O/tdc_markov32_ref.time
grep -E "Elapsed" $O/tdc_markov32_gpu.time $O/tdc_markov32_ref.time
cmp /tmp/m32.gpu.tdc /tmp/m32.ref.tdc && echo "32 MiB archive identical to the reference driver's"
# ncu: launch list of one archive step + full capture of the write kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_tile_kernel -s 2 -c 1 -o $O/ncu_enc_write_dna30 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_enc.log 2>&1
ls -la $O

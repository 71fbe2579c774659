#!/usr/bin/env bash
# gpurun call 18 (1 GPU): stream stages behind the BWT (mtf / rle / encode) — GPU parity tests, plugin tests, stage timings
mkdir -p gpurun_out/r18
O=gpurun_out/r18
( time timeout 600 python -m pytest tests/test_stream_stages.py tests/test_plugin.py -m gpu -x -q ) > $O/pytest_gpu_stream.log 2>&1
tail -5 $O/pytest_gpu_stream.log
timeout 400 python tools/chain_bench.py 28 > $O/chain_bench_rep28.txt 2>&1; cat $O/chain_bench_rep28.txt
python - <<'PY'
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
open("/tmp/rep256.txt", "wb").write(synth.repetitive(1 << 28, 3)[:-1].tobytes())
open("/tmp/rep32.txt", "wb").write(synth.repetitive(1 << 25, 3)[:-1].tobytes())
PY
TIMEFORMAT='%R s wall, %U s user'
{ time timeout 300 ./build/tdc_gpu_only -a "bwt:mtf:rle:encode(huff)" /tmp/rep256.txt -o /tmp/r256.gpu.tdc --force ; } 2>> $O/tdc_chain_times.txt
echo "  ^ tdc_gpu_only bwt:mtf:rle:encode(huff), 256 MiB repetitive text: $(stat -c %s /tmp/r256.gpu.tdc) bytes out" >> $O/tdc_chain_times.txt
{ time timeout 300 ./build/tdc_gpu_only -a "bwt:mtf:rle:encode(huff)" /tmp/rep32.txt -o /tmp/r32.gpu.tdc --force ; } 2>> $O/tdc_chain_times.txt
echo "  ^ tdc_gpu_only, 32 MiB" >> $O/tdc_chain_times.txt
{ time timeout 300 ./build/tdc_ref -a "bwt:mtf:rle:encode(huff)" /tmp/rep32.txt -o /tmp/r32.ref.tdc --force ; } 2>> $O/tdc_chain_times.txt
echo "  ^ tdc_ref (unmodified reference, 1 core), 32 MiB" >> $O/tdc_chain_times.txt
cmp /tmp/r32.gpu.tdc /tmp/r32.ref.tdc && echo "32 MiB archives identical" >> $O/tdc_chain_times.txt
cat $O/tdc_chain_times.txt

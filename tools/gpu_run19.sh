#!/usr/bin/env bash
# gpurun call 19 (1 GPU): tdc driver wall times again after the glog-shim fix (DLOG operands no longer evaluated): chain of config 3, lzss_lcp host/device encode A/B
mkdir -p gpurun_out/r19
O=gpurun_out/r19
python - <<'PY'
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
open("/tmp/rep256.txt", "wb").write(synth.repetitive(1 << 28, 3)[:-1].tobytes())
open("/tmp/rep32.txt", "wb").write(synth.repetitive(1 << 25, 3)[:-1].tobytes())
open("/tmp/markov256.txt", "wb").write(synth.markov_text(1 << 28, 77)[:-1].tobytes())
open("/tmp/markov32.txt", "wb").write(synth.markov_text(1 << 25, 78)[:-1].tobytes())
PY
TIMEFORMAT='%R s wall, %U s user'
T=$O/tdc_times.txt
run() { label="$1"; shift; { time timeout 300 "$@" > /dev/null ; } 2>> $T; echo "  ^ $label" >> $T; }
run "tdc_gpu_only bwt:mtf:rle:encode(huff), 256 MiB repetitive (warm-up of the page cache)" ./build/tdc_gpu_only -a "bwt:mtf:rle:encode(huff)" /tmp/rep256.txt -o /tmp/r256.gpu.tdc --force
run "tdc_gpu_only bwt:mtf:rle:encode(huff), 256 MiB repetitive" ./build/tdc_gpu_only -a "bwt:mtf:rle:encode(huff)" /tmp/rep256.txt -o /tmp/r256.gpu.tdc --force
run "tdc_gpu_only bwt:mtf:rle:encode(huff), 32 MiB repetitive" ./build/tdc_gpu_only -a "bwt:mtf:rle:encode(huff)" /tmp/rep32.txt -o /tmp/r32.gpu.tdc --force
run "tdc_ref (unmodified reference, 1 core) bwt:mtf:rle:encode(huff), 32 MiB repetitive" ./build/tdc_ref -a "bwt:mtf:rle:encode(huff)" /tmp/rep32.txt -o /tmp/r32.ref.tdc --force
cmp /tmp/r32.gpu.tdc /tmp/r32.ref.tdc && echo "32 MiB chain archives identical ($(stat -c %s /tmp/r32.gpu.tdc) bytes)" >> $T
run "tdc_gpu_only lzss_lcp(coder=huff), 256 MiB markov, device encode" ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.dev.tdc --force
./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.dev.tdc --force --stats > $O/tdc_markov256_dev.stats 2>/dev/null
export TDCGPU_HOST_ENCODE=1
run "tdc_gpu_only lzss_lcp(coder=huff), 256 MiB markov, TDCGPU_HOST_ENCODE=1" ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.host.tdc --force
./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.host.tdc --force --stats > $O/tdc_markov256_host.stats 2>/dev/null
unset TDCGPU_HOST_ENCODE
cmp /tmp/m256.dev.tdc /tmp/m256.host.tdc && echo "256 MiB lzss archives identical" >> $T
run "tdc_gpu_only lzss_lcp(coder=huff), 32 MiB markov" ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.gpu.tdc --force
run "tdc_ref lzss_lcp(coder=huff), 32 MiB markov" ./build/tdc_ref -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.ref.tdc --force
cmp /tmp/m32.gpu.tdc /tmp/m32.ref.tdc && echo "32 MiB lzss archives identical" >> $T
cat $T

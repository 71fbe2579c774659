#!/usr/bin/env bash
# gpurun call 15 (1 GPU): lpf_tile_kernel with warp-local merge levels (+ configuration variants), tdc driver A/B for the device encoder
mkdir -p gpurun_out/r15
O=gpurun_out/r15
timeout 600 python tools/lpf_variants.py 30 > $O/lpf_variants.txt 2>&1; cat $O/lpf_variants.txt
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > $O/pytest_gpu_parity.log 2>&1; tail -4 $O/pytest_gpu_parity.log
python - <<'PY'
import sys; sys.path.insert(0, ".")
from tudocomp_b200 import synth
open("/tmp/markov256.txt", "wb").write(synth.markov_text(1 << 28, 77)[:-1].tobytes())
open("/tmp/markov32.txt", "wb").write(synth.markov_text(1 << 25, 78)[:-1].tobytes())
PY
TIMEFORMAT='%R s wall, %U s user'
for mode in dev host; do
  if [ $mode = host ]; then export TDCGPU_HOST_ENCODE=1; else unset TDCGPU_HOST_ENCODE; fi
  for rep in 1 2; do
    { time ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov256.txt -o /tmp/m256.$mode.tdc --force --stats > $O/tdc_markov256_$mode.stats ; } 2>> $O/tdc_times.txt
    echo "  ^ tdc_gpu_only lzss_lcp(huff) 256 MiB markov, encode on $mode (run $rep)" >> $O/tdc_times.txt
  done
done
unset TDCGPU_HOST_ENCODE
cmp /tmp/m256.dev.tdc /tmp/m256.host.tdc && echo "256 MiB archives identical: $(stat -c %s /tmp/m256.dev.tdc) bytes" >> $O/tdc_times.txt
{ time ./build/tdc_gpu_only -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.gpu.tdc --force > /dev/null ; } 2>> $O/tdc_times.txt
echo "  ^ tdc_gpu_only 32 MiB markov" >> $O/tdc_times.txt
{ time ./build/tdc_ref -a "lzss_lcp(coder=huff)" /tmp/markov32.txt -o /tmp/m32.ref.tdc --force --stats > $O/tdc_markov32_ref.stats ; } 2>> $O/tdc_times.txt
echo "  ^ tdc_ref (unmodified reference, 1 core) 32 MiB markov" >> $O/tdc_times.txt
cmp /tmp/m32.gpu.tdc /tmp/m32.ref.tdc && echo "32 MiB archive identical to the reference driver's" >> $O/tdc_times.txt
cat $O/tdc_times.txt

"""Config 3 of BASELINE.json on the GPU box: bwt -> mtf -> rle -> encode(bit) with device-resident hand-over inside one
context is not what the tdc chain does (it hands host buffers from stage to stage), so both are timed:
  * per-stage kernels (CUDA events via tdcgpu_profile_*) with host buffers through the C ABI, as the plugin calls them;
  * the reference's CPU stages (oracle/_ref) on a 16 MiB sample of the same data, single thread.
Usage: python tools/chain_bench.py [log2_bytes=28]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import tudocomp_b200 as tdc  # noqa: E402
from tudocomp_b200 import synth  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
lib = tdc.load()
t = synth.repetitive(1 << lg, 3)
codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)
with tdc.Context(lib, 0) as c:
    for rep in range(2):
        lib.profile_reset()
        lib.profile_enable(True)
        w0 = time.perf_counter()
        c.set_text(t)
        c.build(tdc.SA | tdc.BWT)
        b = c.get(tdc.BWT)
        w1 = time.perf_counter()
        m = c.mtf_encode(b)
        w2 = time.perf_counter()
        r = c.rle_encode(m, 0)
        w3 = time.perf_counter()
        c.literal_histogram_of(r)
        e = c.literal_encode(codes, lens)
        w4 = time.perf_counter()
        c.sync()
        lib.profile_enable(False)
        prof = lib.profile()
    n = t.size
    print(f"repetitive 2^{lg} B: bwt(SA+gather+D2H) {1e3 * (w1 - w0):.1f} ms, mtf {1e3 * (w2 - w1):.1f} ms, rle {1e3 * (w3 - w2):.1f} ms "
          f"({r.size} B out), encode(bit) {1e3 * (w4 - w3):.1f} ms ({e.size} B out); host buffers, wall clock incl. PCIe")
    stage = {"mtf": ("mtf_summaries", "mtf_scan_kernel", "mtf_apply"), "rle": ("rle_first_head_kernel", "rle_next_head_kernel", "rle_count", "rle_scan", "rle_write"),
             "encode": ("stream_histogram_kernel", "lit_count", "lit_scan", "lit_header_kernel", "lit_write")}
    for s, ks in stage.items():
        ms = {k: round(prof[k]["ms"], 3) for k in ks if k in prof}
        tot = sum(ms.values())
        print(f"  {s}: kernels {tot:.3f} ms = {n / 1e6 / tot:.0f} MB/ms-input..." if False else f"  {s}: kernels {tot:.3f} ms ({n / 1e9 / (tot / 1e3):.1f} GB/s of stage input) {ms}")
# CPU reference on a 16 MiB sample of the same BWT
from conftest import Reference  # noqa: E402

ref = Reference()
sb = b[: 1 << 24]
t0 = time.perf_counter(); sm = ref.stream_stage(0, sb); t1 = time.perf_counter(); sr = ref.stream_stage(1, sm, 0); t2 = time.perf_counter()
ref.stream_stage(2, sr); t3 = time.perf_counter()
print(f"reference CPU stages on the first 2^24 B of the same BWT (1 core): mtf {sb.size / 1e6 / (t1 - t0):.1f} MB/s, rle {sm.size / 1e6 / (t2 - t1):.1f} MB/s, "
      f"encode(bit) {sr.size / 1e6 / (t3 - t2):.1f} MB/s")

"""A/B of lpf_tile_kernel configurations on the GPU box: the same text through variant builds of the library
(build/variants/libtdcgpu_<name>.so, made by hand with -DLPF_THREADS_CFG / -DLPF_CHUNK_CFG), per-kernel CUDA-event times."""
import glob
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tudocomp_b200 import _abi, synth  # noqa: E402

workloads = [("dna", synth.dna(1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28, 2)),
             ("markov", synth.markov_text(1 << 27, 1))]
libs = [("default", os.path.join(ROOT, "tudocomp_b200", "libtdcgpu.so"))] + \
       [(os.path.basename(p)[len("libtdcgpu_"):-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libtdcgpu_*.so")))]
for wname, t in workloads:
    ref_sum = None
    for name, path in libs:
        lib = _abi.TdcGpuLib(path)
        with _abi.Context(lib, 0) as c:
            c.set_text(t)
            c.build(_abi.SA | _abi.ISA | _abi.LCP)
            c.factorize(3)  # warm-up
            lib.profile_reset()
            lib.profile_enable(True)
            for _ in range(3):
                z, mn, mx = c.factorize(3)
            c.sync()
            lib.profile_enable(False)
            prof = lib.profile()
            f = c.factors(z)
            chk = zlib.crc32(f.tobytes())
            ref_sum = ref_sum if ref_sum is not None else chk
            print(f"{wname} n={t.size} {name:10s} lpf_tile_kernel {prof['lpf_tile_kernel']['ms'] / 3:8.3f} ms  factorize-kernels "
                  f"{sum(v['ms'] for v in prof.values()) / 3:8.3f} ms  z={z} crc_ok={chk == ref_sum}", flush=True)

"""A/B of compile-time kernel configurations on the GPU box: the same text through variant builds of the library
(build/variants/libtdcgpu_<name>.so, built with -D<MACRO>_CFG=...), per-kernel CUDA-event times of build + factorize.
Usage: python tools/variant_bench.py [log2_bytes=28]"""
import glob
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tudocomp_b200 import _abi, synth  # noqa: E402

t = synth.dna(1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28), 2)
libs = [("default", os.path.join(ROOT, "tudocomp_b200", "libtdcgpu.so"))] + \
       [(os.path.basename(p)[len("libtdcgpu_"):-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libtdcgpu_*.so")))]
ref = None
for name, path in libs:
    lib = _abi.TdcGpuLib(path)
    with _abi.Context(lib, 0) as c:
        c.set_text(t)
        c.build(_abi.SA | _abi.ISA | _abi.LCP)
        c.factorize(3)
        lib.profile_reset()
        lib.profile_enable(True)
        for _ in range(3):
            c.set_text(t)
            c.build(_abi.SA | _abi.ISA | _abi.LCP)
            z, mn, mx = c.factorize(3)
        c.sync()
        lib.profile_enable(False)
        prof = lib.profile()
        chk = zlib.crc32(c.factors(z).tobytes())
        ref = ref if ref is not None else chk
        tot = sum(v["ms"] for v in prof.values()) / 3
        top = {k: round(v["ms"] / 3, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}
        print(f"{name:12s} kernels {tot:8.3f} ms  crc_ok={chk == ref}  {top}", flush=True)

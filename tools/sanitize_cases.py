"""Small inputs through every kernel family, meant to run under compute-sanitizer (tools/gpu_r2.sh sanitize):
index (both record layouts via TDCGPU_SA_MODE), Phi/PLCP route, factoriser, device encoder, stream stages, checkers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tudocomp_b200 as tdc  # noqa: E402
from tudocomp_b200 import synth  # noqa: E402

lib = tdc.load()
ALL = tdc.SA | tdc.ISA | tdc.LCP | tdc.PHI | tdc.PLCP | tdc.BWT
for name, t in (("dna", synth.dna(200000, 1)), ("markov", synth.markov_text(150000, 2)),
                ("repetitive", synth.repetitive(120000, 3, block=3000, p=0.01)), ("tailrun", synth.with_sentinel(np.full(5000, 65, np.uint8)))):
    with tdc.Context(lib, 0) as c:
        c.set_text(t)
        c.build(tdc.SA | tdc.LCP)          # direct LCP route (seeded)
        c.set_text(t)
        c.build(ALL)                       # Phi / PLCP route
        z, mn, mx = c.factorize(3)
        res = c.check(3, z)
        assert res["ok"], (name, res)
        hist, _ = c.literal_histogram()
        nbits = c.encode(np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8))
        arc = c.encoded_chunks(1 << 16)
        bwt = c.get(tdc.BWT)
        m = c.mtf_encode(bwt)
        r = c.rle_encode(m)
        c.literal_histogram_of(r)
        c.literal_encode(np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8))
        packed = c.get_packed(tdc.SA, 20)
        print(name, "ok", t.size, z, nbits, arc.size, r.size, c.sa_stats())
print("sanitize cases done")

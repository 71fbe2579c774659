"""One text of 2^31 bytes or more on ONE GPU (wide indices): 32-bit unsigned indices reach 2^32, the
reference's default build stops at 2^31 (divsufsort's signed BufferWrapper) and needs -DLEN_BITS=40 beyond.  The text is
generated on the device; every slot / position is checked by the device checkers (csrc/check.cu).
Usage: python tools/wide_run.py [n_body=2281701376] [workload=dna|markov]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tudocomp_b200 as tdc  # noqa: E402
from tudocomp_b200 import synth  # noqa: E402

n_body = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 31) + (1 << 27)
workload = sys.argv[2] if len(sys.argv) > 2 else "dna"
lib = tdc.load()
dev = "cuda:0"
if workload == "markov":
    body = synth.markov_text_device(n_body, 7, dev)
else:
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    body = torch.empty(n_body, dtype=torch.uint8, device=dev)
    for lo in range(0, n_body, 1 << 27):
        hi = min(n_body, lo + (1 << 27))
        body[lo:hi] = lut[torch.randint(0, 4, (hi - lo,), generator=g, device=dev, dtype=torch.int64)]
text = torch.zeros(n_body + 1, dtype=torch.uint8, device=dev)
text[:n_body] = body
del body
torch.cuda.synchronize()
n = n_body + 1
out = {"workload": workload, "n": n, "index_bits": 32}
with tdc.Context(lib, 0) as c:
    for rep in range(2):
        t0 = time.perf_counter()
        c.set_text_device(text.data_ptr(), n)
        c.build(tdc.SA | tdc.ISA | tdc.LCP)
        z, mn, mx = c.factorize(3)
        c.sync()
        out["ms_step_%d" % rep] = round(1e3 * (time.perf_counter() - t0), 2)
    out.update(factors=int(z), factor_len=[int(mn), int(mx)], max_lcp=int(c.max_lcp()), sa_stats=c.sa_stats())
    out["MBps"] = round(n_body / 1e6 / (out["ms_step_1"] / 1e3), 1)
    t0 = time.perf_counter()
    out["verify"] = c.check(3, int(z))
    out["verify_s"] = round(time.perf_counter() - t0, 2)
print(json.dumps(out))
sys.exit(0 if out["verify"].get("ok") else 1)

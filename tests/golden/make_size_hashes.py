"""SHA-256 fingerprints of the UNMODIFIED reference's results at the sizes BASELINE.json's configs name.

Run in the build container only (needs /root/reference through oracle/_ref/libtdcref.so):

    python tests/golden/make_size_hashes.py --case dna_2p30 --job index      # one job, one process (~15-25 min at 2^30)
    python tests/golden/make_size_hashes.py --all [--parallel 2]             # every case x job, then merge
    python tests/golden/make_size_hashes.py --merge                          # fragments -> tests/golden/size_hashes.json

Jobs per case (each is one run of the reference, nothing is derived on our side):
  index    TextDS<> with SA|ISA|LCP (SADivSufSort, ISAFromSA, PhiFromSA -> PLCPFromPhi -> LCPFromPLCP) + bwt::bwt:
           sha256 of the little-endian u32 arrays SA, ISA, LCP, of the n BWT bytes, and max_lcp
  factors  LZSSLCPCompressor<spy coder>(threshold=3): sha256 of the (pos,src,len) u32 triples, z, flen_min/max, fdist_max
  bit/huff LZSSLCPCompressor<BitCoder|HuffmanCoder>(threshold=3) raw archive: sha256 + length
The committed JSON is what `tests/test_gpu_sizes.py` and `bench.py` (verified: true) compare the device results with on
the GPU box, where neither /root/reference nor 15 CPU-minutes per array exist.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
FRAG = os.path.join(HERE, "size_hashes")
THRESHOLD = 3

# name -> (generator, body bytes, seed); the 2^30 / 1e8 cases are exactly bench.py's rank-0 texts for the three workloads
CASES = {
    "markov_1e8": ("markov", 100_000_000, 1),     # BASELINE config 1
    "dna_2p30": ("dna", 1 << 30, 2),              # config 2 (and the bench headline)
    "repetitive_2p30": ("repetitive", 1 << 30, 3),  # config 3
    "dna_2p26": ("dna", 1 << 26, 2),              # mid sizes: above the 2^22-element partitioned-scatter switch,
    "markov_2p26": ("markov", 1 << 26, 1),        # cheap enough for every GPU test run
    "repetitive_2p26": ("repetitive", 1 << 26, 3),
    "dna_2p28": ("dna", 1 << 28, 2),              # config 5's block size
    "markov_2p28": ("markov", 1 << 28, 5),
}
JOBS = ("index", "factors", "bit", "huff")


def gen(case):
    from tudocomp_b200 import synth

    kind, n_body, seed = CASES[case]
    return {"dna": synth.dna, "markov": synth.markov_text, "repetitive": synth.repetitive}[kind](n_body, seed)


def sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


def run_job(case, job):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtdcref.so"))
    lib.tdcref_lzss_lcp_factors.restype = ctypes.c_int64
    lib.tdcref_lzss_lcp_compress.restype = ctypes.c_int64
    lib.tdcref_last_error.restype = ctypes.c_char_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    t = gen(case)
    n = int(t.size)
    out = {"case": case, "job": job, "n": n, "generator": list(CASES[case]), "text_sha256": sha(t)}
    t0 = time.time()
    if job == "index":
        sa, isa, lcp = (np.empty(n, np.uint32) for _ in range(3))
        bwt = np.empty(n, np.uint8)
        mx = ctypes.c_uint32(0)
        rc = lib.tdcref_index_bwt(P(t), ctypes.c_uint64(n), P(sa), P(isa), P(lcp), P(bwt), ctypes.byref(mx))
        assert rc == 0, lib.tdcref_last_error()
        out.update(sa=sha(sa), isa=sha(isa), lcp=sha(lcp), bwt=sha(bwt), max_lcp=int(mx.value),
                   sa_head=[int(x) for x in sa[:4]], lcp_sum=int(lcp.sum(dtype=np.uint64)))
    elif job == "factors":
        hdr = (ctypes.c_uint64 * 3)()
        cap = n // 2 + 16
        f = np.empty((cap, 3), np.uint32)
        z = lib.tdcref_lzss_lcp_factors(P(t), ctypes.c_uint64(n), ctypes.c_uint32(THRESHOLD), P(f), ctypes.c_uint64(cap), hdr)
        assert z >= 0, lib.tdcref_last_error()
        f = f[:z]
        out.update(threshold=THRESHOLD, z=int(z), factors=sha(f), flen_min=int(hdr[0]), flen_max=int(hdr[1]), fdist_max=int(hdr[2]),
                   pos=sha(f[:, 0]), src=sha(f[:, 1]), len=sha(f[:, 2]))
    else:
        coder = {"bit": 0, "huff": 1}[job]
        cap = n + n // 4 + 4096
        arc = np.empty(cap, np.uint8)
        secs = ctypes.c_double(0)
        ln = lib.tdcref_lzss_lcp_compress(P(t), ctypes.c_uint64(n), ctypes.c_uint32(THRESHOLD), coder, P(arc), ctypes.c_uint64(cap), ctypes.byref(secs))
        assert 0 <= ln <= cap, lib.tdcref_last_error()
        out.update(threshold=THRESHOLD, archive_len=int(ln), archive=sha(arc[:ln]), reference_compress_s=secs.value)
    out["wall_s"] = round(time.time() - t0, 1)
    os.makedirs(FRAG, exist_ok=True)
    with open(os.path.join(FRAG, f"{case}.{job}.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(out))


def merge():
    res = {}
    for fn in sorted(os.listdir(FRAG)):
        if not fn.endswith(".json"):
            continue
        d = json.load(open(os.path.join(FRAG, fn)))
        c = res.setdefault(d["case"], {"n": d["n"], "generator": d["generator"], "text_sha256": d["text_sha256"]})
        assert c["text_sha256"] == d["text_sha256"]
        c[d["job"]] = {k: v for k, v in d.items() if k not in ("case", "job", "n", "generator", "text_sha256")}
    with open(os.path.join(HERE, "size_hashes.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print("merged", {k: sorted(j for j in v if j in JOBS) for k, v in res.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", choices=sorted(CASES))
    ap.add_argument("--job", choices=JOBS)
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--cases", default="", help="comma list for --all (default: every case)")
    ap.add_argument("--parallel", type=int, default=2, help="--all: reference processes at once (a 2^30 job needs 17-27 GB)")
    ap.add_argument("--merge", action="store_true")
    args = ap.parse_args()
    if args.all:
        cases = [c for c in args.cases.split(",") if c] or list(CASES)
        todo = [(c, j) for c in cases for j in JOBS if not os.path.exists(os.path.join(FRAG, f"{c}.{j}.json"))]
        running = []
        while todo or running:
            running = [p for p in running if p.poll() is None]
            while todo and len(running) < args.parallel:
                c, j = todo.pop(0)
                running.append(subprocess.Popen([sys.executable, __file__, "--case", c, "--job", j], stdout=subprocess.DEVNULL))
            time.sleep(2)
        merge()
    elif args.merge:
        merge()
    else:
        run_job(args.case, args.job)


if __name__ == "__main__":
    main()

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libtdcref.so, built by oracle/Makefile from
/root/reference).  Run in the build container only: `python tests/golden/make_golden.py`.  The fixtures are committed so
that the oracle port and the CUDA path can be pinned on the GPU box, where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import Reference  # noqa: E402
from inputs import all_small_cases  # noqa: E402
from tudocomp_b200 import synth  # noqa: E402


def main():
    ref = Reference()
    out = {}
    names = []
    cases = list(all_small_cases())
    # two mid-size seeded cases per BASELINE workload family
    cases += [("dna_64k", synth.dna(65536, 2)), ("markov_64k", synth.markov_text(65536, 1)),
              ("repetitive_64k", synth.repetitive(65536, 3, block=4096, p=0.01))]
    for name, t in cases:
        ds = ref.textds(t)
        small = t.size <= 70000
        names.append(name)
        out[f"{name}/text"] = t
        for thr in (1, 2, 3, 5):
            f, hdr = ref.factors(t, thr)
            out[f"{name}/factors{thr}"] = f
            out[f"{name}/hdr{thr}"] = np.array(hdr, np.uint64)
        # the coder side of the archives below, from the reference's own HuffmanCoder: literal histogram of the
        # threshold-3 parse (counted on the reference's factor list), header bits and code words
        f3 = out[f"{name}/factors3"]
        covered = np.zeros(t.size + 1, np.int64)
        np.add.at(covered, f3[:, 0].astype(np.int64), 1)
        np.add.at(covered, (f3[:, 0].astype(np.int64) + f3[:, 2]), -1)
        lit = np.cumsum(covered[:-1]) == 0
        hist = np.bincount(t[lit], minlength=256).astype(np.uint64)
        out[f"{name}/lit_hist3"] = hist
        head, hbits, codes, lens = ref.literal_coder(1, hist)
        out[f"{name}/huff_head"] = head
        out[f"{name}/huff_head_bits"] = np.array([hbits], np.uint64)
        out[f"{name}/huff_codes"] = codes
        out[f"{name}/huff_lens"] = lens
        for coder, cname in ((0, "bit"), (1, "huff"), (2, "ascii")):
            arc, _ = ref.compress(t, 3, coder)
            out[f"{name}/arc_{cname}"] = arc if small else np.frombuffer(__import__("hashlib").sha256(arc.tobytes()).digest(), np.uint8)
        out[f"{name}/bwt"] = ref.bwt(t)
        out[f"{name}/max_lcp"] = np.array([ds["max_lcp"]], np.uint32)
        for k in ("sa", "isa", "lcp", "phi", "plcp"):
            out[f"{name}/{k}"] = ds[k]
    out["names"] = np.array(names)
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(names), "cases")
    make_stream_stage_vectors(ref)


def stream_stage_inputs():
    """Inputs of the stream stages behind the BWT (mtf / rle / encode): seeded, incl. bytes >= 0x80 in runs (the reference's
    signed-char run counter), long runs, and a real BWT -> MTF chain."""
    rng = np.random.default_rng(2024)
    t = synth.markov_text(20000, 9)
    cases = {
        "empty": np.zeros(0, np.uint8), "one": np.array([65], np.uint8), "pair": np.array([65, 65], np.uint8),
        "run_low": np.full(5000, 97, np.uint8), "run_high": np.full(700, 200, np.uint8),
        "random": rng.integers(0, 256, 20000, dtype=np.uint8), "two_symbols": rng.integers(65, 67, 20000, dtype=np.uint8),
        "runs_mixed": np.repeat(rng.integers(0, 256, 900, dtype=np.uint8), rng.integers(1, 40, 900)),
        "long_runs": np.repeat(rng.integers(0, 100, 30, dtype=np.uint8), rng.integers(1, 3000, 30)),
        "markov": t,
    }
    return cases


def make_stream_stage_vectors(ref):
    out, names = {}, []
    for name, d in stream_stage_inputs().items():
        names.append(name)
        out[f"{name}/in"] = d
        out[f"{name}/mtf"] = ref.stream_stage(0, d)
        for off in (0, 1, 300):
            out[f"{name}/rle{off}"] = ref.stream_stage(1, d, off)
        out[f"{name}/enc_bit"] = ref.stream_stage(2, d)
        out[f"{name}/enc_huff"] = ref.stream_stage(3, d)
        hist = np.bincount(d, minlength=256).astype(np.uint64)
        head, hbits, codes, lens = ref.literal_coder(1, hist)
        out[f"{name}/huff_head"], out[f"{name}/huff_head_bits"] = head, np.array([hbits], np.uint64)
        out[f"{name}/huff_codes"], out[f"{name}/huff_lens"] = codes, lens
    # the chain of config 3 on a real BWT: bwt -> mtf -> rle
    t = synth.markov_text(20000, 9)
    b = ref.bwt(t)
    out["chain/bwt"], out["chain/mtf"] = b, ref.stream_stage(0, b)
    out["chain/rle"] = ref.stream_stage(1, out["chain/mtf"], 0)
    out["names"] = np.array(names)
    path = os.path.join(HERE, "stream_stage_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(names), "cases")


if __name__ == "__main__":
    main()

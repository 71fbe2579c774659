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


if __name__ == "__main__":
    main()

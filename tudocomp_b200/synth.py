"""Seeded synthetic inputs for the workloads BASELINE.json names (there is no network for corpora).

All generators return the *escaped text with its sentinel*: bytes in 0x01..0xFE followed by one 0x00, which is exactly
what the reference driver hands to a ``uses_textds`` compressor (src/tudocomp_driver/tudocomp_driver.cpp:268-270;
escaping is the identity for these alphabets, io/EscapeMap.hpp:39-64).

  markov_text  – order-3 Markov chain over 27 symbols (a-z, space), English-like unigram frequencies (config 1, 4, 5)
  dna          – i.i.d. uniform ACGT (config 2)
  repetitive   – one random lowercase block repeated, each byte mutated independently with probability p (config 3)

Also the reference's own string generators (include/tudocomp/generators/*.hpp) restated for the parity tests.
"""
from __future__ import annotations

import numpy as np

_ENGLISH_FREQ = np.array(
    [8.2, 1.5, 2.8, 4.3, 12.7, 2.2, 2.0, 6.1, 7.0, 0.15, 0.77, 4.0, 2.4, 6.7, 7.5, 1.9, 0.095, 6.0, 6.3, 9.1, 2.8, 0.98,
     2.4, 0.15, 2.0, 0.074, 18.0],
    dtype=np.float64,
)
_SYMS = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ", dtype=np.uint8)


def with_sentinel(body: np.ndarray) -> np.ndarray:
    out = np.empty(body.size + 1, dtype=np.uint8)
    out[:-1] = body
    out[-1] = 0
    return out


def _markov_table(seed: int) -> np.ndarray:
    """Cumulative transition table [27**3, 27] (float32).  Each context gets the unigram distribution skewed by a
    seeded log-normal factor, which gives ~2.7 bits/symbol of conditional entropy (English-like repetitiveness)."""
    rng = np.random.default_rng(seed + 0x5EED)
    w = _ENGLISH_FREQ[None, :] * np.exp(rng.normal(0.0, 1.6, size=(27 ** 3, 27)))
    w /= w.sum(axis=1, keepdims=True)
    cum = np.cumsum(w, axis=1).astype(np.float32)
    cum[:, -1] = 1.0
    return cum


def markov_text(n_body: int, seed: int = 1, lanes: int | None = None) -> np.ndarray:
    """n_body symbols + sentinel.  Generated as `lanes` chains advanced in lock step (vectorised), concatenated."""
    if n_body == 0:
        return with_sentinel(np.empty(0, np.uint8))
    if lanes is None:
        lanes = int(min(65536, max(1, n_body // 1024)))
    steps = -(-n_body // lanes)
    cum = _markov_table(seed)
    rng = np.random.default_rng(seed)
    ctx = rng.integers(0, 27 ** 3, size=lanes, dtype=np.int64)
    out = np.empty((steps, lanes), dtype=np.uint8)
    for s in range(steps):
        u = rng.random(lanes, dtype=np.float32)
        nxt = (cum[ctx] < u[:, None]).sum(axis=1).astype(np.int64)
        np.minimum(nxt, 26, out=nxt)
        out[s] = nxt
        ctx = (ctx * 27 + nxt) % (27 ** 3)
    body = _SYMS[out.T.reshape(-1)[:n_body]]
    return with_sentinel(body)


def markov_text_device(n_body: int, seed: int = 1, device: str = "cuda:0", lanes: int = 1 << 20):
    """The same order-3 chain as markov_text (same transition table), advanced on the GPU with torch's generator: GiB-sized
    inputs of the block-mode / sharded measurements in seconds instead of minutes (numpy: ~2 min per GiB).  Different
    random stream than markov_text, hence different bytes — only for measurements whose checker does not need the CPU
    generator (round trips, device checkers).  Returns a uint8 CUDA tensor of n_body bytes WITHOUT sentinel."""
    import torch

    cum = torch.from_numpy(_markov_table(seed)).to(device)
    syms = torch.from_numpy(_SYMS.copy()).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lanes = int(min(lanes, max(1, n_body // 1024)))
    steps = -(-n_body // lanes)
    ctx = torch.randint(0, 27 ** 3, (lanes,), generator=g, device=device, dtype=torch.int64)
    out = torch.empty((steps, lanes), dtype=torch.uint8, device=device)
    for s in range(steps):
        u = torch.rand(lanes, generator=g, device=device, dtype=torch.float32)
        nxt = (cum[ctx] < u[:, None]).sum(dim=1).clamp_(max=26)
        out[s] = syms[nxt]
        ctx = (ctx * 27 + nxt) % (27 ** 3)
    return out.t().reshape(-1)[:n_body].contiguous()


def dna(n_body: int, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n_body, dtype=np.uint8)]
    return with_sentinel(body)


def repetitive(n_body: int, seed: int = 3, block: int = 1 << 20, p: float = 0.01) -> np.ndarray:
    rng = np.random.default_rng(seed)
    block = max(1, min(block, n_body)) if n_body else 1
    base = rng.integers(97, 123, size=block, dtype=np.uint8)
    reps = -(-n_body // block) if n_body else 0
    body = np.tile(base, reps)[:n_body].copy()
    mut = rng.random(n_body, dtype=np.float32) < p
    body[mut] = rng.integers(97, 123, size=int(mut.sum()), dtype=np.uint8)
    return with_sentinel(body)


# ---- the reference's generators (include/tudocomp/generators/) ------------------------------------------------------
def fib_word(n: int) -> bytes:
    """FibonacciGenerator::generate — FibonacciGenerator.hpp:17-37."""
    if n == 1:
        return b"b"
    if n == 2:
        return b"a"
    vold, old = b"b", b"a"
    for _ in range(2, n):
        vold, old = old, old + vold
    return old


def thue_morse_word(n: int) -> bytes:
    """ThueMorseGenerator::generate — ThueMorseGenerator.hpp:17-33."""
    if n == 0:
        return b"0"
    a = bytearray(b"0")
    for _ in range(1, n):
        a += bytes(ord("1") if c == ord("0") else ord("0") for c in a)
    return bytes(a)


def run_rich_word(n: int) -> bytes:
    """RunRichGenerator::generate — RunRichGenerator.hpp:17-37."""
    t0, t1, t2 = b"0110101101001011010", b"0110101101001", b"01101011010010110101101"
    t3 = t2 + t1
    if n == 0:
        return t0
    if n == 1:
        return t1
    if n == 2:
        return t2
    for i in range(4, n):
        tmp = (t3 + t2) if i % 3 == 0 else (t3 + t0)
        t0, t1, t2, t3 = t1, t2, t3, tmp
    return t3


def random_digits(length: int, seed: int) -> bytes:
    """Same distribution as RandomUniformGenerator (uniform over '0'..'9', RandomUniformGenerator.hpp:24-39); the
    engine differs (numpy instead of std::default_random_engine), which is irrelevant for property tests."""
    rng = np.random.default_rng(seed)
    return bytes(rng.integers(ord("0"), ord("9") + 1, size=length, dtype=np.uint8))


def escape_with_sentinel(raw: bytes) -> np.ndarray:
    """{0}-escape + sentinel as the driver applies it (io/EscapeMap.hpp:39-64): 00 -> FF FE, FF -> FF FF, then 00."""
    out = bytearray()
    for b in raw:
        if b == 0x00:
            out += b"\xff\xfe"
        elif b == 0xFF:
            out += b"\xff\xff"
        else:
            out.append(b)
    out.append(0)
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()

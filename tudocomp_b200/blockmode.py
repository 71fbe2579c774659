"""Block mode (BASELINE.json config 5): a long input is cut into independent blocks, every block gets its own sentinel
and is indexed/factorised on its own; blocks are dealt round-robin to the ranks (one process per GPU).  There is no
data-path collective — the only communication is the barrier and the max-over-ranks of the step time.

The reference has no block mode; per block the semantics are exactly those of running the reference on
`Input(other, from, to)` (include/tudocomp/io/Input.hpp:241) of the escaped input.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Tuple

import numpy as np


def split_blocks(n_body: int, block_body: int) -> List[Tuple[int, int]]:
    """[from, to) byte ranges of the blocks of an n_body-byte (already escaped) input."""
    if block_body <= 0:
        raise ValueError("block size must be positive")
    return [(s, min(s + block_body, n_body)) for s in range(0, n_body, block_body)] or [(0, 0)]


def rank_blocks(num_blocks: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: block b belongs to rank b % world."""
    return list(range(rank, num_blocks, world))


def block_text(body: np.ndarray, rng: Tuple[int, int]) -> np.ndarray:
    """The text a block's TextDS sees: the slice plus its own sentinel."""
    out = np.empty(rng[1] - rng[0] + 1, dtype=np.uint8)
    out[:-1] = body[rng[0]:rng[1]]
    out[-1] = 0
    return out


def run_rank(body: np.ndarray, block_body: int, rank: int, world: int,
             factorize: Callable[[np.ndarray], np.ndarray]) -> Dict[int, np.ndarray]:
    """Factorise this rank's blocks.  `factorize(text_with_sentinel) -> (z,3) uint32 triples` is the device call in
    production (tudocomp_b200.Context) and the oracle in the CPU tests.  Factor positions are block-local."""
    ranges = split_blocks(int(body.size), block_body)
    return {b: factorize(block_text(body, ranges[b])) for b in rank_blocks(len(ranges), rank, world)}


def reduce_step_time(local_ms: float, dist=None) -> float:
    """Step time of the whole job = max over ranks (torch.distributed all_reduce MAX; identity without a group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput_mb_s(total_body_bytes: int, step_ms: float) -> float:
    return total_body_bytes / 1e6 / (step_ms / 1e3)


def gather_block_ids(owned: Iterable[int], dist=None) -> List[int]:
    """All block ids that some rank processed (sanity check that the deal covers every block exactly once)."""
    owned = sorted(owned)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return owned
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, owned)
    return sorted(b for part in out for b in part)

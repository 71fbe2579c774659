"""CPU test of the N>1 host logic (block mode): two gloo ranks deal the blocks, factorise them (the oracle stands in for
the device call — it is the checker here, the GPU path is covered by -m gpu tests) and agree on the job's step time."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from conftest import Oracle
    from tudocomp_b200 import blockmode, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    body = synth.markov_text(50000, 77)[:-1]

    def factorize(t):
        ds = oracle.textds(t)
        return oracle.factorize(ds, t.size, 3)

    mine = blockmode.run_rank(body, 12000, rank, world, factorize)
    all_ids = blockmode.gather_block_ids(mine.keys(), dist)
    step = blockmode.reduce_step_time(10.0 + 5.0 * rank, dist)
    # every block decodes back to its slice
    ok = True
    ranges = blockmode.split_blocks(body.size, 12000)
    for b, f in mine.items():
        t = blockmode.block_text(body, ranges[b])
        ok &= bool(np.array_equal(oracle.decode(f, t), t))
    ret[rank] = (sorted(mine.keys()), all_ids, step, ok)
    dist.barrier()
    dist.destroy_process_group()


def test_block_mode_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == [0, 2, 4] and ret[1][0] == [1, 3]
    assert ret[0][1] == [0, 1, 2, 3, 4] == ret[1][1]
    assert ret[0][2] == ret[1][2] == 15.0  # max over ranks
    assert ret[0][3] and ret[1][3]


def test_split_and_throughput():
    sys.path.insert(0, ROOT)
    from tudocomp_b200 import blockmode

    assert blockmode.split_blocks(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert blockmode.split_blocks(0, 4) == [(0, 0)]
    assert blockmode.rank_blocks(5, 1, 2) == [1, 3]
    assert abs(blockmode.job_throughput_mb_s(2_000_000, 500.0) - 4.0) < 1e-9

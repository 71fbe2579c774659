"""Worker of tests/test_gpu_dist.py: one rank of a torchrun job (NCCL).  Builds the sharded text index of one text over
all ranks and checks this rank's shards against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import tudocomp_b200 as tdc
    from conftest import Oracle
    from sim_dist import check_against_oracle  # the comparison helper only; the context below is the NCCL product path
    from tudocomp_b200 import synth
    from tudocomp_b200.dist import DistContext

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = tdc.load()
    ctx = DistContext.create_nccl(lib, local, dist)
    oracle = Oracle()
    size = int(os.environ.get("TDC_DIST_TEST_BYTES", str(1 << 20)))
    cases = [("markov", synth.markov_text(size, 11)), ("dna", synth.dna(size, 12)),
             ("repetitive", synth.repetitive(size // 2, 13, block=5000, p=0.01)),
             ("run", synth.with_sentinel(np.full(5000, 97, np.uint8))), ("tiny", synth.with_sentinel(np.frombuffer(b"banana", np.uint8)))]
    for name, t in cases:
        info = check_against_oracle(ctx, oracle, t, (3, 5))
        if rank == 0:
            print(f"dist ok: {name} n={t.size} world={world} shard0={info} stats={ctx.stats()}", flush=True)
    # sizes the oracle does not finish in seconds: every slot and every parse position checked on the device
    big = int(os.environ.get("TDC_DIST_TEST_BIG_BYTES", str(1 << 26)))
    for name, t in (("dna_big", synth.dna(big, 14)), ("markov_big", synth.markov_text(big // 2, 15)),
                    ("repetitive_big", synth.repetitive(big // 4, 16, block=1 << 16, p=0.01))):
        ctx.set_text(t)
        ctx.build()
        zl, zt, mn, mx = ctx.factorize(3)
        res = ctx.verify_full(3, zl, local, dist)
        assert res["ok"], (name, rank, res)
        if rank == 0:
            print(f"dist verified: {name} n={t.size} world={world} factors={zt} rank0={res}", flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE: runs the sharded multi-GPU driver (tudocomp_b200/csrc/dist_textds.cu) in the CPU interpreter
build (tests/sim/_build/libtdcsim.so) with torch.distributed/gloo standing in for NCCL, one process per rank."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIM = os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so")

AG_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
A2A_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p,
                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))


def _view(addr, nbytes):
    import torch

    return torch.frombuffer((C.c_uint8 * nbytes).from_address(addr), dtype=torch.uint8)


def make_sim_context(rank: int, world: int, dist=None):
    """dist: initialised torch.distributed (gloo) when world > 1."""
    from tudocomp_b200 import _abi
    from tudocomp_b200.dist import DistContext

    lib = _abi.TdcGpuLib(SIM)

    def allgather(_user, send, recv, nbytes):
        try:
            import torch

            mine = _view(send, nbytes).clone()
            outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, mine)
            _view(recv, nbytes * world).copy_(torch.cat(outs))
            return 0
        except Exception as e:  # noqa: BLE001
            print("allgather callback:", e, file=sys.stderr)
            return -1

    def alltoallv(_user, send, soff, scnt, recv, roff, rcnt):
        try:
            reqs, keep = [], []
            for p in range(world):
                if p == rank:
                    if scnt[p]:
                        _view(recv + roff[p], rcnt[p]).copy_(_view(send + soff[p], scnt[p]))
                    continue
                if scnt[p]:
                    t = _view(send + soff[p], scnt[p]).clone()
                    keep.append(t)
                    reqs.append(dist.isend(t, p))
                if rcnt[p]:
                    reqs.append(dist.irecv(_view(recv + roff[p], rcnt[p]), p))
            for r in reqs:
                r.wait()
            return 0
        except Exception as e:  # noqa: BLE001
            print("alltoallv callback:", e, file=sys.stderr)
            return -1

    ag, a2a = AG_FN(allgather), A2A_FN(alltoallv)
    lib.lib.tdcsim_dist_create.argtypes = [C.c_int, C.c_int, AG_FN, A2A_FN, C.c_void_p, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    lib.check(lib.lib.tdcsim_dist_create(rank, world, ag, a2a, None, C.byref(h)))
    ctx = DistContext(lib, h, rank, world)
    ctx._keepalive = (ag, a2a)
    return ctx


def check_against_oracle(ctx, oracle, t, thresholds, gather=None):
    """Every rank compares its shards with the oracle's full arrays.  Returns the concatenation check inputs."""
    from tudocomp_b200 import _abi

    ds = oracle.textds(t)
    ctx.set_text(t)
    ctx.build(_abi.SA | _abi.ISA | _abi.LCP)
    info = ctx.shard_info()
    lo, cnt, plo, pcnt = info["slot_lo"], info["slot_cnt"], info["pos_lo"], info["pos_cnt"]
    assert np.array_equal(ctx.get(_abi.SA), ds["sa"][lo:lo + cnt]), ("sa", ctx.rank, info)
    assert np.array_equal(ctx.get(_abi.ISA), ds["isa"][plo:plo + pcnt]), ("isa", ctx.rank, info)
    assert np.array_equal(ctx.get(_abi.LCP), ds["lcp"][lo:lo + cnt]), ("lcp", ctx.rank, info)
    assert ctx.max_lcp() == ds["max_lcp"]
    for thr in thresholds:
        zl, zt, mn, mx = ctx.factorize(thr)
        want = oracle.factorize(ds, t.size, thr)
        assert zt == len(want), (thr, zt, len(want))
        mine = want[(want[:, 0] >= plo) & (want[:, 0] < plo + pcnt)]
        f = ctx.factors(zl)
        got = np.stack([f["pos"], f["src"], f["len"]], 1) if zl else np.zeros((0, 3), np.uint32)
        assert np.array_equal(got, mine), (thr, ctx.rank, got[:5], mine[:5])
        wmn, wmx, _ = oracle.factor_stats(want, t.size)
        assert (mn, mx) == (wmn, wmx), (thr, mn, mx, wmn, wmx)
    return info


def worker(rank: int, world: int, port: int, case: str):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import Oracle
    from inputs import generator_strings, roundtrip_batch, small_synthetic

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    ctx = make_sim_context(rank, world, dist)
    if case == "strings":
        cases = list(roundtrip_batch()) + list(generator_strings(7))
        thr = (1, 2, 3)
    elif case == "few":  # a short run for variants of the text upload: every third reference string, one threshold
        cases = list(roundtrip_batch())[::3]
        thr = (2,)
    else:
        cases = [(nm, t) for nm, t in small_synthetic() if t.size <= 12000]
        thr = (3,)
    for name, t in cases:
        try:
            check_against_oracle(ctx, oracle, t, thr)
        except AssertionError as e:
            print(f"[rank {rank}] FAILED case {name!r} n={t.size}: {e}", file=sys.stderr)
            raise
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    worker(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])

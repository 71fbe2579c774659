"""GPU tests of the sharded text index (tdcgpu_dist_*): a single rank on one GPU (always), and one text over all
GPUs of the box through NCCL when there are at least two."""
import os
import subprocess
import sys

import numpy as np
import pytest

import tudocomp_b200 as tdc
from tudocomp_b200 import synth
from tudocomp_b200.dist import DistContext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_single_rank_shards_equal_oracle(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from sim_dist import check_against_oracle

    lib = tdc.load()
    with DistContext.create_nccl(lib, 0, None) as ctx:
        for name, t in (("markov", synth.markov_text(1 << 20, 21)), ("dna", synth.dna(1 << 20, 22)),
                        ("repetitive", synth.repetitive(1 << 19, 23, block=4000, p=0.01)),
                        ("banana", synth.with_sentinel(np.frombuffer(b"banana", np.uint8)))):
            info = check_against_oracle(ctx, oracle, t, (3,))
            assert info["slot_cnt"] == t.size and info["pos_cnt"] == t.size, name


def test_one_text_over_all_gpus_nccl():
    lib = tdc.load()
    ngpu = lib.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = min(ngpu, 8)
    port = 29400 + (os.getpid() % 500)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert r.stdout.count("dist ok:") == 5, r.stdout[-2000:]
    assert r.stdout.count("dist verified:") == 3, r.stdout[-2000:]


def test_single_rank_full_device_verification():
    """The sharded code path on one rank at 2^26 B, every slot / position checked by the device checkers."""
    lib = tdc.load()
    with DistContext.create_nccl(lib, 0, None) as ctx:
        for name, t in (("dna", synth.dna(1 << 26, 14)), ("markov", synth.markov_text(1 << 25, 15))):
            ctx.set_text(t)
            ctx.build()
            zl, zt, _, _ = ctx.factorize(3)
            res = ctx.verify_full(3, zl, 0, None)
            assert res["ok"] and res["checked_slots"] == t.size, (name, res)

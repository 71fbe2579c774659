"""CPU tests of the sharded multi-GPU driver (tudocomp_b200/csrc/dist_textds.cu): the SAME sources run in the CPU
interpreter build, one process per rank, with torch.distributed/gloo standing in for NCCL (tests/sim_dist.py).  Every
rank checks its SA / ISA / LCP shards and its part of the factor list against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.sim


@pytest.fixture(scope="module")
def simbuilt():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])


def _run_world(world, case, timeout=900, env=None):
    port = 29700 + (os.getpid() % 1500) + world + (7 if env else 0)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "sim_dist.py"), str(r), str(world), str(port), case],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, **(env or {})))
             for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{out[-3000:]}"


def test_single_rank_matches_oracle(simbuilt, oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from inputs import roundtrip_batch
    from sim_dist import check_against_oracle, make_sim_context

    ctx = make_sim_context(0, 1)
    for name, t in roundtrip_batch():
        check_against_oracle(ctx, oracle, t, (1, 3))
    ctx.close()


def test_two_ranks_gloo_reference_strings(simbuilt):
    _run_world(2, "strings")


def test_three_ranks_gloo_reference_strings(simbuilt):
    _run_world(3, "strings")


def test_two_ranks_slice_upload_experiment(simbuilt):
    """TDCGPU_DIST_SLICE_UPLOAD=nccl: every rank uploads its n/P slice and the peers exchange the rest through the collective
    (the peer-memory route, the default on GPUs with a mapped window, is covered by tests/test_gpu_dist.py on >= 2 GPUs)."""
    _run_world(2, "few", env={"TDCGPU_DIST_SLICE_UPLOAD": "nccl"})

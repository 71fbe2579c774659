#!/usr/bin/env python
"""Benchmark of the hot path: SA + ISA + LCP construction and lzss_lcp factorisation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload dna|markov|repetitive]
                    [--log2-bytes L]

A "step" is one pass of the hot path over one synthetic text: set text -> TextDS(SA|ISA|LCP) -> Factorize(threshold 3).
  value   input MB/s with the text already resident in HBM (device->device hand-over), CUDA events on the library's
          stream around exactly K steps, max over ranks.
  e2e     the same through the plugin-facing C-ABI call sequence with HOST buffers: pinned host text in (H2D inside the
          timed region), factor list out to pinned host memory (D2H inside the timed region).
  N > 1   block mode: every rank owns one GPU and an independent text of the same size (different seed), no data-path
          collective ("scaling": "weak"); NCCL is used only for the barrier and the max over ranks.
  --impl reference   times the reference's own CPU implementation (oracle/_ref, the unmodified tudocomp headers) on a
          bounded sample of the same workload, rank 0 only.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the image sets NCCL_DEBUG=VERSION, which makes NCCL print a banner on STDOUT next to the one JSON line (WARN prints it
# too: the banner's level is below WARN); whatever NCCL still has to say goes to stderr
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ.pop("NCCL_DEBUG", None)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "input MB/s for SA+LCP+lzss_lcp factorization"
THRESHOLD = 3
CPU_SAMPLE_LOG2 = 24  # 16 MiB of the same text: ~13 s of single-core reference work


def gen_text(workload: str, n_body: int, seed: int) -> np.ndarray:
    from tudocomp_b200 import synth

    if workload == "dna":
        return synth.dna(n_body, seed)
    if workload == "markov":
        return synth.markov_text(n_body, seed)
    if workload == "repetitive":
        return synth.repetitive(n_body, seed)
    raise SystemExit(f"unknown workload {workload}")


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.samples, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference (oracle/_ref) on a bounded sample.  The only place bench.py executes oracle/.
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(sample: np.ndarray, steps: int):
    path = os.path.join(ROOT, "oracle", "_ref", "libtdcref.so")
    kind = "reference"
    if os.path.exists(path):
        lib = ctypes.CDLL(path)
        lib.tdcref_lzss_lcp_factors.restype = ctypes.c_int64

        def one():
            hdr = (ctypes.c_uint64 * 3)()
            z = lib.tdcref_lzss_lcp_factors(sample.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(sample.size),
                                            ctypes.c_uint32(THRESHOLD), None, ctypes.c_uint64(0), hdr)
            assert z >= 0
            return z
    else:  # the reference did not travel: fall back to the C restatement (kind "port")
        kind = "port"
        opath = os.path.join(ROOT, "oracle", "libtdcoracle.so")
        if not os.path.exists(opath):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libtdcoracle.so"])
        lib = ctypes.CDLL(opath)
        lib.tdcoracle_lzss_lcp_factorize.restype = ctypes.c_int64
        n = sample.size
        sa, isa, lcp = (np.zeros(n, np.uint32) for _ in range(3))
        out = np.zeros((n, 3), np.uint32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

        def one():
            assert lib.tdcoracle_textds(P(sample), ctypes.c_uint32(n), P(sa), P(isa), P(lcp), None, None, None) == 0
            return lib.tdcoracle_lzss_lcp_factorize(P(sa), P(isa), P(lcp), ctypes.c_uint32(n), ctypes.c_uint32(THRESHOLD), P(out), ctypes.c_uint64(n))

    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    return kind, times


_REF_BLOCKS = None  # set before the fork: the workers of cpu_reference_parallel read their block from here


def _ref_block_worker(b):
    kind, times = cpu_reference_run(_REF_BLOCKS[b], 1)
    return kind, times[0]


def cpu_reference_parallel(blocks, steps: int):
    """All host cores: the reference is single-threaded, so the only way it can use C cores is C independent processes,
    each running the unmodified code on its own block of the workload text (what `tdc` on C files in parallel would do).
    One step = all C blocks once, wall clock from the common start to the last worker's end."""
    import multiprocessing as mp

    global _REF_BLOCKS
    _REF_BLOCKS = blocks
    kind, times = "reference", []
    with mp.get_context("fork").Pool(len(blocks)) as pool:
        for _ in range(steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_block_worker, range(len(blocks)), chunksize=1)
            times.append(time.perf_counter() - t0)
            kind = res[0][0]
    return kind, times


def gen_text_device(workload: str, n_body: int, seed: int, device: int):
    """Multi-GB texts of the sharded configuration, generated ON the GPU (torch's Philox generator: the same seed gives
    the same bytes on every rank) — numpy needs ~10 s per GB.  Returns a uint8 CUDA tensor of n_body + 1 bytes (sentinel)."""
    import torch

    if workload != "dna":
        t = torch.from_numpy(gen_text(workload, n_body, seed))
        return t.to(f"cuda:{device}")
    g = torch.Generator(device=f"cuda:{device}")
    g.manual_seed(seed)
    out = torch.empty(n_body + 1, dtype=torch.uint8, device=f"cuda:{device}")
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=f"cuda:{device}")
    chunk = 1 << 27
    for lo in range(0, n_body, chunk):
        hi = min(n_body, lo + chunk)
        idx = torch.randint(0, 4, (hi - lo,), generator=g, device=f"cuda:{device}", dtype=torch.int64)
        out[lo:hi] = lut[idx]
    out[n_body] = 0
    return out


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    return peak, src


def kernel_table(prof):
    return {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}


def dominant_roofline(prof, dev_ms, peak, peak_src, with_traffic=True):
    dom = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    if not dom:
        return None
    d = prof[dom]
    plb, plm = d["bytes"] / max(d["launches"], 1), d["ms"] / max(d["launches"], 1)
    ach = plb / 1e9 / (plm / 1e3) if d["bytes"] else None
    traffic, traffic_src = None, None
    if with_traffic:
        # DRAM bytes per launch: the committed `ncu --set full` capture of this kernel (profiles/traffic.json) gives dram
        # bytes per algorithmic byte; scaled to the average launch of this run
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
            if tj and plb:
                traffic = tj["dram_per_algorithmic_byte"] * plb
                traffic_src = "profiles/traffic.json: " + tj["capture"]
        except Exception:
            pass
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches": d["launches"],
            "avg_launch_ms": plm, "algorithmic_bytes_per_launch": plb, "share_of_step": d["ms"] / dev_ms}


def golden_case(workload, n_body, seed):
    """The tests/golden/size_hashes.json entry (fingerprints of the unmodified reference) for exactly this text, if any."""
    try:
        hashes = json.load(open(os.path.join(ROOT, "tests", "golden", "size_hashes.json")))
    except Exception:
        return None, None
    for name, c in hashes.items():
        if c.get("generator") == [workload, n_body, seed] and "factors" in c:
            return name, c
    return None, None


def sha256_of(a) -> str:
    import hashlib

    return hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest()


# ---------------------------------------------------------------------------------------------------------------------
# one text sharded over all ranks (BASELINE config 4)
# ---------------------------------------------------------------------------------------------------------------------
def dist_record(args, lib, rank, local_rank, world, dist, n_body, seed, steps, warmup, verify):
    """One text of n_body bytes sharded over all ranks: distributed prefix-doubling SA with all-to-all exchanges of rank
    buckets, LCP, lzss_lcp factorisation.  Strong scaling: value = n_body / step time (max over ranks).  Returns the
    record on rank 0 (None elsewhere).  world == 1 uses the same sharded code path on one rank when it fits, which it
    does not above ~1.5e9 B (84 B per suffix): then the single-GPU context serves as the N = 1 point."""
    import torch

    import tudocomp_b200 as tdc
    from tudocomp_b200 import blockmode
    from tudocomp_b200.dist import DistContext

    single = world == 1 and n_body > 1_500_000_000
    d_text = gen_text_device(args.workload, n_body, seed, local_rank)  # the same text on every rank
    n = int(d_text.numel())
    h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_text.copy_(d_text)
    torch.cuda.synchronize()
    ctx = tdc.Context(lib, local_rank) if single else DistContext.create_nccl(lib, local_rank, dist if world > 1 else None)

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(host):
        if single:
            if host:
                ctx.set_text_host_ptr(h_text.data_ptr(), n)
            else:
                ctx.set_text_device(d_text.data_ptr(), n)
            ctx.build(tdc.SA | tdc.ISA | tdc.LCP)
            z, mn, mx = ctx.factorize(THRESHOLD)
            return z, z, mn, mx
        ctx.set_text_ptr(h_text.data_ptr() if host else d_text.data_ptr(), n, on_device=not host)
        ctx.build()
        return ctx.factorize(THRESHOLD)

    zl, zt, mn, mx = step(False)
    h_factors = torch.empty(int(zl * 1.05) + 1024, 3, dtype=torch.int32).pin_memory()
    for _ in range(max(warmup - 1, 0)):
        step(False)
    lib.profile_reset()
    lib.profile_enable(True)
    barrier()
    launches0 = lib.launch_count()
    ctx.event_record(0)
    for _ in range(steps):
        zl, zt, mn, mx = step(False)
    ctx.event_record(1)
    barrier()
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = lib.launch_count() - launches0
    lib.profile_enable(False)
    prof = lib.profile()
    phases = ctx.phases()
    stats = ctx.sa_stats() if single else ctx.stats()
    info = {"slot_lo": 0, "slot_cnt": n, "pos_lo": 0, "pos_cnt": n} if single else ctx.shard_info()
    verified = None
    if verify:  # outside the timed regions, on the results of the last resident step
        try:
            verified = ctx.check(THRESHOLD, zl) if single else ctx.verify_full(THRESHOLD, zl, local_rank, dist if world > 1 else None)
            verified["how"] = "every SA/ISA/LCP slot and every parse position, device checkers (csrc/check.cu)"
        except Exception as e:  # noqa: BLE001  (e.g. the gathered arrays do not fit next to the shards)
            verified = {"ok": None, "error": f"{type(e).__name__}: {e}"[:300]}
    resident_result = (int(zt), int(mn), int(mx))
    barrier()
    t1 = time.perf_counter()
    ctx.event_record(2)
    for _ in range(steps):
        zl, zt, mn, mx = step(True)
        ctx.get_factors_into(h_factors.data_ptr(), h_factors.shape[0])
    ctx.event_record(3)
    barrier()
    e2e_ms = max(ctx.event_elapsed_ms(2, 3), 1e3 * (time.perf_counter() - t1))
    # the host-text route (slice upload + peer exchange on several GPUs) must give the verified resident step's factorisation
    e2e_consistent = (int(zt), int(mn), int(mx)) == resident_result
    ms_step = blockmode.reduce_step_time(dev_ms / steps, dist if world > 1 else None)
    ms_step_e2e = blockmode.reduce_step_time(e2e_ms / steps, dist if world > 1 else None)
    ctx.close()
    del d_text, h_text, h_factors
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = load_peaks()
    roof = dominant_roofline(prof, dev_ms, peak, peak_src, with_traffic=False)
    if roof:
        roof["note"] = "rank 0's kernels; the exchanges are the gap between the kernel sum and the step time"
    return {"metric": METRIC, "value": blockmode.job_throughput_mb_s(n_body, ms_step), "unit": "MB/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}_{n_body}B_single_text_sharded: distributed SA+ISA+LCP + lzss_lcp(threshold={THRESHOLD})",
                       "text_bytes": n, "threshold": THRESHOLD, "index_bits": 32,
                       "l2_policy": "working set per GPU far larger than the 126 MB L2; no flush needed",
                       "parallelism": (f"one text sharded over {world} GPUs; all-to-all of rank buckets, rank updates and rank requests"
                                       if not single else "N = 1 point of the strong-scaling curve: the single-GPU context on the same text")},
            # slice upload (default on several GPUs with the peer-memory window): every rank copies its n/P slice from the host
            # and the peers exchange the rest over NVLink; otherwise every rank uploads the whole text
            "e2e": {"value": blockmode.job_throughput_mb_s(n_body, ms_step_e2e), "unit": "MB/s",
                    "h2d_bytes_per_step": n if (world > 1 and stats.get("p2p") and os.environ.get("TDCGPU_DIST_SLICE_UPLOAD", "1")[:1] not in ("0", "n")) else n * world,
                    "d2h_bytes_per_step": int(12 * zt), "ms_per_step": ms_step_e2e},
            "gpu_launches": int(launches), "roofline": roof, "factors": int(zt), "factor_len": [int(mn), int(mx)],
            "dist_stats": stats, "shard_rank0": info, "verify": verified, "e2e_same_factorisation_as_verified_step": e2e_consistent,
            "kernels": kernel_table(prof),
            "last_step_phases_ms": {k: round(v, 3) for k, v in phases}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dna", choices=["dna", "markov", "repetitive"])
    ap.add_argument("--log2-bytes", type=int, default=30, help="text body size per GPU = 2^L bytes (default 1 GiB)")
    ap.add_argument("--bytes", type=int, default=0, help="text body size in bytes (overrides --log2-bytes), e.g. 4000000000")
    ap.add_argument("--no-verify", action="store_true", help="skip the full device-side verification after the timed regions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-block-driver", action="store_true", help="skip the tdc_block_gpu sub-record (config 5 through the plugin driver)")
    ap.add_argument("--no-dist", action="store_true", help="default mode: skip the sharded-text sub-records (config 4)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the double-buffered (2 contexts) end-to-end measurement")
    ap.add_argument("--dist-bytes", type=int, default=2_000_000_000, help="text size of the sharded sub-record (strong scaling over N)")
    ap.add_argument("--mode", default="block", choices=["block", "dist"],
                    help="'block' (default, the headline): one independent text per GPU (weak scaling, no collective), with the "
                         "sharded-text measurement attached as `dist`; 'dist': ONE text of --bytes sharded over all GPUs is the line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_body = args.bytes if args.bytes > 0 else 1 << args.log2_bytes
    size_name = f"{n_body}B" if args.bytes > 0 else f"2^{args.log2_bytes}B"
    seed_of = {"dna": 2, "markov": 1, "repetitive": 3}[args.workload]
    workload_name = f"{args.workload}_{size_name}_per_gpu: TextDS SA+ISA+LCP + lzss_lcp(threshold={THRESHOLD}) factorize"
    config = {"workload": workload_name, "text_bytes_per_gpu": n_body + 1, "threshold": THRESHOLD, "index_bits": 32,
              "l2_policy": "inputs (>= 1 GiB text, 4 GiB arrays) are far larger than the 126 MB L2; no flush needed",
              "parallelism": f"block mode, {max(world, args.gpus)} independent texts, no collective"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # Bounded sample of the same workload: C blocks of 2^24 B (C = host cores), consecutive pieces of the workload text,
        # one unmodified single-threaded reference process per block, all at once.  `single_core` is one block on one
        # core, the figure that corresponds to the reference's own published single-threaded runs.
        sample_body = min(n_body, 1 << CPU_SAMPLE_LOG2)
        cores = max(1, min(os.cpu_count() or 1, 64, n_body // sample_body))
        body = gen_text(args.workload, sample_body * cores, seed_of)[:-1]  # same generator and seed as rank 0's text
        blocks = []
        for b in range(cores):
            blk = np.empty(sample_body + 1, np.uint8)
            blk[:-1] = body[b * sample_body:(b + 1) * sample_body]
            blk[-1] = 0
            blocks.append(blk)
        kind, t1 = cpu_reference_run(blocks[0], 1)
        single = sample_body / 1e6 / t1[0]
        kind, times = cpu_reference_parallel(blocks, args.warmup + args.steps)
        times = times[args.warmup:]
        mean_t = sum(times) / len(times)
        mbps = cores * sample_body / 1e6 / mean_t
        sample = (f"{cores} consecutive blocks of 2^{int(np.log2(sample_body))} B of the workload text (+ sentinel each), "
                  f"one single-threaded reference process per block, all {cores} at once; single_core = one block alone")
        config["reference_sample"] = (f"NOT one {size_name} text: {cores} x 2^{int(np.log2(sample_body))} B blocks of it (bounded sample; the "
                                      "reference is single-threaded and needs ~15 min per 2^30 B text, tests/golden/size_hashes.json "
                                      "holds its wall times at full size)")
        line = {"impl": "reference", "metric": METRIC, "value": mbps, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * mean_t, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": mbps, "unit": "MB/s", "cores": cores, "kind": kind, "single_core": single, "sample": sample},
                "e2e": {"value": mbps, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------------ our arm
    import torch

    import tudocomp_b200 as tdc
    from tudocomp_b200 import blockmode

    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    lib = tdc.load()  # raises if the CUDA library is missing: no fallback
    verify = not args.no_verify

    if args.mode == "dist":
        sampler = ClockSampler(dev)
        sampler.start()
        line = dist_record(args, lib, rank, local_rank, world, dist, n_body, seed_of, args.steps, args.warmup, verify)
        clocks = sampler.stop()
        if rank == 0:
            line["clocks"] = clocks
            print(json.dumps(line))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    ctx = tdc.Context(lib, dev)
    seed = seed_of + 1000 * rank
    text = gen_text(args.workload, n_body, seed)
    n = int(text.size)
    h_text = torch.from_numpy(text).pin_memory()
    d_text = h_text.to(f"cuda:{dev}", non_blocking=False)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ctx.set_text_device(d_text.data_ptr(), n)
        ctx.build(tdc.SA | tdc.ISA | tdc.LCP)
        return ctx.factorize(THRESHOLD)

    # factor buffer on the host for the e2e arm (pinned), sized after the first step
    z0, _, _ = step_resident()
    h_factors = torch.empty(int(z0 * 1.05) + 1024, 3, dtype=torch.int32).pin_memory()

    def step_e2e(c=None, hf=None, h_in=None):
        c = c or ctx
        c.set_text_host_ptr(h_in if h_in else h_text.data_ptr(), n)
        c.build(tdc.SA | tdc.ISA | tdc.LCP)
        z, mn, mx = c.factorize(THRESHOLD)
        hf = hf if hf is not None else h_factors
        c.get_factors_into(hf.data_ptr() if hasattr(hf, "data_ptr") else hf.ctypes.data, hf.shape[0])
        return z, mn, mx

    for _ in range(args.warmup):
        z, mn, mx = step_resident()

    # ---- timed: resident ----
    sampler = ClockSampler(dev)
    sampler.start()
    lib.profile_reset()
    lib.profile_enable(True)
    barrier()
    launches0 = lib.launch_count()
    ctx.event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        z, mn, mx = step_resident()
    ctx.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = lib.launch_count() - launches0
    lib.profile_enable(False)
    prof = lib.profile()
    stats = ctx.sa_stats()
    phases = ctx.phases()

    # ---- verification of the last timed step's results, outside the timed region ----
    verified = None
    if verify:
        try:
            verified = ctx.check(THRESHOLD, z)  # every slot and every parse position, on the device
            verified["how"] = "device checkers (csrc/check.cu): ISA[SA[i]]==i, Burkhardt-Kaerkkaeinen order, LCP by direct comparison, the reference's PSV/NSV rule at every parse position"
            gname, gold = golden_case(args.workload, n_body, seed)
            if gold is not None:
                f = ctx.factors(z)
                verified["reference_fingerprint"] = {"case": gname, "what": "sha256 of the factor list == the unmodified reference's (tests/golden/size_hashes.json)",
                                                     "match": bool(z == gold["factors"]["z"] and sha256_of(f) == gold["factors"]["factors"])}
                verified["ok"] = bool(verified["ok"] and verified["reference_fingerprint"]["match"])
                del f
        except Exception as e:  # noqa: BLE001
            verified = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}
        if world > 1:
            flag = torch.tensor([0 if verified.get("ok") else 1], device=f"cuda:{dev}")
            dist.all_reduce(flag)
            verified["all_ranks_ok"] = bool(flag.item() == 0)

    # ---- timed: end to end with host buffers, one context (serial: copy in, compute, copy out) ----
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    ctx.event_record(2)
    t1 = time.perf_counter()
    for _ in range(args.steps):
        ze, _, _ = step_e2e()
    ctx.event_record(3)
    barrier()
    wall_e2e = time.perf_counter() - t1
    ms_e2e_serial = max(ctx.event_elapsed_ms(2, 3), 1e3 * wall_e2e) / args.steps  # host-inclusive, take the larger

    # ---- the same with PAGEABLE host buffers (what the tdc driver's View / std::vector are): staged copies (csrc/host_copy.cu)
    ms_e2e_pageable = None
    try:
        pg_text = text  # numpy, pageable
        pg_factors = np.empty((h_factors.shape[0], 3), np.int32)
        step_e2e(None, pg_factors, pg_text.ctypes.data)
        ctx.sync()
        t1 = time.perf_counter()
        for _ in range(min(args.steps, 5)):
            step_e2e(None, pg_factors, pg_text.ctypes.data)
        ctx.sync()
        ms_e2e_pageable = 1e3 * (time.perf_counter() - t1) / min(args.steps, 5)
        del pg_factors
    except Exception as e:  # noqa: BLE001
        ms_e2e_pageable = f"{type(e).__name__}: {e}"[:200]

    # ---- timed: end to end, double-buffered: two contexts on the GPU, each driven by its own host thread, so that one text's
    # H2D / D2H overlaps the other's kernels (a stream of independent texts is exactly what block mode is).  Same C-ABI calls,
    # same pinned host buffers, every step copies its text in and its factor list out.
    ms_e2e_pipe, pipe_error = None, None
    if not args.no_pipeline:
        ctx2 = None
        try:
            ctx2 = tdc.Context(lib, dev)
            h_factors2 = torch.empty_like(h_factors).pin_memory()
            step_e2e(ctx2, h_factors2)
            counts = [(args.steps + 1) // 2, args.steps // 2]
            errs = []

            def worker(c, hf, k):
                try:
                    for _ in range(k):
                        step_e2e(c, hf)
                except Exception as e:  # noqa: BLE001
                    errs.append(e)

            barrier()
            ctx2.sync()
            t1 = time.perf_counter()
            th = [threading.Thread(target=worker, args=(ctx, h_factors, counts[0])), threading.Thread(target=worker, args=(ctx2, h_factors2, counts[1]))]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            ctx.sync()
            ctx2.sync()
            ms_e2e_pipe = 1e3 * (time.perf_counter() - t1) / args.steps
            if errs:
                raise errs[0]
        except Exception as e:  # noqa: BLE001
            ms_e2e_pipe, pipe_error = None, f"{type(e).__name__}: {e}"[:300]
        finally:
            if ctx2 is not None:
                ctx2.close()
            torch.cuda.empty_cache()
    ms_e2e_local = ms_e2e_pipe if (ms_e2e_pipe and ms_e2e_pipe < ms_e2e_serial) else ms_e2e_serial

    # ---- timed: host text in -> finished lzss_lcp(bit) archive out (device-side lzss::encode_text).
    # Not the headline: the D2H side is the archive instead of the factor list.  BitCoder's literal words (8 bits) need no
    # host-side table; with HuffmanCoder the C++ plugin builds the table from the same device histogram.
    codes, lens = np.arange(256, dtype=np.uint64), np.full(256, 8, np.uint8)

    def step_archive(h_arc_ptr=0, cap=0):
        ctx.set_text_host_ptr(h_text.data_ptr(), n)
        ctx.build(tdc.SA | tdc.ISA | tdc.LCP)
        ctx.factorize(THRESHOLD)
        ctx.literal_histogram()
        nbits = ctx.encode(codes, lens)
        if h_arc_ptr:
            ctx.encoded_into(h_arc_ptr, cap)
        return nbits

    # informational, rank-local (no collective inside: a failure here must neither hang the other ranks nor cost the headline)
    arc_bits, arc_ms, prof_arc, arc_error, arc_match = 0, None, {}, None, None
    try:
        arc_bits = step_archive()
        h_arc = torch.empty(arc_bits // 8 + 64, dtype=torch.uint8).pin_memory()
        nb = step_archive(h_arc.data_ptr(), h_arc.numel())
        gname, gold = golden_case(args.workload, n_body, seed)
        if gold is not None and "bit" in gold and verify:
            ln = gold["bit"]["archive_len"]
            arc_match = bool(ln <= h_arc.numel() and sha256_of(h_arc.numpy()[:ln]) == gold["bit"]["archive"])
        lib.profile_reset()
        lib.profile_enable(True)
        torch.cuda.synchronize()
        ctx.sync()
        ctx.event_record(4)
        t2 = time.perf_counter()
        asteps = min(args.steps, 5)
        for _ in range(asteps):
            step_archive(h_arc.data_ptr(), h_arc.numel())
        ctx.event_record(5)
        ctx.sync()
        arc_ms = max(ctx.event_elapsed_ms(4, 5), 1e3 * (time.perf_counter() - t2)) / asteps
        lib.profile_enable(False)
        prof_arc = {k: v for k, v in lib.profile().items() if k.startswith("enc_")}
    except Exception as e:  # noqa: BLE001
        lib.profile_enable(False)
        arc_error = f"{type(e).__name__}: {e}"
    clocks = sampler.stop()

    # whole-job step time = max over ranks (block mode: no other communication)
    ms_step = blockmode.reduce_step_time(dev_ms / args.steps, dist)
    ms_step_e2e = blockmode.reduce_step_time(ms_e2e_local, dist)
    ms_step_e2e_serial = blockmode.reduce_step_time(ms_e2e_serial, dist)
    total_bytes = n_body * world
    value = blockmode.job_throughput_mb_s(total_bytes, ms_step)
    e2e_value = blockmode.job_throughput_mb_s(total_bytes, ms_step_e2e)

    line = None
    if rank == 0:
        peak, peak_src = load_peaks()
        roof = dominant_roofline(prof, dev_ms, peak, peak_src)
        pipeline_bytes = (25.0 * n + 12.0 * z)  # compulsory traffic of SA+ISA+LCP+factorisation
        pipeline = {"algorithmic_bytes": pipeline_bytes, "achieved_GBps": pipeline_bytes / 1e9 / (ms_step / 1e3),
                    "frac_of_peak": pipeline_bytes / 1e9 / (ms_step / 1e3) / peak}
        line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(12 * ze),
                        "ms_per_step": ms_step_e2e,
                        "mode": ("double-buffered: 2 device contexts per GPU driven by 2 host threads through the C ABI, pinned host text in and "
                                 "factor list out every step" if ms_e2e_local == ms_e2e_pipe else "serial: one context, copy in -> compute -> copy out"),
                        "serial_ms_per_step": ms_step_e2e_serial, "serial_value": blockmode.job_throughput_mb_s(total_bytes, ms_step_e2e_serial),
                        "pipelined_ms_per_step_rank0": ms_e2e_pipe, "pipeline_error": pipe_error,
                        "pageable_serial_ms_per_step_rank0": ms_e2e_pageable},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "pipeline_roofline": pipeline,
                "verified": bool(verified and verified.get("ok") and verified.get("all_ranks_ok", True)) if verify else None,
                "verify": verified,
                "factors": int(z), "factor_len": [int(mn), int(mx)], "sa_stats": stats, "kernels": kernel_table(prof),
                "last_step_phases_ms": {k: round(v, 3) for k, v in phases}, "wall_s_timed": wall}
        if arc_error or not arc_ms:
            line["archive"] = {"error": arc_error or "not measured"}
        else:
            enc_ms = sum(v["ms"] for v in prof_arc.values()) / asteps
            enc_bytes = sum(v["bytes"] for v in prof_arc.values()) / asteps
            line["archive"] = {"what": "rank 0: host text -> lzss_lcp(coder=bit,threshold=3) archive in pinned host memory, through the C ABI "
                                       "(build + factorize + literal histogram + device-side lzss::encode_text + D2H of the archive)",
                               "value": n_body / 1e6 / (arc_ms / 1e3), "unit": "MB/s", "ms_per_step": arc_ms,
                               "h2d_bytes_per_step": n, "d2h_bytes_per_step": int((arc_bits + 7) // 8 + 1), "archive_bits": int(arc_bits),
                               "byte_identical_to_reference": arc_match,
                               "encode_kernels_ms_per_step": round(enc_ms, 3),
                               "encode_algorithmic_GBps": (enc_bytes / 1e9 / (enc_ms / 1e3)) if enc_ms else None,
                               "kernels": kernel_table(prof_arc)}
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
            sample_body = min(n_body, 1 << CPU_SAMPLE_LOG2)
            sample = text[: sample_body + 1].copy()
            sample[-1] = 0
            kind, times = cpu_reference_run(sample, 1)
            line["cpu_baseline"] = {"value": sample_body / 1e6 / times[0], "unit": "MB/s", "cores": 1, "kind": kind,
                                    "host_cores_available": os.cpu_count(),
                                    "sample": f"first 2^{int(np.log2(sample_body))} B of rank 0's text + sentinel, one run; the reference is single-threaded"}
            try:
                plug = plugin_e2e(text, n_body, args.steps)
            except Exception as e:  # noqa: BLE001
                plug = {"error": f"{type(e).__name__}: {e}"[:200]}
            if plug:
                line["plugin_e2e"] = plug
    # ---- config 4 next to the headline: ONE text sharded over all N GPUs (strong scaling over N) ----
    ctx.close()
    del d_text, h_text, h_factors
    torch.cuda.empty_cache()
    if not args.no_dist:
        recs = []
        sizes = [args.dist_bytes] + ([4_000_000_000] if world >= 4 else [])
        for nb in sizes:
            try:
                rec = dist_record(args, lib, rank, local_rank, world, dist, nb, 4, min(args.steps, 5), min(args.warmup, 2), verify)
            except Exception as e:  # noqa: BLE001
                rec = {"error": f"{type(e).__name__}: {e}"[:300], "text_bytes": nb}
                if world > 1:
                    raise  # a rank that fails alone would hang the others in the next collective
            if rank == 0:
                recs.append(rec)
        if rank == 0:
            line["dist"] = recs
    # ---- config 5 through the real plugin driver: tdc_block_gpu over the GPU registry, 256 MiB blocks, Huffman coder ----
    if rank == 0 and not args.no_block_driver:
        try:
            bd = block_driver_record(world, local_rank)
        except Exception as e:  # noqa: BLE001
            bd = {"error": f"{type(e).__name__}: {e}"[:300]}
        if bd:
            line["block_driver"] = bd
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def block_driver_record(world, device):
    """BASELINE config 5 as a cold command-line run: `build/tdc_block_gpu -a "lzss_lcp(coder=huff)" -b 268435456 -g N` on an
    order-3 Markov text in /dev/shm (4 blocks per GPU, generated on the GPU), one worker thread per GPU, wall clock of the
    whole process incl. CUDA start-up.  The per-block times come from the driver's own log.  Absent binary -> no record."""
    import re
    import shutil
    import subprocess
    import torch
    from tudocomp_b200 import synth
    exe = os.path.join(ROOT, "build", "tdc_block_gpu")
    if not os.path.exists(exe) or not os.path.isdir("/dev/shm"):
        return None
    gib = 4 if world == 1 else world
    if shutil.disk_usage("/dev/shm").free < (2 * gib + 1) << 30:
        return {"skipped": "not enough room in /dev/shm"}
    src, dst = "/dev/shm/tdc_bench_block_in.txt", "/dev/shm/tdc_bench_block_out.tdcb"
    try:
        with open(src, "wb") as f:
            for i in range(gib):
                f.write(synth.markov_text_device(1 << 30, 500 + i, f"cuda:{device}").cpu().numpy().tobytes())
        torch.cuda.empty_cache()
        env = dict(os.environ, TDC_BLOCK_VERBOSE="1")
        for k in ("CUDA_VISIBLE_DEVICES",):  # the workers address the GPUs of the box themselves
            env.pop(k, None)
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-a", "lzss_lcp(coder=huff)", "-b", str(1 << 28), "-g", str(world), src, "-o", dst],
                           capture_output=True, text=True, timeout=600, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        comp = [float(x) for x in re.findall(r"compress ([0-9.]+) ms", r.stderr)]
        ready = [float(x) for x in re.findall(r"block buffers ready after ([0-9.]+) ms", r.stderr)]
        out_bytes = os.path.getsize(dst)
        comp_sorted = sorted(comp)
        return {"what": "tdc_block_gpu -a lzss_lcp(coder=huff) -b 268435456 -g N: cold command-line run over the GPU-only registry, file in "
                        "/dev/shm -> container in /dev/shm, wall clock of the whole process (CUDA context creation included)",
                "input_bytes": gib << 30, "blocks": len(comp), "gpus": world, "wall_s": round(wall, 3),
                "MB_per_s": round((gib << 30) / 1e6 / wall, 1), "container_bytes": out_bytes,
                "worker_ready_after_ms": [round(x) for x in ready],
                "compress_ms_per_block_median": round(comp_sorted[len(comp_sorted) // 2], 1) if comp else None,
                "compress_ms_per_block_first": round(comp[0], 1) if comp else None}
    finally:
        for p_ in (src, dst):
            try:
                os.remove(p_)
            except OSError:
                pass


def plugin_e2e(text, n_body, steps):
    """The same step through the C++ plugin the `tdc` driver uses: LZSSLCPCompressor<coder, GpuTextDS>::compress on an
    in-memory Input / Output, K times in one process (build/tdc_plugin_bench, built against the reference headers in the
    build container; absent binary -> no record)."""
    exe = os.path.join(ROOT, "build", "tdc_plugin_bench")
    if not os.path.exists(exe):
        return None
    import shutil
    import tempfile

    out = {}
    try:
        where = next(d for d in ("/dev/shm", tempfile.gettempdir(), ROOT) if os.path.isdir(d) and shutil.disk_usage(d).free > 2 * n_body + (1 << 28))
    except StopIteration:
        return {"error": "no scratch directory with room for the input file"}
    with tempfile.NamedTemporaryFile(dir=where, suffix=".txt") as f:
        f.write(memoryview(text[:-1]))  # the driver adds the sentinel itself
        f.flush()
        for coder in ("bit", "huff"):
            try:
                r = subprocess.run([exe, f.name, coder, str(THRESHOLD), str(max(2, min(steps, 5))), "1"], capture_output=True, text=True, timeout=600)
                rec = json.loads(r.stdout.strip().splitlines()[-1])
                rec["value"] = n_body / 1e6 / (rec["ms_per_step"] / 1e3)
                rec["unit"] = "MB/s"
                out[coder] = rec
            except Exception as e:  # noqa: BLE001
                out[coder] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


if __name__ == "__main__":
    sys.exit(main())

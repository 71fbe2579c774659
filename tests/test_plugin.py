"""The reference-side boundary: the stock `tdc` driver built with the GPU text index plugged into its registry
(tudocomp_b200/plugin).  CPU tests check that the binaries exist, list the GPU provider and fail loudly without a
device; GPU tests check byte-identical archives against the unmodified reference driver and round trips."""
import os
import subprocess

import numpy as np
import pytest

from tudocomp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF, GPU, GPU_ONLY = (os.path.join(ROOT, "build", b) for b in ("tdc_ref", "tdc_gpu", "tdc_gpu_only"))
BLOCK_REF, BLOCK_GPU = (os.path.join(ROOT, "build", b) for b in ("tdc_block_ref", "tdc_block_gpu"))


def _need_bins():
    if not all(os.path.exists(p) for p in (REF, GPU, GPU_ONLY)):
        pytest.skip("tdc drivers not built (bash tudocomp_b200/plugin/build_tdc.sh; needs /root/reference)")


def _run(binary, algo, src, dst, extra=()):
    return subprocess.run([binary, "-a", algo, src, "-o", dst, "--force", *extra], capture_output=True, text=True)


def test_registry_lists_gpu_textds():
    _need_bins()
    out = subprocess.run([GPU, "--list"], capture_output=True, text=True).stdout
    assert "gpu(compress" in out and "Text index built on the GPU" in out
    assert "lzss_lcp(coder, textds" in out and "bwt(textds" in out
    only = subprocess.run([GPU_ONLY, "--list"], capture_output=True, text=True).stdout
    assert "textds = gpu(" in only  # the GPU index is the default of lzss_lcp / bwt in the GPU-only registry


def test_gpu_driver_fails_loudly_without_device(tmp_path):
    _need_bins()
    import tudocomp_b200 as tdc
    if tdc.load().device_count() > 0:
        pytest.skip("a GPU is present")
    src = tmp_path / "in.txt"
    src.write_bytes(b"hello hello hello world world")
    r = _run(GPU, "lzss_lcp(ascii, gpu)", str(src), str(tmp_path / "o.tdc"))
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    # the CPU providers of the same binary still work: the plugin is an addition, not a replacement
    r = _run(GPU, "lzss_lcp(coder=ascii)", str(src), str(tmp_path / "o.tdc"))
    assert r.returncode == 0
    assert (tmp_path / "o.tdc").read_bytes().startswith(b"lzss_lcp(coder=ascii)%30:6:12:6:16:hello ")


@pytest.mark.sim
def test_plugin_logic_over_the_simulator_library(tmp_path):
    """CPU check of the C++ plugin code itself (GpuTextDS.hpp: array hand-over, device factor list, device-side
    encode_text with the reference's Huffman table): the GPU-only driver is run with tests/sim's interpreter build of the
    SAME C ABI in place of libtdcgpu.so (LD_LIBRARY_PATH wins over the binary's RUNPATH).  Test infrastructure only —
    the product library stays the nvcc build; archives must equal the unmodified reference driver's byte for byte."""
    _need_bins()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    simdir = tmp_path / "simlib"
    simdir.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so"), simdir / "libtdcgpu.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(simdir))
    cases = {"markov": synth.markov_text(8000, 5)[:-1].tobytes(), "dna": synth.dna(4000, 6)[:-1].tobytes(),
             "binary_with_escapes": bytes(np.random.default_rng(3).integers(0, 256, 3000, dtype=np.uint8)),
             "run": b"a" * 2000, "empty": b""}
    for name, data in cases.items():
        src = tmp_path / f"{name}.bin"
        src.write_bytes(data)
        for algo in ("lzss_lcp(coder=huff)", "lzss_lcp(coder=bit,threshold=5)", "lzss_lcp(coder=ascii)", "bwt",
                     "bwt:mtf:rle:encode(huff)", "mtf:rle(offset=2):encode(bit)"):  # all four stages of config 3 on the "device"
            a, b = str(tmp_path / "ref.tdc"), str(tmp_path / "sim.tdc")
            assert _run(REF, algo, str(src), a).returncode == 0
            r = subprocess.run([GPU_ONLY, "-a", algo, str(src), "-o", b, "--force"], capture_output=True, text=True, env=env)
            assert r.returncode == 0, (name, algo, r.stderr)
            assert open(a, "rb").read() == open(b, "rb").read(), (name, algo)
    # the GPU-aware chain (GpuChainCompressor.hpp): host-only stages in the middle / at either end of a chain take the
    # fallback routes (download for a host stage, host buffer between two host stages, upload of a host stage's output),
    # and the reference's all-host route (TDCGPU_HOST_CHAIN=1) gives the same bytes as the device-resident one
    src = tmp_path / "markov.bin"
    for algo in ("bwt:noop:mtf:encode(huff)", "noop:mtf:noop:rle:noop", "mtf:encode(ascii):rle", "bwt:mtf:rle:encode(huff)"):
        a, b, c = str(tmp_path / "ref.tdc"), str(tmp_path / "sim.tdc"), str(tmp_path / "simh.tdc")
        assert _run(REF, algo, str(src), a).returncode == 0
        r = subprocess.run([GPU_ONLY, "-a", algo, str(src), "-o", b, "--force"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, (algo, r.stderr)
        r = subprocess.run([GPU_ONLY, "-a", algo, str(src), "-o", c, "--force"], capture_output=True, text=True, env=dict(env, TDCGPU_HOST_CHAIN="1"))
        assert r.returncode == 0, (algo, r.stderr)
        assert open(a, "rb").read() == open(b, "rb").read() == open(c, "rb").read(), algo
        back = str(tmp_path / "back.bin")
        r = subprocess.run([GPU_ONLY, "-d", b, "-o", back, "--force"], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and open(back, "rb").read() == src.read_bytes(), (algo, r.stderr)
    # per-provider GPU classes inside the reference's own TextDS (GpuProviders.hpp; the provider concept of ds/TextDS.hpp:23-29):
    # `textds(sa=gpu)`, `textds(sa=gpu, lcp=gpu, isa=gpu)`; the later providers find the text resident (gpu_text_reused = 1)
    for name in ("markov", "binary_with_escapes", "empty"):
        src = str(tmp_path / f"{name}.bin")
        for algo, sel in (("lzss_lcp(coder=huff)", "lzss_lcp(coder=huff, textds=textds(sa=gpu))"),
                          ("lzss_lcp(coder=bit)", "lzss_lcp(coder=bit, textds=textds(sa=gpu, lcp=gpu, isa=gpu))"),
                          ("lzss_lcp(coder=huff)", 'lzss_lcp(coder=huff, textds=textds(lcp=gpu, isa=gpu, compress="plain"))'),
                          ("bwt", "bwt(textds=textds(sa=gpu))")):
            a, b = str(tmp_path / "ref.tdc"), str(tmp_path / "sim.tdc")
            assert _run(REF, algo, src, a, ["--raw"]).returncode == 0
            r = subprocess.run([GPU, "-a", sel, src, "-o", b, "--force", "--raw", "--stats"], capture_output=True, text=True, env=env)
            assert r.returncode == 0, (name, sel, r.stderr)
            assert open(a, "rb").read() == open(b, "rb").read(), (name, sel)
            if "sa=gpu, lcp=gpu, isa=gpu" in sel and name == "markov":
                assert r.stdout.count('"gpu_text_reused"') == 3 and r.stdout.count('"value": "1"') >= 2, r.stdout[:2000]
    # lcpcomp: the reference's own strategies consume the GPU text index through require_* / release_*
    # (arrays come back bit-packed by the device, compress=delayed); mixed registry, so compare under --raw
    for name in ("markov", "empty"):
        src = str(tmp_path / f"{name}.bin")
        for opts in ("coder=huff", "coder=ascii,comp=heap", "coder=huff,comp=max_lcp", "coder=huff,comp=plcppeaks,dec=compact"):
            a, b = str(tmp_path / "ref.tdc"), str(tmp_path / "sim.tdc")
            assert _run(REF, f"lcpcomp({opts})", src, a, ["--raw"]).returncode == 0
            r = subprocess.run([GPU, "-a", f"lcpcomp({opts},textds=gpu)", src, "-o", b, "--force", "--raw"], capture_output=True, text=True, env=env)
            assert r.returncode == 0, (name, opts, r.stderr)
            assert open(a, "rb").read() == open(b, "rb").read(), (name, opts)
    # the `compress` option of the GPU text index (plain: 32-bit arrays; delayed / compressed: bit-packed on the device)
    src = str(tmp_path / "markov.bin")
    assert _run(REF, "lcpcomp(coder=huff)", src, str(tmp_path / "ref.tdc"), ["--raw"]).returncode == 0
    for cm in ("plain", "compressed"):
        r = subprocess.run([GPU, "-a", f'lcpcomp(coder=huff,textds=gpu(compress="{cm}"))', src, "-o", str(tmp_path / "sim.tdc"), "--force", "--raw"],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, (cm, r.stderr)
        assert open(tmp_path / "ref.tdc", "rb").read() == open(tmp_path / "sim.tdc", "rb").read(), cm
    # the host-side encode_text (A/B switch) gives the same bytes
    r = subprocess.run([GPU_ONLY, "-a", "lzss_lcp(coder=huff)", str(tmp_path / "markov.bin"), "-o", str(tmp_path / "h.tdc"), "--force"],
                       capture_output=True, text=True, env=dict(env, TDCGPU_HOST_ENCODE="1"))
    assert r.returncode == 0, r.stderr
    assert _run(REF, "lzss_lcp(coder=huff)", str(tmp_path / "markov.bin"), str(tmp_path / "ref.tdc")).returncode == 0
    assert open(tmp_path / "h.tdc", "rb").read() == open(tmp_path / "ref.tdc", "rb").read()


@pytest.mark.sim
def test_plugin_in_a_wide_index_build_of_the_reference(tmp_path):
    """-DLEN_BITS=40 (def.hpp:100-114: len_t becomes 64-bit, the archive's length field 64 bits wide): the plugin compiles
    against that build and LZSSLCPCompressor<coder, GpuTextDS> (over the interpreter library) writes the archive that the
    same build's CPU TextDS<> writes — 4 bytes longer than the default build's."""
    wide, narrow = os.path.join(ROOT, "build", "tdc_plugin_bench40"), os.path.join(ROOT, "build", "tdc_plugin_bench")
    if not (os.path.exists(wide) and os.path.exists(narrow)):
        pytest.skip("tdc_plugin_bench40 not built (bash tudocomp_b200/plugin/build_tdc.sh; needs /root/reference)")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    simdir = tmp_path / "simlib"
    simdir.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so"), simdir / "libtdcgpu.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(simdir))
    import json
    for name, data in (("markov", synth.markov_text(9000, 5)[:-1].tobytes()), ("dna", synth.dna(5000, 6)[:-1].tobytes()), ("one", b"a")):
        src = tmp_path / f"{name}.bin"
        src.write_bytes(data)
        for coder in ("huff", "bit"):
            rec = {}
            for exe, gpu in ((wide, "1"), (wide, "0"), (narrow, "1")):
                r = subprocess.run([exe, str(src), coder, "3", "1", gpu, str(tmp_path / "o.bin")], capture_output=True, text=True, env=env)
                assert r.returncode == 0, (name, coder, r.stderr)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                rec[(exe, gpu)] = (d["archive_bytes"], d["archive_fnv1a"])
            assert rec[(wide, "1")] == rec[(wide, "0")], (name, coder, rec)
            assert rec[(wide, "1")][0] == rec[(narrow, "1")][0] + 4, (name, coder, rec)


PROVIDERS_CHECK = os.path.join(ROOT, "build", "tdc_providers_check")


def _providers_cases(tmp_path, big):
    cases = {"markov": synth.markov_text(200000 if big else 9000, 5)[:-1].tobytes(), "dna": synth.dna(150000 if big else 5000, 6)[:-1].tobytes(),
             "repetitive": synth.repetitive(100000 if big else 6000, 7, block=700, p=0.01)[:-1].tobytes(),
             "binary_with_escapes": bytes(np.random.default_rng(3).integers(0, 256, 20000 if big else 3000, dtype=np.uint8)),
             "run": b"a" * 2000, "one": b"a", "empty": b""}
    for k, v in cases.items():
        p = tmp_path / f"prov_{k}.bin"
        p.write_bytes(v)
        yield k, str(p)


@pytest.mark.sim
def test_gpu_providers_equal_reference_providers_over_the_simulator_library(tmp_path):
    """TextDS<GpuSA, GpuPhi, GpuPLCP, GpuLCP, GpuISA> == TextDS<> array by array (values, widths, max_lcp) in every
    CompressMode (plugin/tdc_providers_check.cpp), with the interpreter library in place of libtdcgpu.so."""
    if not os.path.exists(PROVIDERS_CHECK):
        pytest.skip("tdc_providers_check not built (bash tudocomp_b200/plugin/build_tdc.sh; needs /root/reference)")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    simdir = tmp_path / "simlib"
    simdir.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so"), simdir / "libtdcgpu.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(simdir))
    for name, src in _providers_cases(tmp_path, big=False):
        r = subprocess.run([PROVIDERS_CHECK, src], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and '"providers_equal": true' in r.stdout, (name, r.stdout, r.stderr)


@pytest.mark.gpu
def test_gpu_providers_equal_reference_providers(tmp_path):
    if not os.path.exists(PROVIDERS_CHECK):
        pytest.skip("tdc_providers_check not built")
    for name, src in _providers_cases(tmp_path, big=True):
        r = subprocess.run([PROVIDERS_CHECK, src], capture_output=True, text=True)
        assert r.returncode == 0 and '"providers_equal": true' in r.stdout, (name, r.stdout, r.stderr)


def _block_container(path):
    """(block_bytes, algo, [archive bytes per block]) of a tdc_block container (tudocomp_b200/plugin/tdc_block.cpp):
    header, archives in completion order, index (offset, length per block, in block order), index offset, end mark."""
    import struct
    c = open(path, "rb").read()
    assert c.startswith(b"TDCBLOCK2\n") and c.endswith(b"TDCBEND\n")
    p = 10
    blk, nb = struct.unpack_from("<QQ", c, p)
    p += 16
    (al,) = struct.unpack_from("<I", c, p)
    p += 4
    algo = c[p:p + al].decode()
    p += al
    (index_off,) = struct.unpack_from("<Q", c, len(c) - 16)
    assert index_off + 16 * nb + 16 == len(c)
    arcs, covered = [], 0
    for b in range(nb):
        off, ln = struct.unpack_from("<QQ", c, index_off + 16 * b)
        assert p <= off and off + ln <= index_off
        arcs.append(c[off:off + ln])
        covered += ln
    assert covered == index_off - p  # the archives tile the space between header and index
    return blk, algo, arcs


def _need_block_bins():
    if not all(os.path.exists(p) for p in (REF, BLOCK_REF, BLOCK_GPU)):
        pytest.skip("tdc_block drivers not built (bash tudocomp_b200/plugin/build_tdc.sh; needs /root/reference)")


@pytest.mark.sim
def test_block_mode_driver_over_the_simulator_library(tmp_path):
    """Block mode (config 5) end to end on the CPU: tdc_block over the GPU registry (interpreter library in place of
    libtdcgpu.so) produces the same container as tdc_block over the reference registry, every block equals what the stock
    reference driver writes with --raw for that slice, forked workers give the same container, and it round-trips."""
    _need_block_bins()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    simdir = tmp_path / "simlib"
    simdir.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so"), simdir / "libtdcgpu.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(simdir))
    data = (synth.markov_text(5000, 5)[:-1].tobytes() + bytes(np.random.default_rng(1).integers(0, 256, 1500, dtype=np.uint8))
            + synth.dna(3000, 6)[:-1].tobytes())
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    algo, blk = "lzss_lcp(coder=huff)", 3000
    ref_c, gpu_c, ref3_c = (str(tmp_path / n) for n in ("ref.tdcb", "gpu.tdcb", "ref3.tdcb"))
    assert subprocess.run([BLOCK_REF, "-a", algo, "-b", str(blk), str(src), "-o", ref_c], capture_output=True).returncode == 0
    r = subprocess.run([BLOCK_GPU, "-a", algo, "-b", str(blk), str(src), "-o", gpu_c], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert open(ref_c, "rb").read() == open(gpu_c, "rb").read()
    # -c: one device context kept alive across the blocks of a worker (TDCGPU_CTX_CACHE=1): same bytes
    r = subprocess.run([BLOCK_GPU, "-a", algo, "-b", str(blk), "-c", str(src), "-o", gpu_c + ".c"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert open(ref_c, "rb").read() == open(gpu_c + ".c", "rb").read()
    # forked workers: the same archives per block (their order inside the container is the order of completion), decodable
    assert subprocess.run([BLOCK_REF, "-a", algo, "-b", str(blk), "-g", "3", str(src), "-o", ref3_c], capture_output=True).returncode == 0
    assert _block_container(ref3_c) == _block_container(ref_c)
    assert subprocess.run([BLOCK_REF, "-d", ref3_c, "-o", str(tmp_path / "back3.bin")], capture_output=True).returncode == 0
    assert open(tmp_path / "back3.bin", "rb").read() == data
    b, a, arcs = _block_container(gpu_c)
    assert (b, a, len(arcs)) == (blk, algo, -(-len(data) // blk))
    for i, arc in enumerate(arcs):
        sl = tmp_path / "slice.bin"
        sl.write_bytes(data[i * blk:(i + 1) * blk])
        assert _run(REF, algo, str(sl), str(tmp_path / "slice.tdc"), ["--raw"]).returncode == 0
        assert open(tmp_path / "slice.tdc", "rb").read() == arc, i
    back = str(tmp_path / "back.bin")
    assert subprocess.run([BLOCK_REF, "-d", gpu_c, "-o", back], capture_output=True).returncode == 0
    assert open(back, "rb").read() == data
    # damaged containers are refused (exit code 1, "Error: ..."), never decoded into something else
    good = open(gpu_c, "rb").read()
    for name, bad in (("truncated", good[:-9]), ("no_end_mark", good[:-1] + b"X"), ("index_offset_past_the_end", good[:-16] + (2 ** 40).to_bytes(8, "little") + good[-8:]),
                      ("not_a_container", b"hello world, this is not a container at all")):
        p = tmp_path / f"bad_{name}.tdcb"
        p.write_bytes(bad)
        r = subprocess.run([BLOCK_REF, "-d", str(p), "-o", str(tmp_path / "bad.out")], capture_output=True, text=True)
        assert r.returncode == 1 and "Error" in r.stderr, (name, r.returncode, r.stderr)


@pytest.mark.sim
def test_block_mode_block_size_edge_cases_over_the_simulator_library(tmp_path):
    """Block sizes from 1 byte to larger than the file, data with bytes that need escaping: the GPU registry's container (interpreter
    library) equals the reference registry's and decodes to the input."""
    _need_block_bins()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tudocomp_b200", "csrc"), "sim"])
    simdir = tmp_path / "simlib"
    simdir.mkdir()
    os.symlink(os.path.join(ROOT, "tests", "sim", "_build", "libtdcsim.so"), simdir / "libtdcgpu.so")
    env = dict(os.environ, LD_LIBRARY_PATH=str(simdir))
    rng = np.random.default_rng(12)
    data = bytes(rng.integers(0, 256, 3000, dtype=np.uint8)) + b"a" * 500 + bytes(rng.integers(97, 100, 2000, dtype=np.uint8))
    for blk in (1, 7, 999, 5500, 100000):
        d = data[:40] if blk == 1 else (data[2990:3200] if blk == 7 else data)  # (few blocks: every block is a whole interpreter run)
        src = tmp_path / f"in{blk}.bin"
        src.write_bytes(d)
        g, r_ = str(tmp_path / "g.tdcb"), str(tmp_path / "r.tdcb")
        rr = subprocess.run([BLOCK_GPU, "-a", "lzss_lcp(coder=huff)", "-b", str(blk), str(src), "-o", g], capture_output=True, text=True, env=env)
        assert rr.returncode == 0, (blk, rr.stderr)
        assert subprocess.run([BLOCK_REF, "-a", "lzss_lcp(coder=huff)", "-b", str(blk), str(src), "-o", r_], capture_output=True).returncode == 0
        assert open(g, "rb").read() == open(r_, "rb").read(), blk
        back = str(tmp_path / "back.bin")
        assert subprocess.run([BLOCK_REF, "-d", g, "-o", back], capture_output=True).returncode == 0
        assert open(back, "rb").read() == d, blk


def _inputs(tmp_path):
    cases = {
        "markov": synth.markov_text(300000, 5)[:-1].tobytes(),
        "dna": synth.dna(300000, 6)[:-1].tobytes(),
        "repetitive": synth.repetitive(200000, 7, block=3000, p=0.01)[:-1].tobytes(),
        "binary_with_escapes": bytes(np.random.default_rng(3).integers(0, 256, 50000, dtype=np.uint8)),
        "empty": b"",
        "one": b"a",
    }
    out = {}
    for k, v in cases.items():
        p = tmp_path / f"{k}.bin"
        p.write_bytes(v)
        out[k] = str(p)
    return out


@pytest.mark.gpu
def test_archives_byte_identical_to_reference_driver(tmp_path):
    _need_bins()
    for name, src in _inputs(tmp_path).items():
        for coder in ("bit", "huff", "ascii"):
            for thr in (3, 5):
                a, b, c = (str(tmp_path / f"{name}.{coder}.{thr}.{x}") for x in ("ref", "gpu", "only"))
                cpu_algo = f"lzss_lcp(coder={coder},threshold={thr})"
                assert _run(REF, cpu_algo, src, a, ["--raw"]).returncode == 0
                r = _run(GPU, f"lzss_lcp(coder={coder},textds=gpu,threshold={thr})", src, b, ["--raw"])
                assert r.returncode == 0, r.stderr
                assert open(a, "rb").read() == open(b, "rb").read(), (name, coder, thr)
                # GPU-only registry: identical -a string, so the archive INCLUDING the driver header is identical
                assert _run(REF, cpu_algo, src, a).returncode == 0
                r = _run(GPU_ONLY, cpu_algo, src, c)
                assert r.returncode == 0, r.stderr
                assert open(a, "rb").read() == open(c, "rb").read(), (name, coder, thr, "with header")
        # decompress the GPU-built archive with the unmodified reference driver
        back = str(tmp_path / f"{name}.back")
        r = subprocess.run([REF, "-d", c, "-o", back, "--force"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(back, "rb").read() == open(src, "rb").read(), name


@pytest.mark.gpu
def test_per_provider_gpu_classes(tmp_path):
    """`textds(sa=gpu, ...)`: GPU providers inside the reference's TextDS, the rest of the pipeline is the reference's CPU code."""
    _need_bins()
    for name, src in _inputs(tmp_path).items():
        for algo, sel in (("lzss_lcp(coder=huff)", "lzss_lcp(coder=huff, textds=textds(sa=gpu, lcp=gpu, isa=gpu))"),
                          ("lzss_lcp(coder=bit,threshold=5)", "lzss_lcp(coder=bit, threshold=5, textds=textds(sa=gpu))"),
                          ("bwt", "bwt(textds=textds(sa=gpu))")):
            a, b = str(tmp_path / "ref.tdc"), str(tmp_path / "gpu.tdc")
            assert _run(REF, algo, src, a, ["--raw"]).returncode == 0
            assert _run(GPU, sel, src, b, ["--raw"]).returncode == 0, (name, sel)
            assert open(a, "rb").read() == open(b, "rb").read(), (name, sel)


@pytest.mark.gpu
def test_bwt_chain_byte_identical(tmp_path):
    _need_bins()
    for name, src in _inputs(tmp_path).items():
        a, b = str(tmp_path / f"{name}.bwt.ref"), str(tmp_path / f"{name}.bwt.gpu")
        algo = "bwt:mtf:rle:encode(huff)"
        assert _run(REF, algo, src, a).returncode == 0
        r = _run(GPU_ONLY, algo, src, b)
        assert r.returncode == 0, r.stderr
        assert open(a, "rb").read() == open(b, "rb").read(), name
        if name == "binary_with_escapes":
            continue  # the reference's own bwt chain does not round-trip escaped binary input (identical archives above)
        back = str(tmp_path / f"{name}.bwt.back")
        assert subprocess.run([REF, "-d", b, "-o", back, "--force"], capture_output=True).returncode == 0
        assert open(back, "rb").read() == open(src, "rb").read(), name


@pytest.mark.gpu
def test_gpu_textds_arrays_through_stats(tmp_path):
    """The GPU provider logs the same phase titles / bit widths the reference providers log."""
    _need_bins()
    src = tmp_path / "m.txt"
    src.write_bytes(synth.markov_text(100000, 8)[:-1].tobytes())
    r = subprocess.run([GPU_ONLY, "-a", "lzss_lcp(coder=bit)", str(src), "-o", str(tmp_path / "o"), "--force", "--stats"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for title in ("Construct Text DS", "Factorize", "Encode", "factors", "threshold", "gpu_ms:Encode: bit stream"):
        assert title in r.stdout, title


@pytest.mark.gpu
def test_lcpcomp_with_gpu_text_index(tmp_path):
    """lcpcomp's strategies (compressors/lcpcomp/compress/*.hpp) on arrays built by the GPU, byte-identical
    raw archives, and the reference driver decodes them."""
    _need_bins()
    for name, src in _inputs(tmp_path).items():
        for opts in ("coder=huff", "coder=huff,comp=heap", "coder=ascii,comp=max_lcp", "coder=huff,comp=plcppeaks,dec=compact"):
            a, b = str(tmp_path / f"{name}.lcpcomp.ref"), str(tmp_path / f"{name}.lcpcomp.gpu")
            assert _run(REF, f"lcpcomp({opts})", src, a, ["--raw"]).returncode == 0
            r = _run(GPU, f"lcpcomp({opts},textds=gpu)", src, b, ["--raw"])
            assert r.returncode == 0, r.stderr
            assert open(a, "rb").read() == open(b, "rb").read(), (name, opts)


@pytest.mark.gpu
def test_host_and_device_encode_agree(tmp_path):
    """TDCGPU_HOST_ENCODE=1 keeps the reference's encode_text on the host (factor list copied back); the default encodes
    on the device.  Same bytes."""
    _need_bins()
    src = tmp_path / "m.txt"
    src.write_bytes(synth.markov_text(400000, 9)[:-1].tobytes())
    for coder in ("huff", "bit"):
        a, b = str(tmp_path / "dev.tdc"), str(tmp_path / "host.tdc")
        assert _run(GPU_ONLY, f"lzss_lcp(coder={coder})", str(src), a).returncode == 0
        r = subprocess.run([GPU_ONLY, "-a", f"lzss_lcp(coder={coder})", str(src), "-o", b, "--force"], capture_output=True, text=True,
                           env=dict(os.environ, TDCGPU_HOST_ENCODE="1"))
        assert r.returncode == 0, r.stderr
        assert open(a, "rb").read() == open(b, "rb").read(), coder


@pytest.mark.gpu
def test_block_mode_driver_on_the_gpu(tmp_path):
    """tdc_block over the GPU registry: same container as over the reference registry, round trip through the reference."""
    _need_block_bins()
    data = synth.markov_text(3 << 20, 11)[:-1].tobytes()
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    algo = "lzss_lcp(coder=huff)"
    ref_c, gpu_c = str(tmp_path / "ref.tdcb"), str(tmp_path / "gpu.tdcb")
    assert subprocess.run([BLOCK_REF, "-a", algo, "-b", str(1 << 20), str(src), "-o", ref_c], capture_output=True).returncode == 0
    r = subprocess.run([BLOCK_GPU, "-a", algo, "-b", str(1 << 20), str(src), "-o", gpu_c], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(ref_c, "rb").read() == open(gpu_c, "rb").read()
    back = str(tmp_path / "back.bin")
    assert subprocess.run([BLOCK_REF, "-d", gpu_c, "-o", back], capture_output=True).returncode == 0
    assert open(back, "rb").read() == data

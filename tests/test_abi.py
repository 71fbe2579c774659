"""CPU tests of the drop-in boundary: the CUDA library loads, exports every symbol include/tdcgpu.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import tudocomp_b200 as tdc
from tudocomp_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tdcgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tdcgpu_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert len(syms) >= 15
    from tudocomp_b200 import dist as tdist

    assert sorted(syms) == sorted(_abi.EXPORTS + tdist.DIST_EXPORTS)
    lib = ctypes.CDLL(tdc.lib_path())
    for s in syms:
        assert hasattr(lib, s), s


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "tdcgpu.h")).read()
    for cite in ("ds/TextDS.hpp:247-292", "LZSSLCPCompressor.hpp:60-115", "ds/bwt.hpp:19-22", "LZSSFactors.hpp:13-20"):
        assert cite in hdr


def test_no_cpu_fallback_without_device():
    lib = tdc.load()
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tdc.TdcGpuError) as e:
        tdc.Context(lib, 0)
    assert e.value.code == -1
    with pytest.raises(tdc.TdcGpuError):
        tdc.TextDS(np.frombuffer(b"banana\0", np.uint8))


def test_multi_gpu_entry_points_fail_loudly_without_device():
    from tudocomp_b200 import dist as tdist

    lib = tdc.load()
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(tdc.TdcGpuError) as e:
        tdist.DistContext.create_nccl(lib, 0, None)
    assert e.value.code == -1 and "no CUDA device" in str(e.value)


def test_missing_sentinel_is_rejected_like_the_reference():
    with pytest.raises(ValueError, match="Input has no sentinel"):
        tdc.TextDS(np.frombuffer(b"banana", np.uint8))


def test_product_never_imports_oracle_or_simulator():
    # the package must not reference the test-only libraries
    pkg = os.path.join(ROOT, "tudocomp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libtdcoracle" not in src and "libtdcsim" not in src and "libtdcref" not in src, f

"""CPU tests: the C-ABI library loads and exports every symbol include/ompmc_b200.h declares; struct
layouts of the ctypes binding match the compiled library.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from ompmc_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return api.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ompmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omc_gpu_\w+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ompmc_b200.h but not exported"
    assert sorted(api.EXPORTS) == syms, "python binding and header disagree on the entry points"


def test_struct_layouts(lib):
    assert lib.omc_gpu_abi_sizeof(0) == C.sizeof(api.MediaTables)
    assert lib.omc_gpu_abi_sizeof(1) == C.sizeof(api.Geometry)
    assert lib.omc_gpu_abi_sizeof(2) == C.sizeof(api.SourceDosxyz)
    assert lib.omc_gpu_abi_sizeof(3) == C.sizeof(api.SourceMatrad)
    assert lib.omc_gpu_abi_sizeof(4) == api.RECORD_DTYPE.itemsize
    assert lib.omc_gpu_abi_sizeof(5) == C.sizeof(api.Counters)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to construct a context (and the oracle is never linked)."""
    import subprocess
    out = subprocess.run(["nm", "-D", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_run_histories" not in out
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.omc_gpu_create(C.byref(h), 0) != 0
    with pytest.raises(api.OmcGpuError):
        api.GpuTransport(0)


def test_sm100a_cubin_present():
    out = __import__("subprocess").run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_multi_gpu_entry_points_without_a_device():
    """The multi-GPU layer has no CPU path either: omc_gpu_multi_create() reports "no CUDA device" (rc 3) where there is none;
    the host-only helpers work without one (sharding rule, NCCL id when an NCCL is installed)."""
    import ctypes as C
    import torch
    lib = api.load_library()
    lo, cnt = C.c_longlong(0), C.c_longlong(0)
    assert lib.omc_gpu_shard_range(100, 10, 3, 4, C.byref(lo), C.byref(cnt)) == 0 and (lo.value, cnt.value) == (108, 2)
    assert lib.omc_gpu_shard_range(0, 10, 4, 4, C.byref(lo), C.byref(cnt)) != 0          # rank out of range
    assert lib.omc_gpu_shard_range(0, -1, 0, 1, C.byref(lo), C.byref(cnt)) != 0
    if not torch.cuda.is_available():
        m = C.c_void_p()
        assert lib.omc_gpu_multi_create(C.byref(m), 2, None) == 3 and not m.value
        with __import__("pytest").raises(api.OmcGpuError):
            api.MultiGpuTransport(ndev=2)

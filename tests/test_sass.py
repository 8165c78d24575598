"""Static checks of the shipped sm_100a code (cuobjdump -sass, no GPU needed): the design claims of BASELINE.json's north_star
that are visible in the instruction stream -- (b) queue records and coefficient pairs move with 128-bit loads / stores,
(d) dose is scored with fp32 reductions (REDG.E.ADD.F32: fire-and-forget, no return value) in the wavefront kernels and with fp64
atomics in the lock-step parity kernel, (c) the hot step kernels keep the Philox state in registers (no local-memory spill
traffic beyond the call frames of the rare __noinline__ samplers) -- and that no tensor-core or Hopper-only instruction appears."""
import re
import subprocess

import pytest

from ompmc_b200 import api, build

KERNELS = {
    "esize": "_ZN3omc12esize_kernelENS_10DevProblemENS_8WaveArgsE",
    "ech": "_ZN3omc10edo_kernelILi1EEEvNS_10DevProblemENS_8WaveArgsE",
    "ebca": "_ZN3omc10edo_kernelILi2EEEvNS_10DevProblemENS_8WaveArgsE",
    "misc": "_ZN3omc11misc_kernelENS_10DevProblemENS_8WaveArgsE",
    "lockstep": "_ZN3omc15lockstep_kernelENS_10DevProblemEPNS_4PartEixxiPKS1_",
    "format_e": "_ZN3omc13format_kernelILi0EEEvPKdxPKNS_5Pow10EPcPNS_14FormatFallbackEPjj",
}


@pytest.fixture(scope="module")
def sass():
    build.build()
    out = {}
    for name, sym in KERNELS.items():
        r = subprocess.run(["cuobjdump", "-sass", "-fun", sym, api.LIB_PATH], capture_output=True, text=True)
        assert "Function :" in r.stdout, f"kernel {name} ({sym}) not found in the library"
        out[name] = r.stdout
    return out


def count(text, pattern):
    return len(re.findall(pattern, text))


def test_wave_kernels_move_records_with_128_bit_accesses(sass):
    for k in ("esize", "ech", "ebca", "misc"):
        assert count(sass[k], r"LDG\.E\.128") >= 6, k          # 16-byte lanes of the queue records, {c1, c0} coefficient pairs
        assert count(sass[k], r"STG\.E\.128") >= 6, k


def test_dose_scoring_instructions(sass):
    for k in ("ech", "ebca", "esize", "misc"):
        assert count(sass[k], r"REDG\.E\.ADD\.F32") >= 1, k    # fp32 chunk grid, reduction without return value
        assert count(sass[k], r"ATOMG\.E\.ADD\.F32") == 0, k
    assert count(sass["lockstep"], r"REDG\.E\.ADD\.F64|ATOMG\.E\.ADD\.F64") >= 1      # parity kernel: fp64 atomics on the batch grid


def test_step_kernels_keep_their_state_in_registers(sass):
    # the condensed-history / boundary-crossing / step-size kernels: a few dozen STATIC local accesses at most (call frames of the
    # rare out-of-line samplers, plus the spills of the 80-register cap = 6 resident blocks per SM, measured 1.303e8 against
    # 1.243e8 histories/s with 96 registers and no spills); the first version of the CH kernel EXECUTED 12.6 % LDL + STL
    #  because its generator state lived in local memory
    for k, limit in (("ech", 48), ("ebca", 24), ("esize", 24)):
        n = count(sass[k], r"\b(LDL|STL)\b")
        total = count(sass[k], r"/\*[0-9a-f]{4}\*/")
        assert n <= limit, f"{k}: {n} local-memory instructions of {total}"


def test_no_tensor_core_or_foreign_instructions(sass):
    for k, text in sass.items():
        assert count(text, r"\b(HMMA|IMMA|DMMA|QGMMA|UTCHMMA|UTCQMMA|WGMMA|HGMMA)\b") == 0, k     # nothing here is a dense contraction


def test_formatter_uses_wide_integer_multiplies_and_vector_stores(sass):
    assert count(sass["format_e"], r"IMAD\.WIDE\.U32|IMAD\.HI\.U32|IMAD\.WIDE") >= 8       # 64 x 128-bit product of omc_format.cuh
    assert count(sass["format_e"], r"STG\.E\.128") >= 1                                    # staged records leave as 16-byte words
    assert count(sass["format_e"], r"\b(LDL|STL)\b") == 0

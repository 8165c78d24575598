"""The .3ddose text conversion (ompmc_b200/csrc/omc_format.cuh; outputResults(), omc_dosxyz.c:859-877).

CPU part: the same __host__ __device__ functions the formatting kernel runs, compiled for the host (libomc_format_host.so),
against glibc's snprintf("%e ") / snprintf("%f ") -- every value the device certifies must be byte-identical, and the values it
refuses (rounding ties, negative / non-finite numbers, other field widths) must be flagged, never guessed.
GPU part: omc_gpu_test_format (one block of a file through format_kernel + the host splice of the refused values) and
omc_gpu_write_3ddose against the per-value fprintf loop of the reference's writer."""
import ctypes as C
import os

import numpy as np
import pytest

from ompmc_b200 import build, problem as P

libc = C.CDLL(None)
libc.snprintf.restype = C.c_int


def c_format(spec: bytes, values) -> bytes:
    """what `for v in values: fprintf(fp, spec, v)` writes (glibc itself)"""
    buf = C.create_string_buffer(512)
    out = bytearray()
    for v in values:
        n = libc.snprintf(buf, C.c_size_t(512), spec, C.c_double(float(v)))
        out += buf.raw[:n]
    return bytes(out)


def py_format(spec: str, values) -> bytes:
    """CPython's % formatting is correctly rounded (round-half-even) like glibc for finite numbers; fast path for bulk data"""
    return "".join([spec % v for v in values.tolist()]).encode()


def edge_values_e():
    v = [0.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1e-99, 9.9999995e-100, 1e99, 9.9999994e99, 1e100, 1e-100,
         -1.0, -0.0, float("nan"), float("inf"), -float("inf"), 1.0, 10.0, 0.1, 0.5, 123456.5, 1234567.5, 12345675.0, 1000000.5,
         9999999.5, 99999995.0, 9.9999995, 9.9999994999, 1.602e-10, 6.25e-17]
    for k in range(-30, 10):
        x = 10.0 ** k
        v += [x, np.nextafter(x, 0), np.nextafter(x, np.inf), 9.9999995 * x, 1.0000005 * x, 2.5 * x]
    v += [i + 0.5 for i in range(1000000, 1000200)]                 # exact ties at the 7th digit
    v += [(2 * i + 1) * 2.0 ** -s for i in range(40) for s in range(1, 60, 3)]
    return np.array(v, dtype=np.float64)


def edge_values_f():
    v = [0.0, 0.9999999, 1.0, 0.5, 0.25, 0.0078125, 0.0234375, 9.9999994, 9.9999995, 9.99999951, 10.0, 15.9, 16.0, 1e10, 5e-324, 1e-7,
         4.9999999e-7, 5e-7, 5.0000001e-7, -0.5, -0.0, float("nan"), float("inf"), 0.3333333333, 0.9999995, 0.99999949]
    v += [2.0 ** -s for s in range(1, 80)] + [3 * 2.0 ** -s for s in range(2, 40)] + [i / 128.0 for i in range(1280)]
    return np.array(v, dtype=np.float64)


@pytest.fixture(scope="module")
def hostfmt():
    build.build_host()
    lib = C.CDLL(build.FORMAT_LIB)
    lib.omc_format_host.restype = C.c_longlong
    lib.omc_format_host.argtypes = [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]

    def run(mode, values):
        values = np.ascontiguousarray(values, dtype=np.float64)
        w = lib.omc_format_width(mode)
        text = np.zeros(values.size * w, dtype=np.uint8)
        flags = np.zeros(values.size, dtype=np.uint8)
        n = lib.omc_format_host(mode, values.size, values.ctypes.data, text.ctypes.data, flags.ctypes.data)
        assert n == int(flags.sum())
        return text.reshape(values.size, w), flags.astype(bool)
    return run


def check_against_glibc(run, mode, values, bulk):
    spec = "%e " if mode == 0 else "%f "
    text, flags = run(mode, values)
    w = text.shape[1]
    ok = values[~flags]
    want = py_format(spec, ok) if bulk else c_format(spec.encode(), ok)
    assert len(want) == ok.size * w, "a certified value has another width in glibc"
    assert text[~flags].tobytes() == want
    return flags


def test_e_format_equals_glibc_on_dose_like_and_random_values(hostfmt):
    rng = np.random.default_rng(7)
    dose = rng.random(400000) * 10.0 ** rng.uniform(-22, -6, 400000)
    assert not check_against_glibc(hostfmt, 0, dose, True).any()            # nothing refused on ordinary data
    bits = rng.integers(0, 2 ** 63, 300000, dtype=np.uint64).view(np.float64)
    bits = bits[np.isfinite(bits)]
    flags = check_against_glibc(hostfmt, 0, bits, True)
    small = (bits >= 1e-99) & (bits < 1e100)
    assert not flags[small].any() and flags[~small & (bits != 0)].all()       # only three-digit exponents are left to the host


def test_e_format_edge_cases_are_exact_or_refused(hostfmt):
    v = edge_values_e()
    flags = check_against_glibc(hostfmt, 0, v, False)
    assert flags[np.isnan(v) | np.isinf(v) | np.signbit(v)].all()
    ties = np.array([i + 0.5 for i in range(1000000, 1000200)])
    assert hostfmt(0, ties)[1].all()                                        # exact ties: never guessed
    assert not hostfmt(0, np.array([0.0, 1.0, 5e-100 * 2, 9.9999994e99, 1.602e-10]))[1].any()


def test_f_format_is_exact_including_ties(hostfmt):
    rng = np.random.default_rng(11)
    v = np.concatenate([rng.random(300000), rng.random(100000) * 9.9999994, rng.random(100000) * 10.0 ** rng.uniform(-12, 0, 100000),
                        rng.integers(0, 2 ** 27, 200000) / 2.0 ** 24])      # dyadic rationals: exact ties at the 6th decimal occur
    assert not check_against_glibc(hostfmt, 1, v, True).any()
    e = edge_values_f()
    flags = check_against_glibc(hostfmt, 1, e, False)
    refused = np.array([not np.isfinite(x) or np.signbit(x) or len("%f" % x) != 8 for x in e])       # other width or not a number
    assert np.array_equal(flags, refused)
    assert hostfmt(1, np.array([0.0078125]))[0].tobytes() == b"0.007812 " and hostfmt(1, np.array([0.0234375]))[0].tobytes() == b"0.023438 "


# ---- GPU ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_device_formatter_block_equals_fprintf_loop(gpu, tmp_path, mode):
    rng = np.random.default_rng(3 + mode)
    spec = "%e " if mode == 0 else "%f "
    edge = edge_values_e() if mode == 0 else edge_values_f()
    body = rng.random(300000) * (10.0 ** rng.uniform(-22, -6, 300000) if mode == 0 else 1.0)
    v = np.concatenate([edge, body, edge[::-1], body[:1000] * 0.0, edge])          # refused values at both ends and in the middle
    path = str(tmp_path / "blk.txt")
    gpu.test_format(mode, v, path)
    finite = np.isfinite(v)
    want = bytearray()
    i = 0
    while i < v.size:                                                             # glibc for the specials, CPython for the bulk
        if finite[i]:
            j = i
            while j < v.size and finite[j]:
                j += 1
            want += py_format(spec, v[i:j]); i = j
        else:
            want += c_format(spec.encode(), v[i:i + 1]); i += 1
    assert open(path, "rb").read() == bytes(want) + b"\n"
    for n in (0, 1, 255, 256, 257):                                                # empty and ragged blocks
        gpu.test_format(mode, body[:n], path)
        assert open(path, "rb").read() == py_format(spec, body[:n]) + b"\n"


@pytest.mark.gpu
def test_device_formatter_streams_several_chunks(gpu, tmp_path):
    """more values than one 4 Mi chunk: double-buffered kernel / copy / fwrite, the last chunk ragged, a refused value in each"""
    n = (4 << 20) + 70001
    rng = np.random.default_rng(5)
    v = rng.random(n) * 1e-12
    v[[17, (4 << 20) - 1, 4 << 20, n - 1]] = [-1.0, 1000000.5, float("nan"), 1e-120]
    path = str(tmp_path / "big.txt")
    gpu.test_format(0, v, path)
    got = open(path, "rb").read()
    idx = [0, 17, 18, (4 << 20) - 1, 4 << 20, (4 << 20) + 1, n - 1, n]
    want = bytearray()
    for a, b in zip(idx[:-1], idx[1:]):
        want += c_format(b"%e ", v[a:b]) if b - a == 1 else py_format("%e ", v[a:b])
    assert got == bytes(want) + b"\n"


@pytest.mark.gpu
def test_write_3ddose_on_device_equals_reference_writer_format(gpu, tmp_path):
    from oracle.gen_fixtures import golden_problem
    prob, ph, cfg = golden_problem("golden_tissue4_6MV")
    gpu.load_problem(prob)
    gpu.set_option("kernel", 1)
    gpu.reset_tallies()
    nhist, nb, nper = P.batch_plan("60000", "6")
    for ib in range(nb):
        gpu.run_batch(ib * nper, nper)
    dose, unc = gpu.accumulate_results(ph.med_densities, nper, nb)
    path = str(tmp_path / "dev.3ddose")
    gpu.write_3ddose(path, ph.med_densities, nper, nb)
    ref = str(tmp_path / "host.3ddose")
    P.write_3ddose(ref, ph, dose, unc)                                             # the reference's formats, value by value
    assert open(path, "rb").read() == open(ref, "rb").read()
    assert (dose > 0).sum() > 1000 and (unc == 0.9999999).any()
    with pytest.raises(Exception):
        gpu.write_3ddose(str(tmp_path / "no such dir" / "x.3ddose"), ph.med_densities, nper, nb)

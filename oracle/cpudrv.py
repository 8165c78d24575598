"""ctypes driver for the two CPU checkers.  TEST INFRASTRUCTURE (oracle/).

* ``RefTransport``    -> oracle/_ref/libompmc_ref[_omp].so : the unmodified reference + harness
                         (oracle/ref_harness.c)
* ``OracleTransport`` -> oracle/libomc_oracle.so            : this repo's plain-C restatement
                         (oracle/omc_oracle.c)
Both export the same entry points under a different prefix (``ref_`` / ``orc_``).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class HistoryRecord(C.Structure):
    _fields_ = [("ndraws", C.c_uint), ("ir_start", C.c_int), ("ndeposit", C.c_uint), ("flags", C.c_uint),
                ("edep", C.c_double)]


RECORD_DTYPE = np.dtype([("ndraws", "<u4"), ("ir_start", "<i4"), ("ndeposit", "<u4"), ("flags", "<u4"), ("edep", "<f8")])


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class _CpuTransport:
    prefix = ""

    def __init__(self, libpath: str):
        if not os.path.exists(libpath):
            raise FileNotFoundError(libpath)
        self.lib = C.CDLL(libpath)
        self.path = libpath
        f = self._f
        f("load_problem").argtypes = [C.c_char_p]; f("load_problem").restype = C.c_int
        f("set_rng").argtypes = [C.c_int, C.c_int, C.c_int]
        f("set_nsplit").argtypes = [C.c_int]
        f("nreg").restype = C.c_int
        f("run_histories").argtypes = [C.c_longlong, C.c_longlong, C.c_void_p]
        f("get_endep").argtypes = [C.c_void_p]
        f("get_accum").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        f("time_batches").argtypes = [C.c_longlong, C.c_longlong, C.c_int]; f("time_batches").restype = C.c_double
        f("num_threads").restype = C.c_int
        f("set_num_threads").argtypes = [C.c_int]
        f("test_geometry").argtypes = [C.c_int] + [C.c_void_p] * 7
        f("test_rng").argtypes = [C.c_longlong, C.c_int, C.c_void_p]
        f("run_particle").argtypes = [C.c_longlong, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        f("test_samplers").argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        try:
            f("test_photon_tau").argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        except AttributeError:      # the matRad build of the reference harness has no dosxyz geometry hook
            pass

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- problem ---------------------------------------------------------------------------
    def load_problem(self, problem, seeds=(97, 33)):
        """problem: path of a full blob, or a dict as built by ompmc_b200.problem.build_problem()."""
        self._f("set_rng")(1, int(seeds[0]), int(seeds[1]))
        if isinstance(problem, dict):
            from ompmc_b200.problem import save_blob
            fd, path = tempfile.mkstemp(suffix=".blob"); os.close(fd)
            try:
                save_blob(path, problem)
                rc = self._f("load_problem")(path.encode())
            finally:
                os.unlink(path)
        else:
            rc = self._f("load_problem")(str(problem).encode())
        if rc != 0:
            raise RuntimeError(f"{self.prefix}load_problem failed rc={rc}")
        self.nreg = self._f("nreg")()

    def set_rng(self, mode: str, seeds=(97, 33)):
        self._f("set_rng")({"ranmar": 0, "philox": 1}[mode], int(seeds[0]), int(seeds[1]))

    def set_nsplit(self, n: int):
        self._f("set_nsplit")(int(n))

    def set_beamlet(self, ibeamlet: int):
        fn = self._f("set_beamlet"); fn.argtypes = [C.c_int]
        fn(int(ibeamlet))

    # -- hot path --------------------------------------------------------------------------
    def run_histories(self, first: int, n: int, records: bool = False):
        rec = np.zeros(n, dtype=RECORD_DTYPE) if records else None
        self._f("run_histories")(first, n, rec.ctypes.data if records else None)
        return rec

    def accum_endep(self):
        self._f("accum_endep")()

    def reset_score(self):
        self._f("reset_score")()

    def get_endep(self) -> np.ndarray:
        out = np.zeros(self.nreg)
        self._f("get_endep")(out.ctypes.data)
        return out

    def set_endep(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        assert grid.size == self.nreg
        fn = self._f("set_endep"); fn.argtypes = [C.c_void_p]
        fn(grid.ctypes.data)

    def get_accum(self):
        a = np.zeros(self.nreg); a2 = np.zeros(self.nreg); e = C.c_double(0.0)
        self._f("get_accum")(a.ctypes.data, a2.ctypes.data, C.byref(e))
        return a, a2, e.value

    def time_batches(self, first: int, nperbatch: int, nbatch: int) -> float:
        return float(self._f("time_batches")(first, nperbatch, nbatch))

    def num_threads(self) -> int:
        return int(self._f("num_threads")())

    def set_num_threads(self, n: int):
        self._f("set_num_threads")(int(n))

    # -- unit hooks ------------------------------------------------------------------------
    def test_geometry(self, xyzuvw, ir, ustep_in):
        xyzuvw = np.ascontiguousarray(xyzuvw, dtype=np.float64); ir = np.ascontiguousarray(ir, dtype=np.int32)
        ustep_in = np.ascontiguousarray(ustep_in, dtype=np.float64)
        n = len(ir)
        idisc = np.zeros(n, np.int32); irnew = np.zeros(n, np.int32); us = np.zeros(n); tp = np.zeros(n)
        self._f("test_geometry")(n, xyzuvw.ctypes.data, ir.ctypes.data, ustep_in.ctypes.data, idisc.ctypes.data,
                                 irnew.ctypes.data, us.ctypes.data, tp.ctypes.data)
        return idisc, irnew, us, tp

    def test_rng(self, hist: int, n: int) -> np.ndarray:
        out = np.zeros(n)
        self._f("test_rng")(hist, n, out.ctypes.data)
        return out

    def test_samplers(self, which: int, inputs, first: int = 0) -> np.ndarray:
        """One sampler function on explicit inputs: n records of 8 doubles in -> n records of 8 doubles out
        (OMC_SAMPLER_* of include/ompmc_b200.h), record i drawing from the Philox stream of history first + i."""
        a = np.zeros((len(inputs), 8))
        inputs = np.asarray(inputs, dtype=np.float64)
        a[:, :inputs.shape[1]] = inputs
        out = np.zeros_like(a)
        self._f("test_samplers")(int(which), len(a), a.ctypes.data, int(first), out.ctypes.data)
        return out

    def test_photon_tau(self, e, xyzuvw, s) -> tuple[np.ndarray, np.ndarray]:
        """Optical depth of straight photon paths of length s (photon() src/ompmc.c:1951-2019) and the region at their end."""
        xyzuvw = np.asarray(xyzuvw, dtype=np.float64).reshape(-1, 6)
        n = len(xyzuvw)
        a = np.zeros((n, 8))
        a[:, 0] = e; a[:, 1:7] = xyzuvw; a[:, 7] = s
        out = np.zeros((n, 2))
        self._f("test_photon_tau")(n, a.ctypes.data, out.ctypes.data)
        return out[:, 0], out[:, 1].astype(np.int64)

    def run_particle(self, hist: int, iq: int, e: float, xyzuvw, ir: int, wt: float = 1.0):
        xyzuvw = np.ascontiguousarray(xyzuvw, dtype=np.float64)
        rec = np.zeros(1, dtype=RECORD_DTYPE)
        self._f("run_particle")(hist, iq, e, xyzuvw.ctypes.data, ir, wt, rec.ctypes.data)
        return rec[0]


class RefTransport(_CpuTransport):
    prefix = "ref_"

    def __init__(self, omp: bool = False, matrad: bool = False):
        super().__init__(os.path.join(HERE, "_ref", "libompmc_ref_matrad_omp.so" if omp else "libompmc_ref_matrad.so") if matrad
                         else ref_lib_path(omp))
        if not matrad:
            self.lib.ref_init_from_inp.argtypes = [C.c_char_p]
            self.lib.ref_dump_problem.argtypes = [C.c_char_p]

    def init_from_inp(self, stem: str):
        self.lib.ref_init_from_inp(stem.encode())
        self.nreg = self.lib.ref_nreg()

    def dump_problem(self, path: str):
        if self.lib.ref_dump_problem(path.encode()) != 0:
            raise RuntimeError("ref_dump_problem failed")


class OracleTransport(_CpuTransport):
    prefix = "orc_"

    def __init__(self):
        super().__init__(oracle_lib_path())
        self.lib.orc_test_ranmar.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.lib.orc_get_work.argtypes = [C.c_void_p]

    def test_ranmar(self, ixx: int, jxx: int, n: int) -> np.ndarray:
        out = np.zeros(n)
        self.lib.orc_test_ranmar(ixx, jxx, n, out.ctypes.data)
        return out

    def work_per_history(self) -> dict:
        """Per-history operation counts since reset_score() (SURVEY.md 8d's B_alg inputs)."""
        w = np.zeros(7, dtype=np.uint64)
        self.lib.orc_get_work(w.ctypes.data)
        n = max(int(w[6]), 1)
        names = ["ausgab", "howfar", "hownear", "pwlf", "mscat", "spin"]
        return {k: float(w[i]) / n for i, k in enumerate(names)}


def ref_lib_path(omp: bool = False) -> str:
    return os.path.join(HERE, "_ref", "libompmc_ref_omp.so" if omp else "libompmc_ref.so")


def oracle_lib_path() -> str:
    return os.path.join(HERE, "libomc_oracle.so")


def have_ref(omp: bool = False, matrad: bool = False) -> bool:
    if matrad:
        return os.path.exists(os.path.join(HERE, "_ref", "libompmc_ref_matrad_omp.so" if omp else "libompmc_ref_matrad.so"))
    return os.path.exists(ref_lib_path(omp))

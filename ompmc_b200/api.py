"""ctypes binding of the C-ABI (include/ompmc_b200.h) -- the reference-facing entry points.

``GpuTransport`` mirrors, call for call, what a reference user code does around its batch loop
(ucodes/omc_dosxyz/omc_dosxyz.c:1155-1283):

    init*()                       -> GpuTransport.load_problem(problem dict)
    for ibatch: {initHistory(); shower();} x nperbatch; accumEndep()
                                  -> run_batch(first_history, nperbatch)
    accumulateResults()/output    -> get_tallies() + ompmc_b200.problem.accumulate_results()

There is no CPU fallback: if the CUDA library is missing or no GPU is present, construction fails.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import problem as P

HERE = os.path.dirname(os.path.abspath(__file__))
BEAMLETS_PER_PASS, BEAMLET_GRID_BUDGET = 64, 64 * 1073741824        # OMC_BEAMLETS_PER_PASS, OMC_BEAMLET_GRID_BUDGET of the header
# (OMPMC_B200_LIB: an A/B build of the same library for the measurement scripts, e.g. built with OMC_NVCC_FLAGS; never a fallback)
LIB_PATH = os.environ.get("OMPMC_B200_LIB") or os.path.join(HERE, "libompmc_b200.so")

KERNEL_LOCKSTEP, KERNEL_WAVEFRONT = 0, 1
DEFAULT_KERNEL = KERNEL_WAVEFRONT    # production default used by bench.py (nsplit == 1)

RECORD_DTYPE = np.dtype([("ndraws", "<u4"), ("ir_start", "<i4"), ("ndeposit", "<u4"), ("flags", "<u4"), ("edep", "<f8")])
COUNTER_NAMES = ["histories", "kernel_launches", "photon_steps", "electron_steps", "deposits", "rng_draws", "errors"]

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)

_MEDIA_F64 = ["ge0", "ge1", "gmfp0", "gmfp1", "gbr10", "gbr11", "gbr20", "gbr21", "cohe0", "cohe1",
              "ray_xgrid", "ray_fcum", "ray_b_array", "ray_c_array"]


class MediaTables(C.Structure):
    """omc_media_tables -- field order must match include/ompmc_b200.h."""
    _fields_ = (
        [("nmed", C.c_int)]
        + [(n, PD) for n in ["ge0", "ge1", "gmfp0", "gmfp1", "gbr10", "gbr11", "gbr20", "gbr21", "cohe0", "cohe1"]]
        + [(n, PD) for n in ["ray_xgrid", "ray_fcum", "ray_b_array", "ray_c_array"]]
        + [("ray_i_array", PI)]
        + [(n, PD) for n in ["ray_pmax0", "ray_pmax1"]]
        + [(n, PD) for n in ["dl1", "dl2", "dl3", "dl4", "dl5", "dl6", "bpar0", "bpar1", "delcm", "zbrang"]]
        + [(n, PD) for n in ["esig0", "esig1", "psig0", "psig1", "ededx0", "ededx1", "pdedx0", "pdedx1",
                             "ebr10", "ebr11", "pbr10", "pbr11", "pbr20", "pbr21", "tmxs0", "tmxs1",
                             "blcce0", "blcce1", "etae_ms0", "etae_ms1", "etap_ms0", "etap_ms1",
                             "q1ce_ms0", "q1ce_ms1", "q1cp_ms0", "q1cp_ms1", "q2ce_ms0", "q2ce_ms1", "q2cp_ms0", "q2cp_ms1",
                             "range_ep", "e_array", "eke0", "eke1"]]
        + [("sig_ismonotone", PI)]
        + [(n, PD) for n in ["esig_e", "psig_e", "xcc", "blcc"]]
        + [(n, C.c_double) for n in ["b2spin_min", "dbeta2i", "espml", "dleneri", "dqq1i"]]
        + [("spin_rej", PD)]
        + [(n, PD) for n in ["ums", "fms", "wms"]]
        + [("ims", PI)]
        + [(n, C.c_double) for n in ["dllambi", "dqmsi"]]
        + [(n, PD) for n in ["pegs_ap", "pegs_ae", "pegs_te", "pegs_thmoll", "pegs_rho"]]
        + [("pegs_meke", PI)]
    )


class Geometry(C.Structure):
    _fields_ = [("isize", C.c_int), ("jsize", C.c_int), ("ksize", C.c_int), ("xbounds", PD), ("ybounds", PD), ("zbounds", PD),
                ("med", PI), ("rhof", PD), ("pcut", PD), ("ecut", PD)]


class SourceDosxyz(C.Structure):
    _fields_ = [("spectrum", C.c_int), ("charge", C.c_int), ("energy", C.c_double), ("deltak", C.c_double),
                ("cdfinv1", PD), ("cdfinv2", PD), ("ssd", C.c_double),
                ("xinl", C.c_double), ("xinu", C.c_double), ("yinl", C.c_double), ("yinu", C.c_double),
                ("xsize", C.c_double), ("ysize", C.c_double),
                ("ixinl", C.c_int), ("ixinu", C.c_int), ("iyinl", C.c_int), ("iyinu", C.c_int)]


class SourceMatrad(C.Structure):
    _fields_ = ([("spectrum", C.c_int), ("charge", C.c_int), ("energy", C.c_double), ("deltak", C.c_double),
                 ("cdfinv1", PD), ("cdfinv2", PD), ("nbeams", C.c_int), ("nbixels", C.c_int), ("ibeam", PI)]
                + [(n, PD) for n in ["xsource", "ysource", "zsource", "xcorner", "ycorner", "zcorner",
                                     "xside1", "yside1", "zside1", "xside2", "yside2", "zside2"]])


class Counters(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in COUNTER_NAMES] + [("reserved", C.c_ulonglong * 9)]


EXPORTS = ["omc_gpu_create", "omc_gpu_destroy", "omc_gpu_last_error", "omc_gpu_set_media", "omc_gpu_set_geometry",
           "omc_gpu_set_source_dosxyz", "omc_gpu_set_source_matrad", "omc_gpu_set_vrt", "omc_gpu_set_seed", "omc_gpu_set_option",
           "omc_gpu_run_histories", "omc_gpu_accum_batch", "omc_gpu_run_batch", "omc_gpu_start_batch", "omc_gpu_finish_batches",
           "omc_gpu_completed_batches", "omc_gpu_synchronize", "omc_gpu_get_tallies",
           "omc_gpu_get_batch_grid", "omc_gpu_accumulate_results", "omc_gpu_write_3ddose", "omc_gpu_test_format", "omc_gpu_run_beamlets", "omc_gpu_fetch_columns", "omc_gpu_reset_tallies", "omc_gpu_device_ptrs", "omc_gpu_stream", "omc_gpu_get_counters",
           "omc_gpu_get_history_records", "omc_gpu_test_geometry", "omc_gpu_test_rng", "omc_gpu_test_particles", "omc_gpu_test_samplers",
           "omc_gpu_abi_sizeof",
           "omc_gpu_shard_range", "omc_gpu_comm_unique_id", "omc_gpu_comm_init", "omc_gpu_comm_rank", "omc_gpu_comm_size", "omc_gpu_comm_sum", "omc_gpu_comm_gather_columns",
           "omc_gpu_multi_create", "omc_gpu_multi_destroy", "omc_gpu_multi_size", "omc_gpu_multi_device", "omc_gpu_multi_last_error",
           "omc_gpu_multi_set_media", "omc_gpu_multi_set_geometry", "omc_gpu_multi_set_source_dosxyz", "omc_gpu_multi_set_source_matrad",
           "omc_gpu_multi_set_vrt", "omc_gpu_multi_set_seed", "omc_gpu_multi_set_option", "omc_gpu_multi_reset_tallies",
           "omc_gpu_multi_run_batch", "omc_gpu_multi_synchronize", "omc_gpu_multi_get_tallies", "omc_gpu_multi_accumulate_results",
           "omc_gpu_multi_write_3ddose", "omc_gpu_multi_get_counters", "omc_gpu_multi_run_beamlets", "omc_gpu_multi_fetch_columns"]


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m ompmc_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.omc_gpu_create.argtypes = [C.POINTER(H), C.c_int]
    lib.omc_gpu_destroy.argtypes = [H]; lib.omc_gpu_destroy.restype = None
    lib.omc_gpu_last_error.argtypes = [H]; lib.omc_gpu_last_error.restype = C.c_char_p
    lib.omc_gpu_set_media.argtypes = [H, C.POINTER(MediaTables)]
    lib.omc_gpu_set_geometry.argtypes = [H, C.POINTER(Geometry)]
    lib.omc_gpu_set_source_dosxyz.argtypes = [H, C.POINTER(SourceDosxyz)]
    lib.omc_gpu_set_source_matrad.argtypes = [H, C.POINTER(SourceMatrad)]
    lib.omc_gpu_set_vrt.argtypes = [H, C.c_int]
    lib.omc_gpu_set_seed.argtypes = [H, C.c_int, C.c_int]
    lib.omc_gpu_set_option.argtypes = [H, C.c_char_p, C.c_longlong]
    lib.omc_gpu_run_histories.argtypes = [H, C.c_longlong, C.c_longlong, C.c_int]
    lib.omc_gpu_accum_batch.argtypes = [H]
    lib.omc_gpu_run_batch.argtypes = [H, C.c_longlong, C.c_longlong, C.c_int]
    lib.omc_gpu_start_batch.argtypes = [H, C.c_longlong, C.c_longlong, C.c_int]
    lib.omc_gpu_finish_batches.argtypes = [H]
    lib.omc_gpu_completed_batches.argtypes = [H]
    lib.omc_gpu_synchronize.argtypes = [H]
    lib.omc_gpu_get_tallies.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.omc_gpu_get_batch_grid.argtypes = [H, C.c_void_p]
    lib.omc_gpu_accumulate_results.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.omc_gpu_write_3ddose.argtypes = [H, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.omc_gpu_test_format.argtypes = [H, C.c_int, C.c_longlong, C.c_void_p, C.c_char_p]
    lib.omc_gpu_run_beamlets.argtypes = [H, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_longlong)]
    lib.omc_gpu_fetch_columns.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.omc_gpu_reset_tallies.argtypes = [H, C.c_int]
    lib.omc_gpu_device_ptrs.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    lib.omc_gpu_stream.argtypes = [H]; lib.omc_gpu_stream.restype = C.c_void_p
    lib.omc_gpu_get_counters.argtypes = [H, C.POINTER(Counters)]
    lib.omc_gpu_get_history_records.argtypes = [H, C.c_void_p, C.c_longlong]
    lib.omc_gpu_test_geometry.argtypes = [H, C.c_int] + [C.c_void_p] * 7
    lib.omc_gpu_test_rng.argtypes = [H, C.c_longlong, C.c_int, C.c_void_p]
    lib.omc_gpu_abi_sizeof.argtypes = [C.c_int]
    lib.omc_gpu_test_particles.argtypes = [H, C.c_int] + [C.c_void_p] * 5 + [C.c_longlong, C.c_void_p]
    lib.omc_gpu_test_samplers.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
    # multi-GPU
    lib.omc_gpu_shard_range.argtypes = [C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.omc_gpu_comm_unique_id.argtypes = [C.c_char_p]
    lib.omc_gpu_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    lib.omc_gpu_comm_rank.argtypes = [H]; lib.omc_gpu_comm_size.argtypes = [H]
    lib.omc_gpu_comm_sum.argtypes = [H, C.c_void_p, C.c_int]
    lib.omc_gpu_comm_gather_columns.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.omc_gpu_multi_create.argtypes = [C.POINTER(H), C.c_int, C.c_void_p]
    lib.omc_gpu_multi_destroy.argtypes = [H]; lib.omc_gpu_multi_destroy.restype = None
    lib.omc_gpu_multi_size.argtypes = [H]
    lib.omc_gpu_multi_device.argtypes = [H, C.c_int]; lib.omc_gpu_multi_device.restype = C.c_void_p
    lib.omc_gpu_multi_last_error.argtypes = [H]; lib.omc_gpu_multi_last_error.restype = C.c_char_p
    lib.omc_gpu_multi_set_media.argtypes = [H, C.POINTER(MediaTables)]
    lib.omc_gpu_multi_set_geometry.argtypes = [H, C.POINTER(Geometry)]
    lib.omc_gpu_multi_set_source_dosxyz.argtypes = [H, C.POINTER(SourceDosxyz)]
    lib.omc_gpu_multi_set_source_matrad.argtypes = [H, C.POINTER(SourceMatrad)]
    lib.omc_gpu_multi_set_vrt.argtypes = [H, C.c_int]
    lib.omc_gpu_multi_set_seed.argtypes = [H, C.c_int, C.c_int]
    lib.omc_gpu_multi_set_option.argtypes = [H, C.c_char_p, C.c_longlong]
    lib.omc_gpu_multi_reset_tallies.argtypes = [H, C.c_int]
    lib.omc_gpu_multi_run_batch.argtypes = [H, C.c_longlong, C.c_longlong, C.c_int]
    lib.omc_gpu_multi_synchronize.argtypes = [H]
    lib.omc_gpu_multi_get_tallies.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.omc_gpu_multi_accumulate_results.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.omc_gpu_multi_write_3ddose.argtypes = [H, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.omc_gpu_multi_get_counters.argtypes = [H, C.POINTER(Counters)]
    lib.omc_gpu_multi_run_beamlets.argtypes = [H, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                               C.c_void_p, C.POINTER(C.c_longlong)]
    lib.omc_gpu_multi_fetch_columns.argtypes = [H, C.c_void_p, C.c_void_p]
    return lib


def shard_range_c(first: int, n: int, rank: int, world: int) -> tuple[int, int]:
    """omc_gpu_shard_range(): the slice of a batch that omc_gpu_run_batch() gives rank `rank` (host arithmetic, no GPU needed)."""
    lib = load_library()
    lo, cnt = C.c_longlong(0), C.c_longlong(0)
    if lib.omc_gpu_shard_range(int(first), int(n), int(rank), int(world), C.byref(lo), C.byref(cnt)) != 0:
        raise ValueError("omc_gpu_shard_range: bad arguments")
    return lo.value, cnt.value


_NCCL_PRELOADED = False


def preload_nccl() -> None:
    """Map the NCCL that PyTorch bundles BEFORE the library resolves NCCL with dlopen("libnccl.so.2") (csrc/omc_nccl.h): a
    Python process must hold one copy of NCCL, and torch's libtorch_cuda.so needs the symbols of its own (newer) one -- were the
    system library mapped first under the same SONAME, a later `import torch` would fail.  Plain C user codes have no such
    constraint and get the system NCCL."""
    global _NCCL_PRELOADED
    if _NCCL_PRELOADED:
        return
    _NCCL_PRELOADED = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            path = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OmcGpuError(RuntimeError):
    pass


def problem_structs(prob: dict):
    """The C structs of include/ompmc_b200.h over the arrays of a problem dict: (media, geometry, source, keep-alive list)."""
    keep = []

    def pd(name):
        a = _f64(prob[name]); keep.append(a)
        return a.ctypes.data_as(PD)

    def pi(name):
        a = _i32(prob[name]); keep.append(a)
        return a.ctypes.data_as(PI)
    mt = MediaTables()
    mt.nmed = int(prob["nmed"][0])
    for name, typ in MediaTables._fields_[1:]:
        if typ is PD:
            setattr(mt, name, pd(name))
        elif typ is PI:
            setattr(mt, name, pi(name))
        else:
            setattr(mt, name, float(prob[name][0]))
    g = Geometry()
    g.isize, g.jsize, g.ksize = int(prob["isize"][0]), int(prob["jsize"][0]), int(prob["ksize"][0])
    g.xbounds, g.ybounds, g.zbounds = pd("xbounds"), pd("ybounds"), pd("zbounds")
    g.med, g.rhof, g.pcut, g.ecut = pi("region_med"), pd("region_rhof"), pd("region_pcut"), pd("region_ecut")
    if "mr_nbeamlets" in prob:                     # matRad beamlet source (omc_matrad.c)
        m = SourceMatrad()
        m.spectrum, m.charge = int(prob["src_spectrum"][0]), int(prob["src_charge"][0])
        m.energy, m.deltak = float(prob["src_energy"][0]), float(prob["src_deltak"][0])
        m.cdfinv1, m.cdfinv2 = pd("src_cdfinv1"), pd("src_cdfinv2")
        m.nbixels, m.nbeams = int(prob["mr_nbeamlets"][0]), len(prob["mr_xsource"])
        m.ibeam = pi("mr_ibeam")
        for k in ["xsource", "ysource", "zsource", "xcorner", "ycorner", "zcorner", "xside1", "yside1", "zside1",
                  "xside2", "yside2", "zside2"]:
            setattr(m, k, pd("mr_" + k))
        return mt, g, m, keep
    s = SourceDosxyz()
    s.spectrum, s.charge = int(prob["src_spectrum"][0]), int(prob["src_charge"][0])
    s.energy, s.deltak = float(prob["src_energy"][0]), float(prob["src_deltak"][0])
    s.cdfinv1, s.cdfinv2 = pd("src_cdfinv1"), pd("src_cdfinv2")
    for k in ["ssd", "xinl", "xinu", "yinl", "yinu", "xsize", "ysize"]:
        setattr(s, k, float(prob["src_" + k][0]))
    for k in ["ixinl", "ixinu", "iyinl", "iyinu"]:
        setattr(s, k, int(prob["src_" + k][0]))
    return mt, g, s, keep


class GpuTransport:
    """One GPU context == one ``omc_gpu_handle``."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.omc_gpu_create(C.byref(self.h), device)
        if rc != 0:
            raise OmcGpuError(f"omc_gpu_create(device={device}) failed rc={rc}: CUDA device required, no CPU fallback")
        self.device = device
        self._keep = []
        self.nreg = 0

    def close(self):
        if self.h:
            self.lib.omc_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise OmcGpuError(f"{what} failed (rc={rc}): {self.lib.omc_gpu_last_error(self.h).decode()}")

    # -- problem upload ----------------------------------------------------------------------
    _PFX = "omc_gpu_"

    def load_problem(self, prob: dict, seeds=(97, 33)):
        mt, g, src, keep = problem_structs(prob)
        f = lambda name: getattr(self.lib, self._PFX + name)         # noqa: E731
        self._ck(f("set_media")(self.h, C.byref(mt)), self._PFX + "set_media")
        self._ck(f("set_geometry")(self.h, C.byref(g)), self._PFX + "set_geometry")
        self.nreg = g.isize * g.jsize * g.ksize + 1
        kind = "set_source_matrad" if isinstance(src, SourceMatrad) else "set_source_dosxyz"
        self._ck(f(kind)(self.h, C.byref(src)), self._PFX + kind)
        self._ck(f("set_vrt")(self.h, int(prob["nsplit"][0])), self._PFX + "set_vrt")
        self._ck(f("set_seed")(self.h, int(seeds[0]), int(seeds[1])), self._PFX + "set_seed")
        del keep   # arrays were copied to the device by the set_* calls

    def set_option(self, key: str, value: int):
        self._ck(self.lib.omc_gpu_set_option(self.h, key.encode(), int(value)), f"omc_gpu_set_option({key})")

    def set_nsplit(self, n: int):
        self._ck(self.lib.omc_gpu_set_vrt(self.h, int(n)), "omc_gpu_set_vrt")

    # -- hot path ----------------------------------------------------------------------------
    def run_histories(self, first: int, n: int, records: bool = False, ibeamlet: int = -1):
        self.set_option("record_histories", 1 if records else 0)
        self._ck(self.lib.omc_gpu_run_histories(self.h, first, n, ibeamlet), "omc_gpu_run_histories")
        if records:
            rec = np.zeros(n, dtype=RECORD_DTYPE)
            self._ck(self.lib.omc_gpu_get_history_records(self.h, rec.ctypes.data, n), "omc_gpu_get_history_records")
            return rec
        return None

    def accum_batch(self):
        self._ck(self.lib.omc_gpu_accum_batch(self.h), "omc_gpu_accum_batch")

    accum_endep = accum_batch

    def run_batch(self, first: int, n: int, ibeamlet: int = -1):
        self._ck(self.lib.omc_gpu_run_batch(self.h, first, n, ibeamlet), "omc_gpu_run_batch")

    def run_beamlets(self, first: int, nhist: int, nbatch: int, ib0: int, nb: int, rel_threshold: float, med_densities: np.ndarray,
                     fetch: bool = True):
        """Beamlets [ib0, ib0+nb) in one pass + device-side column assembly: (jc[nb+1], ir[nnz], val[nnz]); fetch=False leaves
        rows and values on the device (jc only)."""
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        assert dens.size == self.nreg - 1
        jc = np.zeros(nb + 1, dtype=np.int64)
        tot = C.c_longlong(0)
        self._ck(self.lib.omc_gpu_run_beamlets(self.h, int(first), int(nhist), int(nbatch), int(ib0), int(nb), float(rel_threshold),
                                               dens.ctypes.data, jc.ctypes.data, C.byref(tot)), "omc_gpu_run_beamlets")
        if not fetch:
            return jc, None, None
        ir = np.zeros(max(tot.value, 1), dtype=np.int64); val = np.zeros(max(tot.value, 1))
        self._ck(self.lib.omc_gpu_fetch_columns(self.h, ir.ctypes.data, val.ctypes.data), "omc_gpu_fetch_columns")
        return jc, ir[:tot.value], val[:tot.value]

    def beamlet_capacity(self) -> int:
        """default beamlets per run_beamlets() pass: OMC_BEAMLETS_PER_PASS, fewer when their fp32 dose grids exceed
        OMC_BEAMLET_GRID_BUDGET (include/ompmc_b200.h, with the measurement behind the 64)"""
        return max(1, min(BEAMLETS_PER_PASS, int(BEAMLET_GRID_BUDGET // (self.nreg * 4))))

    def start_batch(self, first: int, n: int, ibeamlet: int = -1):
        """Pipelined batch, accumulation left to the caller (see include/ompmc_b200.h)."""
        self.set_option("record_histories", 0)
        self._ck(self.lib.omc_gpu_start_batch(self.h, first, n, ibeamlet), "omc_gpu_start_batch")

    def finish_batches(self):
        self._ck(self.lib.omc_gpu_finish_batches(self.h), "omc_gpu_finish_batches")

    def completed_batches(self) -> int:
        return int(self.lib.omc_gpu_completed_batches(self.h))

    def synchronize(self):
        self._ck(self.lib.omc_gpu_synchronize(self.h), "omc_gpu_synchronize")

    def get_endep(self) -> np.ndarray:
        out = np.zeros(self.nreg)
        self._ck(self.lib.omc_gpu_get_batch_grid(self.h, out.ctypes.data), "omc_gpu_get_batch_grid")
        return out

    def get_tallies(self):
        a = np.zeros(self.nreg); a2 = np.zeros(self.nreg); e = C.c_double(0.0)
        self._ck(self.lib.omc_gpu_get_tallies(self.h, a.ctypes.data, a2.ctypes.data, C.addressof(e)), "omc_gpu_get_tallies")
        return a, a2, e.value

    get_accum = get_tallies

    def accumulate_results(self, med_densities: np.ndarray, nhist: int, nbatch: int, iout: int = 1):
        """accumulateResults() on the device (omc_dosxyz.c:719-799): (dose[nvox], rel_sigma[nvox])."""
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        assert dens.size == self.nreg - 1
        dose = np.zeros(self.nreg); unc = np.zeros(self.nreg)
        self._ck(self.lib.omc_gpu_accumulate_results(self.h, int(iout), int(nhist), int(nbatch), dens.ctypes.data, dose.ctypes.data,
                                                     unc.ctypes.data), "omc_gpu_accumulate_results")
        return dose[1:], unc[1:]

    def write_3ddose(self, path: str, med_densities: np.ndarray, nhist: int, nbatch: int, iout: int = 1):
        """outputResults() (omc_dosxyz.c:801-886): accumulateResults() and the .3ddose text both produced on the device."""
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        assert dens.size == self.nreg - 1
        self._ck(self.lib.omc_gpu_write_3ddose(self.h, os.fsencode(path), int(iout), int(nhist), int(nbatch), dens.ctypes.data),
                 "omc_gpu_write_3ddose")

    def test_format(self, mode: int, values: np.ndarray, path: str):
        """unit-test hook: `values` through the device formatter ("%e " mode 0, "%f " mode 1) into `path`."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.lib.omc_gpu_test_format(self.h, int(mode), v.size, v.ctypes.data, os.fsencode(path)), "omc_gpu_test_format")

    def reset_tallies(self, which: int = 0):
        self._ck(self.lib.omc_gpu_reset_tallies(self.h, which), "omc_gpu_reset_tallies")

    reset_score = reset_tallies

    def counters(self) -> dict:
        c = Counters()
        self._ck(self.lib.omc_gpu_get_counters(self.h, C.byref(c)), "omc_gpu_get_counters")
        out = {n: int(getattr(c, n)) for n in COUNTER_NAMES}
        if any(c.reserved[:4]):
            out["reserved"] = [int(v) for v in c.reserved[:4]]
        return out

    def stream_ptr(self) -> int:
        return int(self.lib.omc_gpu_stream(self.h) or 0)

    def device_ptrs(self):
        e, a, a2, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        self._ck(self.lib.omc_gpu_device_ptrs(self.h, C.byref(e), C.byref(a), C.byref(a2), C.byref(n)), "omc_gpu_device_ptrs")
        return int(e.value), int(a.value), int(a2.value), int(n.value)

    # -- unit hooks --------------------------------------------------------------------------
    def test_geometry(self, xyzuvw, ir, ustep_in):
        xyzuvw = _f64(xyzuvw); ir = _i32(ir); ustep_in = _f64(ustep_in)
        n = len(ir)
        idisc = np.zeros(n, np.int32); irnew = np.zeros(n, np.int32); us = np.zeros(n); tp = np.zeros(n)
        self._ck(self.lib.omc_gpu_test_geometry(self.h, n, xyzuvw.ctypes.data, ir.ctypes.data, ustep_in.ctypes.data,
                                                idisc.ctypes.data, irnew.ctypes.data, us.ctypes.data, tp.ctypes.data),
                 "omc_gpu_test_geometry")
        return idisc, irnew, us, tp

    def test_particles(self, iq, e, xyzuvw, ir, wt=None, first_history: int = 0):
        iq = _i32(iq); e = _f64(e); xyzuvw = _f64(xyzuvw); ir = _i32(ir)
        n = len(iq)
        wt = _f64(np.ones(n) if wt is None else wt)
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        self._ck(self.lib.omc_gpu_test_particles(self.h, n, iq.ctypes.data, e.ctypes.data, xyzuvw.ctypes.data, ir.ctypes.data,
                                                 wt.ctypes.data, first_history, rec.ctypes.data), "omc_gpu_test_particles")
        return rec

    def test_samplers(self, which: int, inputs, first_history: int = 0) -> np.ndarray:
        """One sampler of the PRODUCTION kernels on explicit inputs (OMC_SAMPLER_* of include/ompmc_b200.h): n records of up to
        8 doubles in -> n records of 8 doubles out, record i drawing from the Philox stream of history first_history + i."""
        inputs = np.asarray(inputs, dtype=np.float64)
        a = np.zeros((len(inputs), 8))
        a[:, :inputs.shape[1]] = inputs
        out = np.zeros_like(a)
        self._ck(self.lib.omc_gpu_test_samplers(self.h, int(which), len(a), a.ctypes.data, int(first_history), out.ctypes.data),
                 "omc_gpu_test_samplers")
        return out

    def test_rng(self, hist: int, n: int) -> np.ndarray:
        out = np.zeros(n)
        self._ck(self.lib.omc_gpu_test_rng(self.h, hist, n, out.ctypes.data), "omc_gpu_test_rng")
        return out

    # -- multi-GPU: NCCL inside the library (one process per GPU) ------------------------------
    def comm_unique_id(self) -> bytes:
        preload_nccl()
        buf = C.create_string_buffer(128)
        rc = self.lib.omc_gpu_comm_unique_id(buf)
        if rc != 0:
            raise OmcGpuError(f"omc_gpu_comm_unique_id failed rc={rc}: NCCL not available")
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes):
        assert len(uid) == 128
        preload_nccl()
        self._ck(self.lib.omc_gpu_comm_init(self.h, int(rank), int(world), uid), "omc_gpu_comm_init")

    def comm_gather_columns(self, nbeamlets: int, mine: dict):
        """{beamlet: (rows, values)} of this rank -> (jc, ir, val) of the complete matrix on every rank (collective)."""
        cnt = np.zeros(nbeamlets)
        for b, (r, _) in mine.items():
            cnt[b] = len(r)
        cnt = self.comm_sum(cnt)
        jc = np.zeros(nbeamlets + 1, dtype=np.int64)
        jc[1:] = np.cumsum(cnt.astype(np.int64))
        flag = np.zeros(nbeamlets, dtype=np.uint8)
        flag[list(mine)] = 1
        order = sorted(mine)
        ir_m = np.ascontiguousarray(np.concatenate([mine[b][0] for b in order]) if order else np.zeros(0), dtype=np.int64)
        val_m = np.ascontiguousarray(np.concatenate([mine[b][1] for b in order]) if order else np.zeros(0), dtype=np.float64)
        tot = int(jc[-1])
        ir = np.zeros(max(tot, 1), dtype=np.int64); val = np.zeros(max(tot, 1))
        self._ck(self.lib.omc_gpu_comm_gather_columns(self.h, int(nbeamlets), jc.ctypes.data, flag.ctypes.data, ir_m.ctypes.data,
                                                      val_m.ctypes.data, ir.ctypes.data, val.ctypes.data), "omc_gpu_comm_gather_columns")
        return jc, ir[:tot], val[:tot]

    def comm_sum(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._ck(self.lib.omc_gpu_comm_sum(self.h, v.ctypes.data, v.size), "omc_gpu_comm_sum")
        return v


class MultiGpuTransport(GpuTransport):
    """``omc_gpu_multi``: several GPUs of this node behind one handle in ONE process (one host thread per device inside the
    library, NCCL between the devices); same calls as GpuTransport for the batch loop and the results."""
    _PFX = "omc_gpu_multi_"

    def __init__(self, ndev: int = 0, device_ids=None):
        self.lib = load_library()
        preload_nccl()
        self.h = C.c_void_p()
        ids = None if device_ids is None else _i32(device_ids)
        rc = self.lib.omc_gpu_multi_create(C.byref(self.h), int(ndev if ids is None else len(ids)), None if ids is None else ids.ctypes.data)
        if rc != 0:
            raise OmcGpuError(f"omc_gpu_multi_create failed rc={rc}: CUDA devices (and NCCL for more than one) required")
        self.ndev = int(self.lib.omc_gpu_multi_size(self.h))
        self.device = 0
        self.nreg = 0

    def close(self):
        if self.h:
            self.lib.omc_gpu_multi_destroy(self.h)
            self.h = C.c_void_p()

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise OmcGpuError(f"{what} failed (rc={rc}): {self.lib.omc_gpu_multi_last_error(self.h).decode()}")

    def set_option(self, key: str, value: int):
        self._ck(self.lib.omc_gpu_multi_set_option(self.h, key.encode(), int(value)), f"omc_gpu_multi_set_option({key})")

    def run_batch(self, first: int, n: int, ibeamlet: int = -1):
        self._ck(self.lib.omc_gpu_multi_run_batch(self.h, first, n, ibeamlet), "omc_gpu_multi_run_batch")

    def synchronize(self):
        self._ck(self.lib.omc_gpu_multi_synchronize(self.h), "omc_gpu_multi_synchronize")

    def reset_tallies(self, which: int = 0):
        self._ck(self.lib.omc_gpu_multi_reset_tallies(self.h, which), "omc_gpu_multi_reset_tallies")

    def get_tallies(self):
        a = np.zeros(self.nreg); a2 = np.zeros(self.nreg); e = C.c_double(0.0)
        self._ck(self.lib.omc_gpu_multi_get_tallies(self.h, a.ctypes.data, a2.ctypes.data, C.addressof(e)), "omc_gpu_multi_get_tallies")
        return a, a2, e.value

    def accumulate_results(self, med_densities: np.ndarray, nhist: int, nbatch: int, iout: int = 1):
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        dose = np.zeros(self.nreg); unc = np.zeros(self.nreg)
        self._ck(self.lib.omc_gpu_multi_accumulate_results(self.h, int(iout), int(nhist), int(nbatch), dens.ctypes.data, dose.ctypes.data,
                                                           unc.ctypes.data), "omc_gpu_multi_accumulate_results")
        return dose[1:], unc[1:]

    def write_3ddose(self, path: str, med_densities: np.ndarray, nhist: int, nbatch: int, iout: int = 1):
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        self._ck(self.lib.omc_gpu_multi_write_3ddose(self.h, os.fsencode(path), int(iout), int(nhist), int(nbatch), dens.ctypes.data),
                 "omc_gpu_multi_write_3ddose")

    def counters(self) -> dict:
        c = Counters()
        self._ck(self.lib.omc_gpu_multi_get_counters(self.h, C.byref(c)), "omc_gpu_multi_get_counters")
        return {n: int(getattr(c, n)) for n in COUNTER_NAMES}

    def run_beamlets(self, first: int, nhist: int, nbatch: int, ib0: int, nb: int, rel_threshold: float, med_densities: np.ndarray,
                     per_pass: int = 0):
        dens = np.ascontiguousarray(med_densities, dtype=np.float64)
        jc = np.zeros(nb + 1, dtype=np.int64)
        tot = C.c_longlong(0)
        self._ck(self.lib.omc_gpu_multi_run_beamlets(self.h, int(first), int(nhist), int(nbatch), int(ib0), int(nb), int(per_pass),
                                                     float(rel_threshold), dens.ctypes.data, jc.ctypes.data, C.byref(tot)),
                 "omc_gpu_multi_run_beamlets")
        ir = np.zeros(max(tot.value, 1), dtype=np.int64); val = np.zeros(max(tot.value, 1))
        self._ck(self.lib.omc_gpu_multi_fetch_columns(self.h, ir.ctypes.data, val.ctypes.data), "omc_gpu_multi_fetch_columns")
        return jc, ir[:tot.value], val[:tot.value]

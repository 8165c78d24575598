"""Host-side problem assembly for the ompmc_b200 hot path.

Mirrors the parts of the reference user code that sit *above* the batch loop and feed it:

* ``.egsphant`` writer/reader             -- ``initPhantom()``  ucodes/omc_dosxyz/omc_dosxyz.c:62-175
* per-region transport data                -- ``initRegions()``  omc_dosxyz.c:890-962
* spectrum -> inverse CDF, collimator      -- ``initSource()``   omc_dosxyz.c:368-632
* batch bookkeeping (``atoi`` ints)        -- ``main()``         omc_dosxyz.c:1207-1225  
* batch statistics + ``.3ddose`` writer    -- ``accumulateResults()/outputResults()`` :719-886

Physics tables (``initMediaData()``, src/ompmc.c:5450) are *not* rebuilt here: they arrive as a
"media blob" (oracle/omc_blob.h) dumped from the reference's own init chain
(oracle/gen_fixtures.py), exactly as the C user code would pass its global structs to the C-ABI.
"""
from __future__ import annotations

import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

MXGE, MXEKE, MXRAYFF = 2000, 500, 100
RM = 0.5109989461

# --------------------------------------------------------------------------------------------
# blob container (layout: oracle/omc_blob.h)
# --------------------------------------------------------------------------------------------

def load_blob(path: str) -> dict[str, np.ndarray]:
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"OMCBLOB1":
        raise ValueError(f"{path}: not an OMCBLOB1 file")
    (n,) = struct.unpack_from("<I", data, 8)
    off = 12
    out: dict[str, np.ndarray] = {}
    for _ in range(n):
        name = data[off:off + 32].split(b"\0", 1)[0].decode()
        dtype, _pad, count = struct.unpack_from("<IIQ", data, off + 32)
        off += 48
        np_dtype = np.float64 if dtype == 0 else np.int32
        nbytes = count * np.dtype(np_dtype).itemsize
        out[name] = np.frombuffer(data, dtype=np_dtype, count=count, offset=off).copy()
        off += (nbytes + 7) & ~7
    return out


def save_blob(path: str, arrays: dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"OMCBLOB1")
        f.write(struct.pack("<I", len(arrays)))
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            if arr.dtype == np.float64:
                code = 0
            elif arr.dtype == np.int32:
                code = 1
            else:
                raise TypeError(f"{name}: dtype {arr.dtype} not supported")
            raw = arr.tobytes()
            f.write(name.encode().ljust(32, b"\0")[:32])
            f.write(struct.pack("<IIQ", code, 0, arr.size))
            f.write(raw)
            if len(raw) % 8:
                f.write(b"\0" * (8 - len(raw) % 8))


MEDIA_KEYS_PREFIXES = ("src_", "region_", "med_", "xbounds", "ybounds", "zbounds", "isize", "jsize", "ksize", "nsplit")


def media_only(blob: dict[str, np.ndarray]) -> dict[str, np.ndarray]:
    """Strip geometry / source / vrt entries from a full problem blob."""
    return {k: v for k, v in blob.items() if not k.startswith(MEDIA_KEYS_PREFIXES)}


# --------------------------------------------------------------------------------------------
# phantoms (.egsphant)
# --------------------------------------------------------------------------------------------

@dataclass
class Phantom:
    """struct Geom, omc_dosxyz.c:46-59.  ``med_indices`` are 1-based as in the file."""
    media: list[str]
    xbounds: np.ndarray
    ybounds: np.ndarray
    zbounds: np.ndarray
    med_indices: np.ndarray      # int32 [ksize*jsize*isize], x fastest
    med_densities: np.ndarray    # float64, g/cm3

    @property
    def isize(self) -> int: return len(self.xbounds) - 1
    @property
    def jsize(self) -> int: return len(self.ybounds) - 1
    @property
    def ksize(self) -> int: return len(self.zbounds) - 1
    @property
    def nvox(self) -> int: return self.isize * self.jsize * self.ksize
    @property
    def nreg(self) -> int: return self.nvox + 1


def _bounds(lo: float, n: int, d: float) -> np.ndarray:
    # values as they would be after a %.6f-ish text round trip through the .egsphant
    return np.array([float(f"{lo + i * d:.6f}") for i in range(n + 1)], dtype=np.float64)


def water_phantom(medium: str, n=(61, 61, 60), voxel=(0.5, 0.5, 0.5), density=1.0) -> Phantom:
    """Homogeneous slab, centred in x/y, z from 0 (SURVEY 8d config 1/2 stand-in for WATER.egsphant)."""
    nx, ny, nz = n
    dx, dy, dz = voxel
    return Phantom([medium], _bounds(-0.5 * nx * dx, nx, dx), _bounds(-0.5 * ny * dy, ny, dy), _bounds(0.0, nz, dz),
                   np.ones(nx * ny * nz, dtype=np.int32), np.full(nx * ny * nz, density, dtype=np.float64))


TISSUE4 = ["AIR700ICRU", "LUNG700ICRU", "ICRUTISSUE700ICRU", "ICRPBONE700ICRU"]
TISSUE4_RHO = [0.0012, 0.26, 1.0, 1.85]


def tissue_phantom(n=(81, 81, 80), voxel=(0.3, 0.3, 0.3), kind="tg119") -> Phantom:
    """Heterogeneous 4-media stand-in for TG119 / PROSTATE .egsphant (both missing from the checkout,
    /root/reference/.MISSING_LARGE_BLOBS).  ``tg119``: tissue cylinder in air with a lung and a bone
    insert; ``prostate``: elliptical tissue body in air with two femoral-head-like bone cylinders
    and a low-density rectal pocket.  Density varies a few % voxel to voxel (deterministic) so that
    rhof != 1 paths are exercised."""
    nx, ny, nz = n
    dx, dy, dz = voxel
    xb, yb, zb = _bounds(-0.5 * nx * dx, nx, dx), _bounds(-0.5 * ny * dy, ny, dy), _bounds(0.0, nz, dz)
    xc = 0.5 * (xb[:-1] + xb[1:]); yc = 0.5 * (yb[:-1] + yb[1:]); zc = 0.5 * (zb[:-1] + zb[1:])
    Z, Y, X = np.meshgrid(zc, yc, xc, indexing="ij")
    zmid = 0.5 * (zb[0] + zb[-1]); zhalf = 0.5 * (zb[-1] - zb[0])
    med = np.ones((nz, ny, nx), dtype=np.int32)                 # air
    if kind == "tg119":
        R = 0.46 * min(nx * dx, 2 * zhalf)
        body = X ** 2 + (Z - zmid) ** 2 <= R ** 2
        med[body] = 3
        med[(X + 0.4 * R) ** 2 + (Z - zmid) ** 2 <= (0.22 * R) ** 2] = 2     # lung insert
        med[(X - 0.4 * R) ** 2 + (Z - zmid) ** 2 <= (0.18 * R) ** 2] = 4     # bone insert
    else:
        a, b = 0.47 * nx * dx, 0.45 * 2 * zhalf
        body = (X / a) ** 2 + ((Z - zmid) / b) ** 2 <= 1.0
        med[body] = 3
        for sx in (-1.0, 1.0):
            med[(X - sx * 0.55 * a) ** 2 + (Z - zmid) ** 2 <= (0.16 * a) ** 2] = 4
        med[(X / (0.12 * a)) ** 2 + ((Z - zmid - 0.35 * b) / (0.12 * b)) ** 2 <= 1.0] = 2
    rho = np.asarray(TISSUE4_RHO)[med - 1]
    # deterministic +-3 % texture on tissue / bone / lung (not air)
    tex = 1.0 + 0.03 * np.sin(1.7 * X + 0.9 * Y + 1.3 * Z)
    rho = np.where(med > 1, np.round(rho * tex, 4), rho)
    return Phantom(list(TISSUE4), xb, yb, zb, med.reshape(-1).astype(np.int32), rho.reshape(-1).astype(np.float64))


def write_egsphant(path: str, ph: Phantom) -> None:
    """Format parsed by initPhantom(), omc_dosxyz.c:84-152 (single-digit media indices, x fastest)."""
    with open(path, "w") as f:
        f.write(f"{len(ph.media)}\n")
        for m in ph.media:
            f.write(m + "\n")
        f.write("  ".join(["0.25"] * len(ph.media)) + "\n")
        f.write(f"{ph.isize} {ph.jsize} {ph.ksize}\n")
        for b in (ph.xbounds, ph.ybounds, ph.zbounds):
            f.write(" ".join(f"{v:.6f}" for v in b) + "\n")
        med = ph.med_indices.reshape(ph.ksize, ph.jsize, ph.isize)
        for k in range(ph.ksize):
            for j in range(ph.jsize):
                f.write("".join(str(int(v)) for v in med[k, j]) + "\n")
            f.write("\n")
        rho = ph.med_densities.reshape(ph.ksize, ph.jsize, ph.isize)
        for k in range(ph.ksize):
            for j in range(ph.jsize):
                f.write(" ".join(f"{v:.6g}" for v in rho[k, j]) + "\n")
            f.write("\n")


def read_egsphant(path: str) -> Phantom:
    with open(path) as f:
        nmed = int(f.readline())
        media = ["".join(f.readline().split()) for _ in range(nmed)]
        f.readline()
        nx, ny, nz = (int(v) for v in f.readline().split()[:3])
        toks: list[str] = []

        def take(n: int) -> np.ndarray:
            while len(toks) < n:
                toks.extend(f.readline().split())
            vals = np.array([float(t) for t in toks[:n]]); del toks[:n]
            return vals
        xb, yb, zb = take(nx + 1), take(ny + 1), take(nz + 1)
        med = np.empty((nz, ny, nx), dtype=np.int32)
        for k in range(nz):
            for j in range(ny):
                line = f.readline().rstrip("\n")
                med[k, j] = [ord(c) - 48 for c in line[:nx]]
            f.readline()
        rho = np.array(f.read().split(), dtype=np.float64)[: nx * ny * nz]
    return Phantom(media, xb, yb, zb, med.reshape(-1), rho)


# --------------------------------------------------------------------------------------------
# regions, source, vrt  -> the arrays the C-ABI takes
# --------------------------------------------------------------------------------------------

def init_regions(ph: Phantom, media: dict[str, np.ndarray], ecut: float, pcut: float) -> dict[str, np.ndarray]:
    """initRegions(), omc_dosxyz.c:890-962.  Region 0 = outside (vacuum).  Q10: if PEGS AE > ecut the
    reference leaves region.ecut uninitialised -- refused here instead of reproduced."""
    nreg = ph.nreg
    med = np.empty(nreg, dtype=np.int32); rhof = np.zeros(nreg); pc = np.zeros(nreg); ec = np.zeros(nreg)
    med[0] = -1
    imed = ph.med_indices.astype(np.int32) - 1
    med[1:] = imed
    ap, ae, rho = media["pegs_ap"], media["pegs_ae"], media["pegs_rho"]
    if np.any(imed < 0):
        raise ValueError("vacuum voxels: the reference writes region 0 instead (omc_dosxyz.c:928-932); unsupported")
    if np.any(ae[imed] > ecut):
        raise ValueError("global ecut below PEGS AE: reference leaves region.ecut uninitialised (SURVEY Q10)")
    dens = ph.med_densities
    rhof[1:] = np.where(dens == 0.0, 1.0, dens / rho[imed])
    pc[1:] = np.where(ap[imed] <= pcut, pcut, ap[imed])
    ec[1:] = ecut
    return {"region_med": med, "region_rhof": rhof, "region_pcut": pc, "region_ecut": ec}


def parse_spectrum(text: str):
    lines = text.splitlines()
    hdr = lines[1].split()
    nensrc, enmin, imode = int(hdr[0]), float(hdr[1]), int(hdr[2])
    ens, pdf = [], []
    for ln in lines[2:2 + nensrc]:
        a, b = ln.split()[:2]
        ens.append(float(a)); pdf.append(float(b))
    return nensrc, enmin, imode, np.array(ens), np.array(pdf)


def spectrum_cdfinv(text: str, invdim: int = 1000):
    """Inverse-CDF tables of initSource(), omc_dosxyz.c:396-506 (1000 bins; float32 grid size quirk:
    ``gridsz = 1.0f/deltak`` is evaluated in double because deltak is double)."""
    nensrc, enmin, imode, ensrcd, srcpdf = parse_spectrum(text)
    srcpdf = srcpdf.copy()
    if imode == 1:
        srcpdf[0] *= ensrcd[0] - enmin
        srcpdf[1:] *= ensrcd[1:] - ensrcd[:-1]
    elif imode != 0:
        raise ValueError("Invalid mode number in spectrum file.")
    cdf = np.empty(nensrc)
    acc = 0.0
    for i in range(nensrc):          # sequential sum, same order as the reference
        acc = srcpdf[i] if i == 0 else acc + srcpdf[i]
        cdf[i] = acc
    fnorm = 1.0 / cdf[-1]
    cdf = cdf * fnorm
    gridsz = 1.0 / float(invdim)
    c1 = np.empty(invdim); c2 = np.empty(invdim)
    for k in range(invdim):
        ak = float(k) * gridsz
        i = 0
        while i < nensrc and not (ak <= cdf[i]):
            i += 1
        c1[k] = ensrcd[i - 1] if i != 0 else enmin
        c2[k] = ensrcd[i] - c1[k]
    return c1, c2


def init_source(ph: Phantom, collimator, ssd: float, charge: int = 0, spectrum_text: str | None = None,
                mono_energy: float = 0.0, cdfinv=None) -> dict[str, np.ndarray]:
    """Geometric part of initSource(), omc_dosxyz.c:520-621 (+ spectrum tables)."""
    xinl, xinu, yinl, yinu = (float(v) for v in collimator)
    xb, yb = ph.xbounds, ph.ybounds

    def clamp(lo, hi, b, n):
        if lo < b[0]: lo = b[0]
        if hi <= lo: hi = lo
        if hi > b[n]: hi = b[n]
        if lo > b[n]: lo = b[n]
        il = 0
        while b[il] <= lo and b[il + 1] < lo: il += 1
        iu = il - 1   # the reference reads b[-1] (out of bounds) when il == 0; ixinu/iyinu are only printed
        while (iu < 0 or b[iu] <= hi) and b[iu + 1] < hi: iu += 1
        return lo, hi, il, iu
    xinl, xinu, ixinl, ixinu = clamp(xinl, xinu, xb, ph.isize)
    yinl, yinu, iyinl, iyinu = clamp(yinl, yinu, yb, ph.jsize)
    if cdfinv is None and spectrum_text is not None:
        cdfinv = spectrum_cdfinv(spectrum_text)
    spectrum = 1 if cdfinv is not None else 0
    f = lambda v: np.array([v], dtype=np.float64)
    i = lambda v: np.array([v], dtype=np.int32)
    return {
        "src_spectrum": i(spectrum), "src_charge": i(charge), "src_energy": f(0.0 if spectrum else mono_energy),
        "src_deltak": f(float(len(cdfinv[0])) if spectrum else 0.0),
        "src_cdfinv1": np.asarray(cdfinv[0], dtype=np.float64) if spectrum else np.zeros(1),
        "src_cdfinv2": np.asarray(cdfinv[1], dtype=np.float64) if spectrum else np.zeros(1),
        "src_ssd": f(ssd), "src_xinl": f(xinl), "src_xinu": f(xinu), "src_yinl": f(yinl), "src_yinu": f(yinu),
        "src_xsize": f(xinu - xinl), "src_ysize": f(yinu - yinl),
        "src_ixinl": i(ixinl), "src_ixinu": i(ixinu), "src_iyinl": i(iyinl), "src_iyinu": i(iyinu),
    }


def geometry_arrays(ph: Phantom) -> dict[str, np.ndarray]:
    i = lambda v: np.array([v], dtype=np.int32)
    return {"isize": i(ph.isize), "jsize": i(ph.jsize), "ksize": i(ph.ksize),
            "xbounds": ph.xbounds, "ybounds": ph.ybounds, "zbounds": ph.zbounds,
            "med_indices": ph.med_indices.astype(np.int32), "med_densities": ph.med_densities}


def build_problem(media: dict[str, np.ndarray], ph: Phantom, *, ecut: float, pcut: float, collimator, ssd: float,
                  charge: int = 0, spectrum_text: str | None = None, cdfinv=None, mono_energy: float = 0.0,
                  nsplit: int = 1) -> dict[str, np.ndarray]:
    """Full problem dict == what a reference user code holds in its globals just before the batch loop."""
    prob = dict(media_only(media))
    prob.update(geometry_arrays(ph))
    prob.update(init_regions(ph, media, ecut, pcut))
    prob.update(init_source(ph, collimator, ssd, charge, spectrum_text, mono_energy, cdfinv))
    prob["nsplit"] = np.array([nsplit], dtype=np.int32)
    return prob


# --------------------------------------------------------------------------------------------
# batch bookkeeping + statistics + output
# --------------------------------------------------------------------------------------------

def c_atoi(s: str) -> int:
    """atoi(): leading whitespace, optional sign, digits; stops at the first non-digit ("1e9" -> 1)."""
    s = s.lstrip()
    sign, i = 1, 0
    if s[:1] in "+-":
        sign = -1 if s[0] == "-" else 1
        i = 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    v = sign * int(s[i:j]) if j > i else 0
    v = (v + 2 ** 31) % 2 ** 32 - 2 ** 31
    return v


def batch_plan(ncase: str | int, nbatch: str | int):
    """omc_dosxyz.c:1207-1225: returns (nhist, nbatch, nperbatch) with C int semantics."""
    nhist = c_atoi(str(ncase)); nb = c_atoi(str(nbatch))
    if int(nhist / nb) == 0:
        nhist = nb
    nper = int(nhist / nb)
    return nper * nb, nb, nper


def accumulate_results(ph: Phantom, accum: np.ndarray, accum2: np.ndarray, nhist: int, nbatch: int, iout: int = 1):
    """accumulateResults(), omc_dosxyz.c:719-799: returns (dose[nvox], rel_sigma[nvox])."""
    e = accum[1:] / float(nbatch)
    e2 = accum2[1:] / float(nbatch)
    with np.errstate(divide="ignore", invalid="ignore"):
        unc = np.sqrt((e2 - e * e) / float(nbatch - 1)) / e
    unc = np.where(e != 0.0, unc, 0.9999999)
    if iout:
        vol = (np.diff(ph.zbounds)[:, None, None] * np.diff(ph.ybounds)[None, :, None] * np.diff(ph.xbounds)[None, None, :]).reshape(-1)
        mass = vol * ph.med_densities
        with np.errstate(divide="ignore", invalid="ignore"):
            dose = e * 1.602e-10 / (mass * float(nhist))
    else:
        dose = e / float(nhist)
    dose = np.where(e != 0.0, dose, 0.0)
    air = ph.med_densities < 0.044
    dose = np.where(air, 0.0, dose)
    unc = np.where(air, 0.9999999, unc)
    return dose, unc


def write_3ddose(path: str, ph: Phantom, dose: np.ndarray, unc: np.ndarray) -> None:
    """outputResults(), omc_dosxyz.c:841-879 -- byte-compatible formatting."""
    with open(path, "w") as f:
        f.write("%5d%5d%5d\n" % (ph.isize, ph.jsize, ph.ksize))
        for b in (ph.xbounds, ph.ybounds, ph.zbounds):
            f.write("".join("%f " % v for v in b) + "\n")
        f.write("".join("%e " % v for v in dose) + "\n")
        f.write("".join("%f " % v for v in unc) + "\n")


def read_3ddose(path: str):
    with open(path) as f:
        nx, ny, nz = (int(f.read(5)) for _ in range(3))
        f.readline()
        xb = np.array(f.readline().split(), dtype=float); yb = np.array(f.readline().split(), dtype=float)
        zb = np.array(f.readline().split(), dtype=float)
        dose = np.array(f.readline().split(), dtype=float); unc = np.array(f.readline().split(), dtype=float)
    return (nx, ny, nz), (xb, yb, zb), dose, unc


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden(name: str) -> str:
    return os.path.join(GOLDEN_DIR, name)


# --------------------------------------------------------------------------------------------
# matRad beamlet source (ucodes/omc_matrad/omc_matrad.c): struct Source :507-541, filled from the matRad
# `mcSrc` struct (:702-748).  A bixel is a parallelogram corner + r1*side1 + r2*side2 on the isocentre plane.
# --------------------------------------------------------------------------------------------

def matrad_beamlets(ph: Phantom, gantry_deg=(0.0,), nbix=(3, 3), bixel_cm: float = 0.5, sad_cm: float = 100.0,
                    isocentre=None) -> dict[str, np.ndarray]:
    """Synthetic stand-in for matRad's beamlet export: for every gantry angle (rotation about the y axis,
    0 deg = beam travelling along +z) an nbix[0] x nbix[1] grid of square bixels centred on the isocentre."""
    if isocentre is None:
        isocentre = (0.5 * (ph.xbounds[0] + ph.xbounds[-1]), 0.5 * (ph.ybounds[0] + ph.ybounds[-1]),
                     0.5 * (ph.zbounds[0] + ph.zbounds[-1]))
    c = np.asarray(isocentre, dtype=np.float64)
    src, ib, corner, s1, s2 = [], [], [], [], []
    for b, ang in enumerate(gantry_deg):
        t = math.radians(ang)
        d = np.array([math.sin(t), 0.0, math.cos(t)])            # beam direction
        e1 = np.array([math.cos(t), 0.0, -math.sin(t)])          # in-plane axes of the isocentre plane
        e2 = np.array([0.0, 1.0, 0.0])
        src.append(c - sad_cm * d)
        for j in range(nbix[1]):
            for i in range(nbix[0]):
                corner.append(c + (i - 0.5 * nbix[0]) * bixel_cm * e1 + (j - 0.5 * nbix[1]) * bixel_cm * e2)
                s1.append(bixel_cm * e1); s2.append(bixel_cm * e2); ib.append(b)
    src, corner, s1, s2 = (np.asarray(a, dtype=np.float64) for a in (src, corner, s1, s2))
    out = {"mr_nbeamlets": np.array([len(ib)], dtype=np.int32), "mr_ibeam": np.asarray(ib, dtype=np.int32)}
    for k, a in (("source", src), ("corner", corner), ("side1", s1), ("side2", s2)):
        for ax, name in enumerate("xyz"):
            out[f"mr_{name}{k}"] = np.ascontiguousarray(a[:, ax])
    return out


def build_problem_matrad(media, ph: Phantom, beamlets: dict, *, ecut: float, pcut: float, charge: int = 0, cdfinv=None,
                         mono_energy: float = 0.0, nsplit: int = 1) -> dict[str, np.ndarray]:
    """Problem dict for the matRad user code: same media/geometry/regions, beamlet source instead of the
    dosxyz collimated point source."""
    prob = dict(media_only(media))
    prob.update(geometry_arrays(ph))
    prob.update(init_regions(ph, media, ecut, pcut))
    spectrum = 1 if cdfinv is not None else 0
    f = lambda v: np.array([v], dtype=np.float64)
    i = lambda v: np.array([v], dtype=np.int32)
    prob.update({"src_spectrum": i(spectrum), "src_charge": i(charge), "src_energy": f(0.0 if spectrum else mono_energy),
                 "src_deltak": f(float(len(cdfinv[0])) if spectrum else 0.0),
                 "src_cdfinv1": np.asarray(cdfinv[0], dtype=np.float64) if spectrum else np.zeros(1),
                 "src_cdfinv2": np.asarray(cdfinv[1], dtype=np.float64) if spectrum else np.zeros(1)})
    prob.update(beamlets)
    prob["nsplit"] = i(nsplit)
    return prob


def resample_phantom(ph: Phantom, factor=(2, 2, 2)) -> Phantom:
    """Split every voxel into factor[0] x factor[1] x factor[2] equal sub-voxels (medium and density copied).
    BASELINE config 5 up-samples the PROSTATE grid from 3 mm to 2 mm / 1 mm; an integer split keeps the
    material map exactly (3 mm -> 1 mm is factor 3; 3 mm -> 1.5 mm is factor 2)."""
    fx, fy, fz = factor

    def split(b, f):
        out = [b[0]]
        for i in range(len(b) - 1):
            for k in range(1, f + 1):
                out.append(b[i] + (b[i + 1] - b[i]) * k / f)
        return np.asarray(out, dtype=np.float64)
    med = ph.med_indices.reshape(ph.ksize, ph.jsize, ph.isize)
    rho = ph.med_densities.reshape(ph.ksize, ph.jsize, ph.isize)
    rep = lambda a: np.repeat(np.repeat(np.repeat(a, fz, axis=0), fy, axis=1), fx, axis=2)
    return Phantom(list(ph.media), split(ph.xbounds, fx), split(ph.ybounds, fy), split(ph.zbounds, fz),
                   rep(med).reshape(-1).astype(np.int32), rep(rho).reshape(-1).astype(np.float64))


def resample_phantom_to(ph: Phantom, voxel) -> Phantom:
    """The phantom on voxels of (about) ``voxel`` = (dx, dy, dz) cm, any ratio: same extent, round(extent / size) equal cells per
    axis, volume-weighted mean density, medium = the one filling most of the new voxel (ties: lowest index, vacuum competes).
    Mirror of phantom_resample() in ompmc_b200/host/omc_host_input.h (BASELINE config 5: 3 mm -> 2 mm / 1 mm)."""
    def axis(b, size):
        ext = b[-1] - b[0]
        m = max(1, int(np.floor(ext / size + 0.5)))
        out = b[0] + ext * np.arange(m + 1, dtype=np.float64) / m
        out[-1] = b[-1]
        return out

    def overlap(ob, nb):
        """(n_new, n_old) matrix of overlap lengths"""
        lo = np.maximum(nb[:-1, None], ob[None, :-1])
        hi = np.minimum(nb[1:, None], ob[None, 1:])
        return np.maximum(hi - lo, 0.0)
    xb, yb, zb = axis(ph.xbounds, voxel[0]), axis(ph.ybounds, voxel[1]), axis(ph.zbounds, voxel[2])
    n_new, n_old = (len(xb) - 1, len(yb) - 1, len(zb) - 1), (ph.isize, ph.jsize, ph.ksize)
    uniform = all(np.allclose(np.diff(b), np.diff(b)[0], rtol=1e-9) for b in (ph.xbounds, ph.ybounds, ph.zbounds))
    if uniform and all(m % n == 0 for m, n in zip(n_new, n_old)):
        # integer split of a uniform grid: every new voxel lies inside one old voxel (same result as the general rule below,
        # without its dense overlap matrices -- the 1 mm grid of config 5 has 8.1e7 voxels)
        out = resample_phantom(ph, tuple(m // n for m, n in zip(n_new, n_old)))
        return Phantom(list(ph.media), xb, yb, zb, out.med_indices, out.med_densities)
    wx, wy, wz = overlap(ph.xbounds, xb), overlap(ph.ybounds, yb), overlap(ph.zbounds, zb)
    med = ph.med_indices.reshape(ph.ksize, ph.jsize, ph.isize)
    rho = ph.med_densities.reshape(ph.ksize, ph.jsize, ph.isize)

    def apply(a):           # sum over old cells of w * a, axis by axis (x fastest, as the C loops accumulate)
        a = np.einsum("kji,xi->kjx", a, wx)
        a = np.einsum("kjx,yj->kyx", a, wy)
        return np.einsum("kyx,zk->zyx", a, wz)
    vtot = apply(np.ones_like(rho))
    mass = apply(rho)
    nmed = len(ph.media)
    vol = np.stack([apply((med == m).astype(np.float64)) for m in range(nmed + 1)])
    best = np.argmax(vol, axis=0)            # first maximum = lowest index on ties
    dens = np.where(vtot > 0, mass / np.maximum(vtot, 1e-300), 0.0)
    return Phantom(list(ph.media), xb, yb, zb, best.reshape(-1).astype(np.int32), dens.reshape(-1).astype(np.float64))


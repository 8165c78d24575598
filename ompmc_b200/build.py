"""Build libompmc_b200.so (hand-written CUDA kernels for sm_100a + the C-ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container as well as on the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libompmc_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
# per-source extra flags; the lock-step parity kernel must not contract a*b+c (see omc_lockstep.cu)
SOURCES = {
    "omc_lockstep.cu": ["-fmad=false"],
    "omc_wavefront.cu": [],
    "omc_capi.cu": [],
    "omc_multi.cu": [],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(tag: str, flags: list[str]) -> str:
    """An A/B build of the library with extra nvcc flags -> ompmc_b200/libompmc_b200_<tag>.so (measurement scripts select it with
    OMPMC_B200_LIB); objects go to build/<tag>/."""
    objdir = os.path.join(HERE, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        r = subprocess.run([nvcc()] + ARCH + COMMON + extra + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed for {src} ({tag})")
    lib = os.path.join(HERE, f"libompmc_b200_{tag}.so")
    r = subprocess.run([nvcc()] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ompmc_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            tune = os.environ.get("OMC_NVCC_FLAGS", "").split()     # experiments only, e.g. -DOMC_WAVE_MINBLOCKS=2
            cmd = [nvcc()] + ARCH + COMMON + extra + tune + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}")
    if force or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libompmc_b200.so failed")
    build_host()
    return LIB


HOST_SRC = os.path.join(HERE, "host", "omc_dosxyz_b200.c")
HOST_EXE = os.path.join(HERE, "host", "omc_dosxyz_b200")
MATRAD_SRC = os.path.join(HERE, "host", "omc_matrad_b200.c")
MATRAD_EXE = os.path.join(HERE, "host", "omc_matrad_b200")
HOST_COMMON = os.path.join(HERE, "host", "omc_host_common.h")
TABLES_SRC = os.path.join(HERE, "host", "omc_tables.c")
TABLES_LIB = os.path.join(HERE, "host", "libomc_tables.so")
FORMAT_SRC = os.path.join(CSRC, "omc_format_host.cc")
FORMAT_LIB = os.path.join(HERE, "libomc_format_host.so")


def build_host(force: bool = False) -> str:
    """The plain-C host drivers (omc_dosxyz: batch loop + statistics + .3ddose writer; omc_matrad: beamlet loop + CSC file)
    linked against the C-ABI library."""
    inc = os.path.join(os.path.dirname(HERE), "include")
    hdir = os.path.join(HERE, "host")
    deps = [HOST_COMMON, TABLES_SRC, os.path.join(hdir, "omc_tables.h"), os.path.join(hdir, "omc_host_input.h"), LIB, os.path.join(inc, "ompmc_b200.h")]
    # host-side physics table initialisation (PEGS4 / XCOM / form factors / msnew / spinms -> omc_media_tables), also as a
    # shared library of its own for the CPU tests; no contraction of a*b+c: the tables must equal the reference's bit for bit
    if force or _stale(TABLES_LIB, [TABLES_SRC, os.path.join(hdir, "omc_tables.h"), os.path.join(inc, "ompmc_b200.h")]):
        cmd = ["gcc", "-O2", "-Wall", "-ffp-contract=off", "-shared", "-fPIC", "-o", TABLES_LIB, TABLES_SRC, "-I", inc, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed for host/omc_tables.c")
    # omc_format.cuh (the device's "%e " / "%f " text conversion) compiled for the host: CPU test hook
    if force or _stale(FORMAT_LIB, [FORMAT_SRC, os.path.join(CSRC, "omc_format.cuh")]):
        cmd = ["g++", "-O2", "-Wall", "-std=c++17", "-shared", "-fPIC", "-o", FORMAT_LIB, FORMAT_SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed for csrc/omc_format_host.cc")
    for src, exe in ((HOST_SRC, HOST_EXE), (MATRAD_SRC, MATRAD_EXE)):
        if force or _stale(exe, [src] + deps):
            extra = [TABLES_SRC, "-ffp-contract=off"] if src == HOST_SRC else []
            cmd = ["gcc", "-O2", "-Wall", "-o", exe, src] + extra + ["-I", inc, "-I", hdir, "-L", HERE, "-lompmc_b200", "-lm", "-Wl,-rpath,$ORIGIN/.."]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"gcc failed for {os.path.basename(src)}")
    return HOST_EXE


if __name__ == "__main__":
    if "--variant" in sys.argv:         # build.py --variant TAG -DFOO=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

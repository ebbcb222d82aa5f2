"""Build libmvsnerf_b200.so (C ABI, include/mvsnerf_b200.h) in-tree with nvcc for sm_100a.

    python -m mvsnerf_b200.build [--force] [-v] [--trace] [--probes]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so it travels with a
repository snapshot (it is git-ignored, not gpurun-ignored).

    --trace   libmvsnerf_b200_trace.so: the product sources with the pipeline-timeline hooks compiled in
              (tools/tc_trace.py loads it through MVSN_LIB); the product library never carries them.
    --probes  libmvsnerf_b200_probes.so: csrc/probes/*.cu, the hardware bring-up probes used by tools/ --
              a separate library, so the product exports exactly what include/mvsnerf_b200.h declares.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
PROBES = os.path.join(CSRC, "probes")
LIB_PATH = os.path.join(HERE, "libmvsnerf_b200.so")
PROBES_LIB_PATH = os.path.join(HERE, "libmvsnerf_b200_probes.so")
BUILD_DIR = os.path.join(HERE, "csrc", "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _cu(d):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))


def _sources(probes: bool = False):
    if probes:
        return _cu(PROBES) + [os.path.join(CSRC, "tc_selftest.cu")]
    return _cu(CSRC)


def _digest(flags, probes: bool) -> str:
    """Content hash of everything the library is built from.  Paths enter RELATIVE to the repository
    root, so a prebuilt .so + stamp stay valid when the checkout moves (gpurun copies it elsewhere)."""
    h = hashlib.sha256(" ".join(flags).encode())
    deps = _sources(probes) + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    deps.append(os.path.join(ROOT, "include", "mvsnerf_b200.h"))
    for p in deps:
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, ROOT).encode() + b"\0" + f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, trace: bool = False, probes: bool = False) -> str:
    suffix = "_trace" if trace else "_probes" if probes else ""
    build_dir = BUILD_DIR + suffix
    lib_path = LIB_PATH.replace(".so", suffix + ".so")
    flags = NVCC_FLAGS + (["-DMVSN_TC_TRACE"] if trace else []) + (["-DMVSN_BUILD_PROBES"] if probes else [])
    os.makedirs(build_dir, exist_ok=True)
    stamp = os.path.join(build_dir, "stamp")
    digest = _digest(flags, probes) + suffix
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == digest:
        return lib_path
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources(probes)))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv,
                         probes="--probes" in sys.argv)
    print(path)

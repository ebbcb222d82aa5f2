"""Build libmvsnerf_b200.so (C ABI, include/mvsnerf_b200.h) in-tree with nvcc for sm_100a.

    python -m mvsnerf_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so it travels with a
repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmvsnerf_b200.so")
BUILD_DIR = os.path.join(HERE, "csrc", "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    wip_dir = os.path.join(CSRC, "wip")
    if os.path.isdir(wip_dir):
        deps += sorted(os.path.join(wip_dir, f) for f in os.listdir(wip_dir) if f.endswith(".cu"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mvsnerf_b200.h"))
    for p in deps:
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, trace: bool = False, wip: bool = False) -> str:
    """trace=True builds libmvsnerf_b200_trace.so with the pipeline-timeline hooks compiled in
    (tools/tc_trace.py loads it through MVSN_LIB); the product library never carries them.
    wip=True builds libmvsnerf_b200_wip.so, which additionally contains csrc/wip/*.cu (round-2 work in progress,
    reachable as mlp_mode 3 through MVSN_LIB); the product library never contains it either."""
    suffix = "_trace" if trace else "_wip" if wip else ""
    build_dir = BUILD_DIR + suffix
    lib_path = LIB_PATH.replace(".so", suffix + ".so")
    flags = NVCC_FLAGS + (["-DMVSN_TC_TRACE"] if trace else []) + (["-DMVSN_WIP_PAIR"] if wip else [])
    os.makedirs(build_dir, exist_ok=True)
    stamp = os.path.join(build_dir, "stamp")
    digest = _digest() + suffix
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

    sources = _sources()
    if wip:
        wip_dir = os.path.join(CSRC, "wip")
        sources += sorted(os.path.join(wip_dir, f) for f in os.listdir(wip_dir) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv,
                         wip="--wip" in sys.argv)
    print(path)

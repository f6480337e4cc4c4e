"""Builds libd3dsweep.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree).

    python -m deep3d_aerial_b200.build [--force] [--verbose]

Each translation unit is compiled in its own nvcc process (they run in parallel) into
`csrc/_build/*.o`, then linked into `deep3d_aerial_b200/libd3dsweep.so`.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libd3dsweep.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC,-O2,-Wall",
         "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "d3d_sweep.h"),
                                                                 os.path.abspath(__file__)]
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or not _newer(obj, [src] + headers):
            cmd = [nvcc] + ARCH + [f for f in FLAGS if f != "--use_fast_math=false"] + ["-c", src, "-o", obj]
            if ptxas_info:
                cmd += ["-Xptxas", "-v"]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, rc, log in ex.map(run, jobs):
            if verbose or rc or ptxas_info:
                sys.stderr.write("[nvcc] %s\n%s" % (os.path.basename(src), log))
            if rc:
                raise RuntimeError("nvcc failed on " + src)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]
    if force or jobs or not _newer(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))

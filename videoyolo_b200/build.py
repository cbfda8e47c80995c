"""In-tree build of the CUDA library: videoyolo_b200/csrc/*.cu -> videoyolo_b200/libvyolo.so.

nvcc cross-compiles for sm_100a without a GPU, so this runs in the build container; the .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "libvyolo.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # the arch-specific target: tcgen05/TMEM need the 'a' feature set
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                                   # box_nms / decode arithmetic must not be FMA-contracted
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return n


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), variant: str = "") -> str:
    """variant / extra_flags: development builds (tools/): libvyolo_<variant>.so with extra nvcc flags, own objects."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "vyolo.h")]
    OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + variant if variant else ""))
    SO = os.path.join(HERE, "libvyolo%s.so" % ("_" + variant if variant else ""))
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append(([nvcc] + NVCC_FLAGS + list(extra_flags) + extra + ["-c", s, "-o", o], o))
    def run(job):
        cmd, o = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (o, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return o
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(SO, objs):
        cmd = [nvcc, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                    "--cudart", "static",
                                                    "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map")]
        # static cudart: no dependence on which libcudart the host process (torch) brings; driver
        # entry points (cuTensorMapEncode*) are resolved at run time through cudaGetDriverEntryPoint,
        # so the library also loads on a box without libcuda.so.1 (the CPU-only build container).
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

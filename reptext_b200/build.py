"""Build csrc/*.cu into csrc/librt_reptext.so with nvcc for sm_100a (in-tree; the .so travels with gpurun).

    python -m reptext_b200.build [--force]         the product library
    python -m reptext_b200.build --ab [--force]    csrc/librt_reptext_ab.so: the same library compiled with
                                                   -DRT_AB_VARIANTS (the attention A/B kernels, the hand-off trace);
                                                   load it with RT_LIB=<path> (tools/attn_sweep.py, tools/attn_trace.py)
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "librt_reptext.so")
LIB_AB = os.path.join(CSRC, "librt_reptext_ab.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, bdir: str = "build", defines=()) -> str:
    obj = os.path.join(CSRC, bdir, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(CSRC, bdir, src[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True, ab: bool = False) -> str:
    bdir = "build_ab" if ab else "build"
    lib = LIB_AB if ab else LIB
    defines = ("-DRT_AB_VARIANTS",) if ab else ()
    os.makedirs(os.path.join(CSRC, bdir), exist_ok=True)
    stamp = os.path.join(CSRC, bdir, "digest.txt")
    dig = _digest(" ".join(defines))
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda src: _compile(src, bdir, defines), _sources()))
    cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", lib, *objs, "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(f"built {lib}")
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, ab="--ab" in sys.argv)

"""In-tree nvcc build of ``libb200rt.so`` for sm_100a (no JIT cache: the .so travels with the repo).

    python -m b200rt.build            # or b200rt.build.build()

Objects:
  rt_f32.o  float32 production kernels            (default flags, fused multiply-add on)
  rt_f64.o  float64 parity kernels, same source   (-fmad=false: the reference never contracts a*b+c)
  lbvh.o    LBVH builder (Morton + CUB radix sort + Karras hierarchy + refit)
  c_api.o   extern "C" surface of include/b200rt.h
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "b200rt", "libb200rt.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
UNITS = [
    ("rt_f32.cu", []),
    ("rt_f64.cu", ["-fmad=false"]),
    ("lbvh.cu", []),
    ("c_api.cu", []),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found: the B200 core cannot be built (no CPU fallback exists)")
    return exe


def _sources_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(PKG_DIR), "include", "b200rt.h"))
    return max(os.path.getmtime(p) for p in paths)


def is_stale() -> bool:
    return not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < _sources_mtime()


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """``defines``/``out`` build an experimental variant (e.g. ("-DB2RT_SCAN_UNROLL=2",)) next to the default."""
    lib_path = out or LIB_PATH
    if not force and not defines and not is_stale():
        return LIB_PATH
    tag = "" if not defines else "_" + "_".join(d.replace("-D", "").replace("=", "") for d in defines)
    os.makedirs(BUILD, exist_ok=True)
    exe = nvcc()
    objs, procs = [], []
    for src, extra in UNITS:
        obj = os.path.join(BUILD, src.replace(".cu", tag + ".o"))
        cmd = [exe, *ARCH, *COMMON, *extra, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            print(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
    link = [exe, *ARCH, "-shared", "-o", lib_path, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

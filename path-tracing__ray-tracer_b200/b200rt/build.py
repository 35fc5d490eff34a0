"""In-tree nvcc build of ``libb200rt.so`` for sm_100a (no JIT cache: the .so travels with the repo).

    python -m b200rt.build            # or b200rt.build.build()

Objects:
  rt_f32.o  float32 production kernels            (default flags, fused multiply-add on)
  rt_f64.o  float64 parity kernels, same source   (-fmad=false: the reference never contracts a*b+c)
  lbvh.o    LBVH builder (Morton + CUB radix sort + Karras hierarchy + refit)
  c_api.o   extern "C" surface of include/b200rt.h
  scene_prepare.o  small-scene record derivation (host C++; b2rt_scene_prepare)
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
    ("scene_prepare.cu", []),
]


def variant_path(name: str) -> str:
    """In-tree path of a named build variant (``libb200rt_<name>.so`` next to the default library)."""
    return os.path.join(PKG_DIR, "b200rt", f"libb200rt_{name}.so")


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
    """``defines``/``out`` build an experimental variant (e.g. ("-DB2RT_SCAN_UNROLL=2",)) next to the default.

    Safe under ``torchrun``: an exclusive ``fcntl`` lock serialises the ranks, the ones that waited re-check
    staleness under the lock (so only the first compiles), objects and the library are written to temporary
    names and moved into place with ``os.replace`` so no process can ``dlopen`` a half-written file."""
    import fcntl
    lib_path = out or LIB_PATH
    if not force and not defines and not is_stale():
        return LIB_PATH
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not defines and not is_stale():      # another rank built it while this one waited
                return LIB_PATH
            return _build_locked(verbose, tuple(defines), lib_path)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, defines: tuple, lib_path: str) -> str:
    tag = "" if not defines else "_" + "_".join(d.replace("-D", "").replace("=", "") for d in defines)
    exe = nvcc()
    pid = os.getpid()
    objs, procs = [], []
    for src, extra in UNITS:
        obj = os.path.join(BUILD, src.replace(".cu", tag + ".o"))
        tmp = f"{obj}.{pid}.tmp"
        cmd = [exe, *ARCH, *COMMON, *extra, *defines, "-c", os.path.join(CSRC, src), "-o", tmp]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, tmp, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = None
    for src, tmp, obj, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            print(log)
        if p.returncode != 0:
            failed = failed or f"nvcc failed on {src}:\n{log}"
        else:
            os.replace(tmp, obj)
    if failed:
        for _, tmp, _, _ in procs:
            if os.path.exists(tmp):
                os.remove(tmp)
        raise RuntimeError(failed)
    tmp_lib = f"{lib_path}.{pid}.tmp"
    link = [exe, *ARCH, "-shared", "-o", tmp_lib, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp_lib):
            os.remove(tmp_lib)
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_lib, lib_path)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds the C-ABI shared library (hand-written sm_100a CUDA, no torch headers) in-tree:
``desenet_b200/lib/libdesenet_b200.so``.  nvcc cross-compiles without a GPU; the built .so travels
to the GPU box with the repo snapshot (git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdesenet_b200.so")
SOURCES = ["api.cu", "seg.cu", "detect.cu", "nms.cu", "nms_small.cu", "fused.cu", "coords.cu", "match.cu", "stats.cu", "peer.cu", "headconv.cu", "detconv.cu"]
HEADERS = ["common.cuh", "geom.cuh", "filter_core.cuh", "nms_core.cuh", "nms_launch.cuh", os.path.join("..", "..", "include", "desenet_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static",
]
OBJ_TAG = "obj"
if os.environ.get("DSN_TIMING"):          # tuning build: phase timeline markers, separate .so
    NVCC_FLAGS += ["-DDSN_TIMING", "-rdc=true"]
    LIB = os.path.join(LIBDIR, "libdesenet_b200_timing.so")
    OBJ_TAG = "obj_timing"
VARIANT_FLAGS = []
VARIANT_SOURCES = None                    # sources recompiled with VARIANT_FLAGS (None: all); the others are linked from build/obj
if not os.environ.get("DSN_TIMING") and os.environ.get("DSN_VARIANT"):
    # tuning build "name:-DFOO=1 -DBAR=2[:a.cu,b.cu]": separate .so with extra compile-time switches
    _name, _, _rest = os.environ["DSN_VARIANT"].partition(":")
    _flags, _, _srcs = _rest.partition(":")
    VARIANT_FLAGS = _flags.split()
    VARIANT_SOURCES = _srcs.split(",") if _srcs else None
    LIB = os.path.join(LIBDIR, f"libdesenet_b200_{_name}.so")
    OBJ_TAG = f"obj_{_name}"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the desenet_b200 CUDA library cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "..", "build", OBJ_TAG)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for s in SOURCES:
        if VARIANT_SOURCES is not None and s not in VARIANT_SOURCES:     # reuse the default build's object
            objs.append(os.path.join(HERE, "..", "build", "obj", s.replace(".cu", ".o")))
            continue
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, *VARIANT_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
    tmp = LIB + ".tmp"
    out = subprocess.run([nvcc, "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                          "-o", tmp, *objs], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError(f"link failed:\n{out.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

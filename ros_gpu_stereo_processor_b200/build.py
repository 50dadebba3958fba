"""Builds libb200stereo.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m ros_gpu_stereo_processor_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiments: B200S_NVCC_DEFS="-DB200S_RECT_FTY=8 ..." with B200S_LIB_SUFFIX=_x builds libb200stereo_x.so beside the product
SUFFIX = os.environ.get("B200S_LIB_SUFFIX", "")
LIB = os.path.join(HERE, "libb200stereo%s.so" % SUFFIX)
SOURCES = ["api.cu", "slots.cu", "rectify.cu", "prefilter.cu", "bm_sad.cu", "bm_ws.cu", "bm_vh.cu", "bm_strip.cu", "bm_cuda_compat.cu", "post.cu", "reproject.cu", "intpeak.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-cudart", "static"] + os.environ.get("B200S_NVCC_DEFS", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    objdir = os.path.join(HERE, "build" + SUFFIX)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "handle.h"), os.path.join(CSRC, "bm_common.cuh"), os.path.join(HERE, "..", "include", "b200_stereo.h")]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

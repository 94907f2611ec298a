"""Build libhighway_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhighway_b200.so")
SOURCES = ["hw_sa.cu", "hw_ma.cu", "hw_ma_grp.cu", "hw_ma_tpe.cu", "hw_policy.cu", "hw_replay.cu"]
HEADERS = ["hw_common.cuh", "hw_ma_common.cuh", os.path.join("..", "..", "include", "highway_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",           # reference arithmetic is unfused fp64; never contract a*b+c
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libhighway_b200.so")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    cmd, obj = args
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False, out=None):
    """compile every CUDA source (one nvcc process per file, in parallel) and link them into one shared library; returns its path.
    Objects are kept under build/obj/<flags> and reused while the source, the headers and the flags are unchanged."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor
    lib = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("HW_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DHW_SA_MIN_BLOCKS_D=8
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    objroot = os.path.join(os.path.dirname(HERE), "build", "obj")
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    jobs, objs = [], []
    for s in srcs:
        # per-file experiment flags: HW_NVCC_EXTRA_HW_MA_TPE="-DHW_TPE_ONLY=5 ..." touches that object only
        fflags = cflags + os.environ.get("HW_NVCC_EXTRA_" + s[:-3].upper(), "").split()
        objdir = os.path.join(objroot, hashlib.sha1(" ".join(fflags).encode()).hexdigest()[:12])
        os.makedirs(objdir, exist_ok=True)
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)
        if verbose or not fresh or (force and not os.environ.get("HW_BUILD_REUSE_OBJECTS")):
            jobs.append(([_nvcc()] + fflags + ["-c", "-o", obj, src], obj))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(_compile_one, jobs))
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib + ".tmp"] + objs)
    os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

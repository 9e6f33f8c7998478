"""Build libcosmogp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cosmogp_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcosmogp_b200.so")
SOURCES = ["cgp_api.cu", "cgp_small.cu", "cgp_large.cu", "cgp_fit.cu", "cgp_stream.cu", "cgp_ctx.cu"]
HEADERS = [os.path.join(CSRC, "cgp_internal.h"), os.path.join(CSRC, "cgp_math.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "cosmogp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + ["cgp_small64.cu"]] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    extra = os.environ.get("CGP_BUILD_DEFS", "").split()      # experiment knobs, e.g. CGP_BUILD_DEFS="-DCGP64_GRID_U=1"
    units = [(src, list(extra), src.replace(".cu", ".o")) for src in SOURCES]
    # the static N <= 64 kernel: one translation unit per (dim, task), built in parallel
    units += [("cgp_small64.cu", extra + ["-DCGP64_DIM=%d" % d, "-DCGP64_TASK=%d" % t], "cgp_small64_d%d_t%d.o" % (d, t))
              for d in (1, 2) for t in (0, 1, 2, 4, 5)] + [("cgp_small64.cu", extra + ["-DCGP64_DIM=1", "-DCGP64_TASK=%d" % t], "cgp_small64_d1_t%d.o" % t) for t in (6, 7)]
    for src, defs, objname in units:
        obj = os.path.join(CSRC, objname)
        cmd = ([_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else [])
               + ["-c", os.path.join(CSRC, src), "-o", obj])
        procs.append((objname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

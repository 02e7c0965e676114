"""Build libattndm_b200.so (sm_100a only) in-tree with nvcc.

    python -m attentiondm_b200.build

The library links cudart statically and resolves the one driver symbol it needs
(cuTensorMapEncodeTiled) at run time, so it loads on a CPU-only box too (for the
symbol-export test); every compute entry point needs a B200.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libattndm_b200.so")
SOURCES = ["api.cu", "quant_kernels.cu", "conv_simt.cu", "conv_tc.cu", "misc_kernels.cu", "rowprog.cu"]
HEADERS = ["common.cuh", "conv_common.cuh", os.path.join("..", "..", "include", "attndm_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-cudart", "static"]
# debug builds only, e.g. ATTNDM_NVCC_EXTRA=-DATTNDM_TC_TRACE for tools/conv_trace.py (part of the stamp digest)
NVCC_FLAGS += os.environ.get("ATTNDM_NVCC_EXTRA", "").split()


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


# A/B variants of the library (same sources, extra defines), built next to the default one as
# libattndm_b200_<name>.so and selected at run time with ATTNDM_LIB=<path> (see _ffi.py):
VARIANTS = {"silu_guard": ["-DATTNDM_SILU_GUARD"], "silu_accurate": ["-DATTNDM_SILU_ACCURATE"],
            "tc_trace": ["-DATTNDM_TC_TRACE"], "rp_trace": ["-DATTNDM_RP_TRACE"],
            "epi8": ["-DATTNDM_TC_EPI_WARPS=8"], "epi16": ["-DATTNDM_TC_EPI_WARPS=16"]}


def build(force=False, verbose=False, variant=None):
    global NVCC_FLAGS
    if variant is not None:
        saved = NVCC_FLAGS
        NVCC_FLAGS = NVCC_FLAGS + VARIANTS[variant]
        try:
            return _build(force, verbose, LIB.replace(".so", f"_{variant}.so"), "build_" + variant)
        finally:
            NVCC_FLAGS = saved
    return _build(force, verbose, LIB, "build")


def _build(force, verbose, LIB, objdir_name):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    stamp = LIB + ".stamp"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objdir = os.path.join(HERE, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var[0] if var else None))

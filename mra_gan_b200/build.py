"""Build libmra_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libmra_b200.so")
SOURCES = ["mra_abi.cu"]
HEADERS = ["common.cuh", "conv_plan.h", "conv_naive.cuh", "conv_tc.cuh", "conv_tc_halo.cuh", "conv_tc_phase.cuh", "conv_tc_col.cuh", "conv_special.cuh", "norm.cuh", "norm_stream.cuh", "misc.cuh",
           os.path.join("..", "..", "include", "mra_gan_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [_nvcc(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

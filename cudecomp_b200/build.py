"""Builds cudecomp_b200/lib/libcudecomp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcudecomp.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cc")) + glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "mpi_shim", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every source of csrc/ into one shared library. Returns its path."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [
        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "mpi_shim"), "-I" + CSRC,
        "-shared", "-o", LIB_PATH] + sources() + ["-lrt", "-lpthread"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libcudecomp.so")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))

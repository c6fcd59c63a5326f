"""Builds the library in-tree with nvcc for sm_100a (cross-compiles without a GPU):
  cudecomp_b200/lib/libcudecomp.so          the engine + the MPI-subset shim for machines without MPI
  cudecomp_b200/lib/libcudecomp_realmpi.so  the same objects linked with csrc/realmpi.map: only cudecomp* symbols are
                                            exported, for applications that bring a real MPI (include/cudecomp_b200_mpi.h)
Sources are compiled to objects in parallel (cudecomp_b200/build/, git-ignored) and linked twice."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(LIB_DIR, "libcudecomp.so")
REALMPI_LIB_PATH = os.path.join(LIB_DIR, "libcudecomp_realmpi.so")
VERSION_SCRIPT = os.path.join(CSRC, "realmpi.map")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cc")) + glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "mpi_shim", "*.h"))


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(REALMPI_LIB_PATH):
        return True
    t = min(os.path.getmtime(LIB_PATH), os.path.getmtime(REALMPI_LIB_PATH))
    deps = sources() + _headers() + [VERSION_SCRIPT]
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libcudecomp.so")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)


def build_library(force=False, verbose=False):
    """Compile every source of csrc/ and link both shared libraries. Returns the path of libcudecomp.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "mpi_shim"), "-I" + CSRC]
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs, objects = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        objects.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append([nvcc] + NVCC_FLAGS + inc + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(lambda c: _run(c, verbose), jobs))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
    _run(link + ["-o", LIB_PATH] + objects + ["-lrt", "-lpthread", "-ldl"], verbose)
    _run(link + ["-Xlinker", "--version-script=" + VERSION_SCRIPT, "-o", REALMPI_LIB_PATH] + objects + ["-lrt", "-lpthread", "-ldl"],
         verbose)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))

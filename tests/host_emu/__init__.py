"""ctypes loader of the host-side launch emulator (tests/host_emu/emu.cc). Test infrastructure only: it is compiled with
g++ from the product's host sources (launch_params.cc, plan.cc, geometry.cc) plus the kernels' shared tile decoder
(tiling.h) into tests/host_emu/libcdb_emu.so -- a library the product never loads."""
import ctypes
import glob
import os
import subprocess

import numpy as np

from cudecomp_b200 import capi as cd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cudecomp_b200", "csrc")
LIB = os.path.join(HERE, "libcdb_emu.so")
SOURCES = [os.path.join(HERE, "emu.cc")] + [os.path.join(CSRC, f) for f in ("launch_params.cc", "plan.cc", "geometry.cc")]

_lib = None


def cuda_include():
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "include", "cuda_runtime.h")):
            return os.path.join(c, "include")
    raise RuntimeError("cuda_runtime.h not found")


def build(force=False):
    deps = SOURCES + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I" + cuda_include(), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "include", "mpi_shim"), "-I" + CSRC, "-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the launch emulator failed:\n" + res.stdout + res.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.cdb_emu_run_boxes.restype = ctypes.c_int
        _lib.cdb_emu_choose_grid.restype = ctypes.c_int
        _lib.cdb_emu_choose_grid.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
    return _lib


class EmuError(RuntimeError):
    pass


def run_boxes(boxes, srcs, dsts, es, legal, me=-1, comm_size=0, peer_index=None, tile_bytes=0, peer_order=0,
              kernel_variant=0, grid=0, threads=256):
    """boxes: list of plan dicts (capi box dicts); srcs/dsts: per box the numpy array (any dtype, C-contiguous) that
    holds the source / destination buffer; legal: numpy arrays a launch may touch. Copies on the arrays in place and
    returns the stats dict."""
    n = len(boxes)
    arr = (cd.cudecompB200Box_t * max(n, 1))()
    for i, b in enumerate(boxes):
        arr[i].peer_rank = b["peer_rank"]
        arr[i].is_unpack = b["is_unpack"]
        arr[i].src_offset = b["src_offset"]
        arr[i].dst_offset = b["dst_offset"]
        for k in range(3):
            arr[i].extent[k] = b["extent"][k]
            arr[i].src_stride[k] = b["src_stride"][k]
            arr[i].dst_stride[k] = b["dst_stride"][k]
    pi = (ctypes.c_int32 * max(n, 1))(*(peer_index if peer_index is not None else [0] * n))
    sb = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in srcs])
    db = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in dsts])
    lo = (ctypes.c_void_p * max(len(legal), 1))(*[a.ctypes.data for a in legal])
    ln = (ctypes.c_int64 * max(len(legal), 1))(*[a.nbytes for a in legal])
    stats = (ctypes.c_int64 * 8)()
    err = ctypes.create_string_buffer(512)
    rc = lib().cdb_emu_run_boxes(arr, pi, n, sb, db, lo, ln, len(legal), es, tile_bytes, peer_order, kernel_variant, me,
                                 comm_size, grid, threads, stats, err, 512)
    if rc != 0:
        raise EmuError(err.value.decode())
    return dict(launches=stats[0], bytes_written=stats[1], accesses=stats[2], kinds=stats[3], vec=stats[4], slots=stats[5],
                balanced_grid=stats[6], longest_row=stats[7])


def run_phased(boxes, srcs, dsts, es, legal, nsteps, lag, want_unpack, step, tile_bytes=0, grid=0, threads=256,
               kernel_variant=0, head_percent=25):
    """One rank's fused staged schedule (engine.cc runFusedStaged) prepared by the product's preparePhased; executes the
    push boxes of phase `step` or (want_unpack) the unpack boxes that wait for `step`. Returns the stats dict, or None
    when the schedule cannot run as one phased launch."""
    n = len(boxes)
    arr = (cd.cudecompB200Box_t * max(n, 1))()
    for i, b in enumerate(boxes):
        arr[i].peer_rank = b["peer_rank"]
        arr[i].is_unpack = b["is_unpack"]
        arr[i].step = b.get("step", 0)
        arr[i].src_offset = b["src_offset"]
        arr[i].dst_offset = b["dst_offset"]
        for k in range(3):
            arr[i].extent[k] = b["extent"][k]
            arr[i].src_stride[k] = b["src_stride"][k]
            arr[i].dst_stride[k] = b["dst_stride"][k]
    sb = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in srcs])
    db = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in dsts])
    lo = (ctypes.c_void_p * max(len(legal), 1))(*[a.ctypes.data for a in legal])
    ln = (ctypes.c_int64 * max(len(legal), 1))(*[a.nbytes for a in legal])
    stats = (ctypes.c_int64 * 8)()
    err = ctypes.create_string_buffer(512)
    fn = lib().cdb_emu_run_phased
    fn.restype = ctypes.c_int
    rc = fn(arr, n, nsteps, sb, db, lo, ln, len(legal), es, tile_bytes, lag, int(want_unpack), step, grid, threads,
            kernel_variant, head_percent, stats, err, 512)
    if rc < 0:
        raise EmuError(err.value.decode())
    if rc == 1:
        return None
    return dict(phases=stats[0], bytes_written=stats[1], accesses=stats[2], vec=stats[3], slots=stats[4], boxes=stats[5])


def aligned_array(n, dtype, offset_bytes=0, fill=None):
    """numpy array of n elements whose first byte sits `offset_bytes` past a 256-byte boundary (device allocations are
    256-byte aligned; offsets exercise the narrower vector widths)."""
    itemsize = np.dtype(dtype).itemsize
    raw = np.zeros(n * itemsize + 512, np.uint8)
    start = (-raw.ctypes.data) % 256 + offset_bytes
    out = raw[start:start + n * itemsize].view(dtype)
    if fill is not None:
        out[:] = fill
    return out

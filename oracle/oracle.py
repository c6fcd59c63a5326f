"""ctypes wrapper of the CPU oracle + the analytic known-answer pattern of the reference's tests.

TEST INFRASTRUCTURE ONLY (see the header of cudecomp_oracle.c): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg. The product package
(cudecomp_b200/) never imports this module.

Two independent checkers live here:
  * `Oracle` runs oracle/libcudecomp_oracle.so, the C restatement of the reference's
    pack -> all-to-all -> unpack algorithm (include/internal/transpose.h:196-905, halo.h:40-315);
  * `pattern_pencil` / `halo_reference` are numpy restatements of the generators the reference's own tests use
    to decide pass/fail (tests/ctest/transpose_tests.cc:323-378, tests/ctest/halo_tests.cc:158-272): every cell
    holds its global linear index, so the expected output of any transpose / halo update is known in closed form.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcudecomp_oracle.so")


def build(force=False):
    """Compile the C oracle with the Makefile next to it (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "cudecomp_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B", "libcudecomp_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Grid(ctypes.Structure):
    _fields_ = [("gdims", ctypes.c_int32 * 3), ("gdims_dist", ctypes.c_int32 * 3), ("pdims", ctypes.c_int32 * 2),
                ("col_major", ctypes.c_int32), ("order", (ctypes.c_int32 * 3) * 3)]


class _Pencil(ctypes.Structure):
    _fields_ = [("shape", ctypes.c_int32 * 3), ("lo", ctypes.c_int32 * 3), ("hi", ctypes.c_int32 * 3),
                ("order", ctypes.c_int32 * 3), ("halo", ctypes.c_int32 * 3), ("pad", ctypes.c_int32 * 3),
                ("size", ctypes.c_int64)]


class PencilInfo:
    """Plain-Python pencil description (same fields as cudecompPencilInfo_t)."""

    def __init__(self, shape, lo, hi, order, halo_extents, padding, size):
        self.shape = tuple(int(v) for v in shape)
        self.lo = tuple(int(v) for v in lo)
        self.hi = tuple(int(v) for v in hi)
        self.order = tuple(int(v) for v in order)
        self.halo_extents = tuple(int(v) for v in halo_extents)
        self.padding = tuple(int(v) for v in padding)
        self.size = int(size)

    def as_tuple(self):
        return (self.shape, self.lo, self.hi, self.order, self.halo_extents, self.padding, self.size)

    def __repr__(self):
        return "PencilInfo(shape=%s lo=%s hi=%s order=%s halo=%s pad=%s size=%d)" % self.as_tuple()


def resolve_mem_order(axis_contiguous=(False, False, False), mem_order=None):
    """order[axis][i] = global axis at memory position i (reference src/cudecomp.cc:1120-1133)."""
    if mem_order is not None and mem_order[0][0] >= 0:
        return [list(map(int, row)) for row in mem_order]
    return [[(axis + i) % 3 if axis_contiguous[axis] else i for i in range(3)] for axis in range(3)]


def _i32x3(v):
    if v is None:
        return None
    return (ctypes.c_int32 * 3)(*[int(x) for x in v])


TRANSPOSE_OPS = {"XY": (0, 1), "YZ": (1, 1), "ZY": (2, -1), "YX": (1, -1)}  # (ax, dir), transpose.h:907-953


def transpose_axes(op):
    ax, d = TRANSPOSE_OPS[op]
    return ax, (ax + 1) % 3 if d > 0 else (ax + 2) % 3


class Oracle:
    """One decomposition (global grid + process grid + memory orders), all ranks in this process."""

    def __init__(self, gdims, pdims, axis_contiguous=(False, False, False), mem_order=None, gdims_dist=None,
                 col_major=False):
        self.lib = ctypes.CDLL(build())
        L = self.lib
        L.oracle_pencil_info.argtypes = [ctypes.POINTER(_Grid), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.POINTER(_Pencil)]
        L.oracle_transpose.argtypes = [ctypes.POINTER(_Grid), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_halo.argtypes = [ctypes.POINTER(_Grid), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_shifted_rank.argtypes = [ctypes.POINTER(_Grid)] + [ctypes.c_int] * 5
        L.oracle_transpose_workspace_size.argtypes = [ctypes.POINTER(_Grid)]
        L.oracle_transpose_workspace_size.restype = ctypes.c_int64
        L.oracle_halo_workspace_size.argtypes = [ctypes.POINTER(_Grid), ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.oracle_halo_workspace_size.restype = ctypes.c_int64
        L.oracle_has_empty_pencils.argtypes = [ctypes.POINTER(_Grid), ctypes.c_int]

        self.gdims = tuple(int(v) for v in gdims)
        self.pdims = tuple(int(v) for v in pdims)
        self.nranks = self.pdims[0] * self.pdims[1]
        self.order = resolve_mem_order(axis_contiguous, mem_order)
        dist = tuple(int(v) for v in gdims_dist) if gdims_dist and all(gdims_dist) else self.gdims
        self.gdims_dist = dist
        self.col_major = bool(col_major)
        g = _Grid()
        for i in range(3):
            g.gdims[i] = self.gdims[i]
            g.gdims_dist[i] = dist[i]
            for j in range(3):
                g.order[i][j] = self.order[i][j]
        g.pdims[0], g.pdims[1] = self.pdims
        g.col_major = 1 if col_major else 0
        self._g = g

    # ---- geometry
    def pencil_info(self, rank, axis, halo=None, padding=None):
        p = _Pencil()
        rc = self.lib.oracle_pencil_info(ctypes.byref(self._g), rank, axis, _i32x3(halo), _i32x3(padding),
                                         ctypes.byref(p))
        if rc:
            raise ValueError("invalid halo/padding")
        return PencilInfo(p.shape, p.lo, p.hi, p.order, p.halo, p.pad, p.size)

    def shifted_rank(self, rank, axis, dim, displacement, periodic):
        return self.lib.oracle_shifted_rank(ctypes.byref(self._g), rank, axis, dim, displacement, int(bool(periodic)))

    def transpose_workspace_size(self):
        return self.lib.oracle_transpose_workspace_size(ctypes.byref(self._g))

    def halo_workspace_size(self, rank, axis, halo):
        return self.lib.oracle_halo_workspace_size(ctypes.byref(self._g), rank, axis, _i32x3(halo))

    def has_empty_pencils(self, axis):
        return bool(self.lib.oracle_has_empty_pencils(ctypes.byref(self._g), axis))

    # ---- data movement (all ranks at once; arrays are flat numpy arrays, one per rank)
    def transpose(self, op, inputs, outputs, in_halo=None, out_halo=None, in_pad=None, out_pad=None):
        ax, d = TRANSPOSE_OPS[op] if isinstance(op, str) else op
        es = inputs[0].dtype.itemsize
        n = self.nranks
        ins = (ctypes.c_void_p * n)(*[a.ctypes.data for a in inputs])
        outs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in outputs])
        rc = self.lib.oracle_transpose(ctypes.byref(self._g), ax, d, es, ins, outs, _i32x3(in_halo), _i32x3(out_halo),
                                       _i32x3(in_pad), _i32x3(out_pad))
        if rc:
            raise RuntimeError("oracle_transpose failed with code %d" % rc)

    def halo(self, axis, dim, data, halo, periods=None, padding=None):
        es = data[0].dtype.itemsize
        n = self.nranks
        bufs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in data])
        per = _i32x3([1 if p else 0 for p in periods]) if periods is not None else None
        rc = self.lib.oracle_halo(ctypes.byref(self._g), axis, dim, es, bufs, _i32x3(halo), per, _i32x3(padding))
        if rc:
            raise RuntimeError("oracle_halo failed with code %d" % rc)

    def max_threads(self):
        return self.lib.oracle_max_threads()

    def set_threads(self, n):
        self.lib.oracle_set_threads(int(n))

    def release(self):
        """Frees the library's persistent staging regions (they are as large as the pencils that went through it)."""
        self.lib.oracle_release()


# ------------------------------------------------------------------------------------------------------------
# Analytic known-answer pattern (numpy). dtype names follow cudecompDataType_t.

NP_DTYPES = {"float": np.float32, "double": np.float64, "float_complex": np.complex64, "double_complex": np.complex128}


def _local_coords(pinfo):
    s0, s1, s2 = pinfo.shape
    l0 = np.arange(s0, dtype=np.int64)[None, None, :]
    l1 = np.arange(s1, dtype=np.int64)[None, :, None]
    l2 = np.arange(s2, dtype=np.int64)[:, None, None]
    return [l0, l1, l2]  # arrays broadcastable to shape (s2, s1, s0): memory position 0 is the last numpy axis


def interior_mask(pinfo):
    """Cells that are neither halo nor padding (transpose_tests.cc:314-321). Shape (s2, s1, s0)."""
    loc = _local_coords(pinfo)
    m = np.ones((pinfo.shape[2], pinfo.shape[1], pinfo.shape[0]), dtype=bool)
    for k in range(3):
        h = pinfo.halo_extents[pinfo.order[k]]
        p = pinfo.padding[pinfo.order[k]]
        m &= (loc[k] >= h) & (loc[k] < pinfo.shape[k] - h - p)
    return m


def global_index(pinfo, gdims, wrap=None):
    """Global linear index gx + GX*(gy + gz*GY) of every cell, and validity.

    wrap: None -> coordinates outside the global domain are invalid;
          sequence of 3 bools -> periodic dimensions wrap (halo_tests.cc:229-253)."""
    loc = _local_coords(pinfo)
    glob = [None, None, None]
    valid = np.ones((pinfo.shape[2], pinfo.shape[1], pinfo.shape[0]), dtype=bool)
    for k in range(3):
        ax = pinfo.order[k]
        g = loc[k] + pinfo.lo[k] - pinfo.halo_extents[ax]
        outside = (g < 0) | (g >= gdims[ax])
        if wrap is not None and wrap[ax]:
            g = np.mod(g, gdims[ax])
        else:
            valid = valid & ~outside
        glob[ax] = g
    gi = glob[0] + gdims[0] * (glob[1] + glob[2] * gdims[1])
    gi = np.broadcast_to(gi, valid.shape)
    return gi, valid


def _value(gi, dtype):
    """pencilValue<T> of the reference tests: real -> gi, complex -> (gi, -gi)."""
    if np.issubdtype(dtype, np.complexfloating):
        return (gi + (-1j) * gi).astype(dtype)
    return gi.astype(dtype)


def pattern_pencil(pinfo, gdims, dtype, fill=-1):
    """initializePencil<T>: interior cells hold their global index, halo/padding cells hold `fill`. Flat array."""
    gi, _ = global_index(pinfo, gdims)
    m = interior_mask(pinfo)
    out = np.full(m.shape, fill, dtype=dtype)
    out[m] = _value(gi, dtype)[m]
    return out.reshape(-1)


def halo_reference(pinfo, gdims, dtype, periods, fill=-1):
    """initializeReference<T> of halo_tests.cc:213-236: the pencil after halos of all three dims were updated."""
    gi, valid = global_index(pinfo, gdims, wrap=periods)
    loc = _local_coords(pinfo)
    pad = np.zeros(valid.shape, dtype=bool)
    for k in range(3):
        pad |= loc[k] >= pinfo.shape[k] - pinfo.padding[pinfo.order[k]]
    out = np.full(valid.shape, fill, dtype=dtype)
    ok = valid & ~pad
    out[ok] = _value(gi, dtype)[ok]
    return out.reshape(-1)


def interior_equal(pinfo, expected, actual):
    """pencilMatches of transpose_tests.cc:356-378: exact equality on interior cells only."""
    m = interior_mask(pinfo).reshape(-1)
    return bool(np.array_equal(expected[m], actual[m]))

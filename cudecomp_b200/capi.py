"""ctypes binding of libcudecomp.so: the reference's C API under the reference's names.

Every function below calls the C entry point of the same name (include/cudecomp.h, which cites the reference
declaration each one replaces) with plain pointers and sizes; no torch types cross this boundary. Functions
return the cudecompResult_t code exactly like the C API unless stated otherwise; `check()` turns a non-zero
code into CudecompError for callers that prefer exceptions.

There is no CPU fallback: if the shared library has not been built, importing this module fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcudecomp.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "cudecomp_b200: %s is missing. Build it first (python -m cudecomp_b200.build, or __graft_entry__.build()); "
        "there is no CPU fallback for the transpose engine." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

# ---------------------------------------------------------------------------------------------- enums
CUDECOMP_TRANSPOSE_COMM_MPI_P2P = 1
CUDECOMP_TRANSPOSE_COMM_MPI_P2P_PL = 2
CUDECOMP_TRANSPOSE_COMM_MPI_A2A = 3
CUDECOMP_TRANSPOSE_COMM_NCCL = 4
CUDECOMP_TRANSPOSE_COMM_NCCL_PL = 5
CUDECOMP_TRANSPOSE_COMM_NVSHMEM = 6
CUDECOMP_TRANSPOSE_COMM_NVSHMEM_PL = 7
CUDECOMP_TRANSPOSE_COMM_NVSHMEM_SM = 8

CUDECOMP_HALO_COMM_MPI = 1
CUDECOMP_HALO_COMM_MPI_BLOCKING = 2
CUDECOMP_HALO_COMM_NCCL = 3
CUDECOMP_HALO_COMM_NVSHMEM = 4
CUDECOMP_HALO_COMM_NVSHMEM_BLOCKING = 5

CUDECOMP_FLOAT = -1
CUDECOMP_DOUBLE = -2
CUDECOMP_FLOAT_COMPLEX = -3
CUDECOMP_DOUBLE_COMPLEX = -4

CUDECOMP_AUTOTUNE_GRID_TRANSPOSE = 0
CUDECOMP_AUTOTUNE_GRID_HALO = 1

CUDECOMP_RANK_ORDER_DEFAULT = 0
CUDECOMP_RANK_ORDER_ROW_MAJOR = 1
CUDECOMP_RANK_ORDER_COL_MAJOR = 2

CUDECOMP_RESULT_SUCCESS = 0
CUDECOMP_RESULT_INVALID_USAGE = 1
CUDECOMP_RESULT_NOT_SUPPORTED = 2
CUDECOMP_RESULT_INTERNAL_ERROR = 3
CUDECOMP_RESULT_CUDA_ERROR = 4
CUDECOMP_RESULT_CUTENSOR_ERROR = 5
CUDECOMP_RESULT_MPI_ERROR = 6
CUDECOMP_RESULT_NCCL_ERROR = 7
CUDECOMP_RESULT_NVSHMEM_ERROR = 8
CUDECOMP_RESULT_NVML_ERROR = 9

RESULT_NAMES = {0: "SUCCESS", 1: "INVALID_USAGE", 2: "NOT_SUPPORTED", 3: "INTERNAL_ERROR", 4: "CUDA_ERROR",
                5: "CUTENSOR_ERROR", 6: "MPI_ERROR", 7: "NCCL_ERROR", 8: "NVSHMEM_ERROR", 9: "NVML_ERROR"}

CUDECOMP_GRID_DESC_CONFIG_VERSION = 1
CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION = 1
CUDECOMP_PENCIL_INFO_VERSION = 1

# mpi shim handles (include/mpi_shim/mpi.h)
MPI_COMM_NULL = 0
MPI_COMM_WORLD = 1
MPI_COMM_SELF = 2

PATH_NONE, PATH_LOCAL, PATH_DIRECT, PATH_STAGED = 0, 1, 2, 3

DTYPE_SIZES = {CUDECOMP_FLOAT: 4, CUDECOMP_DOUBLE: 8, CUDECOMP_FLOAT_COMPLEX: 8, CUDECOMP_DOUBLE_COMPLEX: 16}

_i32 = ctypes.c_int32
_i64 = ctypes.c_int64


# -------------------------------------------------------------------------------------------- structs
class cudecompGridDescConfig_t(ctypes.Structure):
    _fields_ = [
        ("struct_size", _i64), ("magic", _i32), ("version", _i32),
        ("gdims", _i32 * 3), ("gdims_dist", _i32 * 3), ("pdims", _i32 * 2), ("rank_order", ctypes.c_int),
        ("transpose_comm_backend", ctypes.c_int), ("transpose_axis_contiguous", ctypes.c_bool * 3),
        ("transpose_mem_order", (_i32 * 3) * 3), ("halo_comm_backend", ctypes.c_int),
    ]


class cudecompGridDescAutotuneOptions_t(ctypes.Structure):
    _fields_ = [
        ("struct_size", _i64), ("magic", _i32), ("version", _i32),
        ("n_warmup_trials", _i32), ("n_trials", _i32), ("grid_mode", ctypes.c_int), ("dtype", ctypes.c_int),
        ("allow_uneven_decompositions", ctypes.c_bool), ("disable_mpi_backends", ctypes.c_bool),
        ("disable_nccl_backends", ctypes.c_bool), ("disable_nvshmem_backends", ctypes.c_bool),
        ("skip_threshold", ctypes.c_double),
        ("autotune_transpose_backend", ctypes.c_bool), ("transpose_use_inplace_buffers", ctypes.c_bool * 4),
        ("transpose_op_weights", ctypes.c_double * 4),
        ("transpose_input_halo_extents", (_i32 * 3) * 4), ("transpose_output_halo_extents", (_i32 * 3) * 4),
        ("transpose_input_padding", (_i32 * 3) * 4), ("transpose_output_padding", (_i32 * 3) * 4),
        ("autotune_halo_backend", ctypes.c_bool), ("halo_extents", _i32 * 3), ("halo_periods", ctypes.c_bool * 3),
        ("halo_axis", _i32), ("halo_padding", _i32 * 3),
    ]


class cudecompPencilInfo_t(ctypes.Structure):
    _fields_ = [
        ("struct_size", _i64), ("magic", _i32), ("version", _i32),
        ("shape", _i32 * 3), ("lo", _i32 * 3), ("hi", _i32 * 3), ("order", _i32 * 3),
        ("halo_extents", _i32 * 3), ("padding", _i32 * 3), ("size", _i64),
    ]


class cudecompB200Box_t(ctypes.Structure):
    _fields_ = [("peer_rank", _i32), ("is_unpack", _i32), ("step", _i32), ("reserved", _i32),
                ("src_offset", _i64), ("dst_offset", _i64),
                ("extent", _i64 * 3), ("src_stride", _i64 * 3), ("dst_stride", _i64 * 3)]


assert ctypes.sizeof(cudecompGridDescConfig_t) == 104
assert ctypes.sizeof(cudecompGridDescAutotuneOptions_t) == 320
assert ctypes.sizeof(cudecompPencilInfo_t) == 96

cudecompHandle_t = ctypes.c_void_p
cudecompGridDesc_t = ctypes.c_void_p

# Every symbol include/cudecomp.h declares (24) + include/cudecomp_b200_ext.h + the mpi shim.
API_SYMBOLS = [
    "cudecompInit", "cudecompInit_F", "cudecompFinalize", "cudecompGridDescCreateVersioned",
    "cudecompGridDescDestroy", "cudecompGridDescConfigSetDefaultsVersioned",
    "cudecompGridDescAutotuneOptionsSetDefaultsVersioned", "cudecompGetPencilInfoVersioned",
    "cudecompGetTransposeWorkspaceSize", "cudecompGetHaloWorkspaceSize", "cudecompGetDataTypeSize", "cudecompMalloc",
    "cudecompFree", "cudecompTransposeCommBackendToString", "cudecompHaloCommBackendToString",
    "cudecompGetGridDescConfigVersioned", "cudecompGetShiftedRank", "cudecompTransposeXToY", "cudecompTransposeYToZ",
    "cudecompTransposeZToY", "cudecompTransposeYToX", "cudecompUpdateHalosX", "cudecompUpdateHalosY",
    "cudecompUpdateHalosZ",
]
EXT_SYMBOLS = ["cudecompB200GetLaunchCount", "cudecompB200GetLastPath", "cudecompB200SetTuning",
               "cudecompB200CheckErrors", "cudecompB200SetPipelineChunks", "cudecompB200SetStagedMode", "cudecompB200SetKernelVariant", "cudecompB200DescribeTransposeBoxes", "cudecompB200DescribeHaloBoxes",
               "cudecompB200PlanTransposeBoxes", "cudecompB200PlanHaloBoxes", "cudecompB200PlanPipelinedTransposeBoxes",
               "cudecompB200SelfTestMailbox", "cudecompB200GetAutotuneCandidates", "cudecompB200GetCumemState",
               "cudecompB200ProbeFdPassing"]
MPI_SHIM_SYMBOLS = ["MPI_Init", "MPI_Init_thread", "MPI_Initialized", "MPI_Finalize", "MPI_Finalized", "MPI_Abort",
                    "MPI_Wtime", "MPI_Get_processor_name", "MPI_Error_string", "MPI_Comm_rank", "MPI_Comm_size",
                    "MPI_Comm_split", "MPI_Comm_split_type", "MPI_Comm_dup", "MPI_Comm_free", "MPI_Comm_c2f",
                    "MPI_Comm_f2c", "MPI_Barrier", "MPI_Bcast", "MPI_Allgather", "MPI_Gather", "MPI_Allreduce",
                    "MPI_Reduce"]

_P = ctypes.POINTER
_vp = ctypes.c_void_p
_i32p = _P(_i32)


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_sig("cudecompInit", ctypes.c_int, [_P(cudecompHandle_t), ctypes.c_int])
_sig("cudecompInit_F", ctypes.c_int, [_P(cudecompHandle_t), ctypes.c_int])
_sig("cudecompFinalize", ctypes.c_int, [cudecompHandle_t])
_sig("cudecompGridDescCreateVersioned", ctypes.c_int,
     [cudecompHandle_t, _P(cudecompGridDesc_t), _P(cudecompGridDescConfig_t), _i64, _i32,
      _P(cudecompGridDescAutotuneOptions_t), _i64, _i32])
_sig("cudecompGridDescDestroy", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t])
_sig("cudecompGridDescConfigSetDefaultsVersioned", ctypes.c_int, [_P(cudecompGridDescConfig_t), _i64, _i32])
_sig("cudecompGridDescAutotuneOptionsSetDefaultsVersioned", ctypes.c_int,
     [_P(cudecompGridDescAutotuneOptions_t), _i64, _i32])
_sig("cudecompGetPencilInfoVersioned", ctypes.c_int,
     [cudecompHandle_t, cudecompGridDesc_t, _P(cudecompPencilInfo_t), _i64, _i32, _i32, _i32p, _i32p])
_sig("cudecompGetTransposeWorkspaceSize", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _P(_i64)])
_sig("cudecompGetHaloWorkspaceSize", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32p, _P(_i64)])
_sig("cudecompGetDataTypeSize", ctypes.c_int, [ctypes.c_int, _P(_i64)])
_sig("cudecompMalloc", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _P(_vp), ctypes.c_size_t])
_sig("cudecompFree", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _vp])
_sig("cudecompTransposeCommBackendToString", ctypes.c_char_p, [ctypes.c_int])
_sig("cudecompHaloCommBackendToString", ctypes.c_char_p, [ctypes.c_int])
_sig("cudecompGetGridDescConfigVersioned", ctypes.c_int,
     [cudecompHandle_t, cudecompGridDesc_t, _P(cudecompGridDescConfig_t), _i64, _i32])
_sig("cudecompGetShiftedRank", ctypes.c_int,
     [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32, _i32, ctypes.c_bool, _i32p])
for _n in ("cudecompTransposeXToY", "cudecompTransposeYToZ", "cudecompTransposeZToY", "cudecompTransposeYToX"):
    _sig(_n, ctypes.c_int,
         [cudecompHandle_t, cudecompGridDesc_t, _vp, _vp, _vp, ctypes.c_int, _i32p, _i32p, _i32p, _i32p, _vp])
for _n in ("cudecompUpdateHalosX", "cudecompUpdateHalosY", "cudecompUpdateHalosZ"):
    _sig(_n, ctypes.c_int,
         [cudecompHandle_t, cudecompGridDesc_t, _vp, _vp, ctypes.c_int, _i32p, _P(ctypes.c_bool), _i32, _i32p, _vp])
_sig("cudecompB200GetLaunchCount", ctypes.c_int, [_P(ctypes.c_uint64)])
_sig("cudecompB200GetLastPath", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32p])
_sig("cudecompB200GetCumemState", ctypes.c_int, [cudecompHandle_t, _i32p])
_sig("cudecompB200ProbeFdPassing", ctypes.c_int, [cudecompHandle_t, _i32p])
_sig("cudecompB200SetTuning", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32])
_sig("cudecompB200CheckErrors", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t])
_sig("cudecompB200SetPipelineChunks", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32])
_sig("cudecompB200SetStagedMode", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32])
_sig("cudecompB200SetKernelVariant", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32])
_sig("cudecompB200SetSchedule", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32, _i32])
_sig("cudecompB200SetTransferMode", ctypes.c_int, [cudecompHandle_t, cudecompGridDesc_t, _i32])
_sig("cudecompB200DescribeTransposeBoxes", _i32,
     [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32, _i32p, _i32p, _i32p, _i32p, _i32, _P(cudecompB200Box_t), _i32])
_sig("cudecompB200DescribeHaloBoxes", _i32,
     [cudecompHandle_t, cudecompGridDesc_t, _i32, _i32, _i32p, _P(ctypes.c_bool), _i32p, _i32,
      _P(cudecompB200Box_t), _i32])
_sig("cudecompB200PlanTransposeBoxes", _i32,
     [_P(cudecompGridDescConfig_t), _i32, _i32, _i32, _i32p, _i32p, _i32p, _i32p, _i32, _P(cudecompB200Box_t), _i32])
_sig("cudecompB200PlanHaloBoxes", _i32,
     [_P(cudecompGridDescConfig_t), _i32, _i32, _i32, _i32p, _P(ctypes.c_bool), _i32p, _i32, _P(cudecompB200Box_t), _i32])
_sig("cudecompB200PlanPipelinedTransposeBoxes", _i32,
     [_P(cudecompGridDescConfig_t), _i32, _i32, _i32, _i32p, _i32p, _i32p, _i32p, _i32, _i32, _P(cudecompB200Box_t), _i32])
_sig("cudecompB200SelfTestMailbox", ctypes.c_int, [cudecompHandle_t, _i32, ctypes.c_uint32])
_sig("MPI_Init", ctypes.c_int, [_vp, _vp])
_sig("MPI_Finalize", ctypes.c_int, [])
_sig("MPI_Comm_rank", ctypes.c_int, [ctypes.c_int, _P(ctypes.c_int)])
_sig("MPI_Comm_size", ctypes.c_int, [ctypes.c_int, _P(ctypes.c_int)])
_sig("MPI_Barrier", ctypes.c_int, [ctypes.c_int])
_sig("MPI_Comm_split", ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P(ctypes.c_int)])
_sig("MPI_Comm_free", ctypes.c_int, [_P(ctypes.c_int)])
_sig("MPI_Allreduce", ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_sig("MPI_Bcast", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_sig("MPI_Allgather", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int])
_sig("MPI_Wtime", ctypes.c_double, [])


class CudecompError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        super().__init__("%s failed: CUDECOMP_RESULT_%s (%d)" % (where or "cuDecomp call", RESULT_NAMES.get(code, "?"),
                                                                code))


def check(code, where=""):
    if code != CUDECOMP_RESULT_SUCCESS:
        raise CudecompError(code, where)
    return code


def _arr3(v):
    """None -> NULL; sequence -> int32[3]."""
    if v is None:
        return None
    return (_i32 * 3)(*[int(x) for x in v])


def _bool3(v):
    if v is None:
        return None
    return (ctypes.c_bool * 3)(*[bool(x) for x in v])


def _ptr(p):
    """Accept ints, c_void_p and objects with data_ptr() (torch tensors)."""
    if p is None:
        return None
    if hasattr(p, "data_ptr"):
        return ctypes.c_void_p(p.data_ptr())
    if isinstance(p, ctypes.c_void_p):
        return p
    return ctypes.c_void_p(int(p))


def _stream(s):
    if s is None:
        return None
    if hasattr(s, "cuda_stream"):
        return ctypes.c_void_p(s.cuda_stream)
    return ctypes.c_void_p(int(s))


# ------------------------------------------------------------------------------- mpi shim (bootstrap)
def MPI_Init():
    return lib.MPI_Init(None, None)


def MPI_Finalize():
    return lib.MPI_Finalize()


def MPI_Comm_rank(comm=MPI_COMM_WORLD):
    r = ctypes.c_int(-1)
    lib.MPI_Comm_rank(comm, ctypes.byref(r))
    return r.value


def MPI_Comm_size(comm=MPI_COMM_WORLD):
    r = ctypes.c_int(-1)
    lib.MPI_Comm_size(comm, ctypes.byref(r))
    return r.value


def MPI_Barrier(comm=MPI_COMM_WORLD):
    return lib.MPI_Barrier(comm)


def MPI_Allreduce_max(value, comm=MPI_COMM_WORLD):
    """max over the ranks of one double (MPI_DOUBLE = (3 << 8) | 8, MPI_MAX = 2 in include/mpi_shim/mpi.h)."""
    v = ctypes.c_double(float(value))
    lib.MPI_Allreduce(ctypes.c_void_p(-1), ctypes.byref(v), 1, (3 << 8) | 8, 2, comm)  # MPI_IN_PLACE
    return v.value


def MPI_Comm_split(comm, color, key):
    out = ctypes.c_int(0)
    lib.MPI_Comm_split(comm, color, key, ctypes.byref(out))
    return out.value


# ----------------------------------------------------------------- the reference API, same names/semantics
def cudecompInit(mpi_comm=MPI_COMM_WORLD):
    """-> (result, handle)"""
    h = cudecompHandle_t()
    res = lib.cudecompInit(ctypes.byref(h), mpi_comm)
    return res, h


def cudecompFinalize(handle):
    return lib.cudecompFinalize(handle)


def cudecompGridDescConfigSetDefaults(config):
    return lib.cudecompGridDescConfigSetDefaultsVersioned(ctypes.byref(config), ctypes.sizeof(config),
                                                          CUDECOMP_GRID_DESC_CONFIG_VERSION)


def cudecompGridDescAutotuneOptionsSetDefaults(options):
    return lib.cudecompGridDescAutotuneOptionsSetDefaultsVersioned(ctypes.byref(options), ctypes.sizeof(options),
                                                                   CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION)


def cudecompGridDescCreate(handle, config, options=None):
    """-> (result, grid_desc). `config` is updated in place like in C."""
    gd = cudecompGridDesc_t()
    res = lib.cudecompGridDescCreateVersioned(
        handle, ctypes.byref(gd), ctypes.byref(config), ctypes.sizeof(config), CUDECOMP_GRID_DESC_CONFIG_VERSION,
        ctypes.byref(options) if options is not None else None,
        ctypes.sizeof(options) if options is not None else 0,
        CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION if options is not None else 0)
    return res, gd


def cudecompGridDescDestroy(handle, grid_desc):
    return lib.cudecompGridDescDestroy(handle, grid_desc)


def cudecompGetPencilInfo(handle, grid_desc, axis, halo_extents=None, padding=None):
    """-> (result, cudecompPencilInfo_t)"""
    p = cudecompPencilInfo_t()
    res = lib.cudecompGetPencilInfoVersioned(handle, grid_desc, ctypes.byref(p), ctypes.sizeof(p),
                                             CUDECOMP_PENCIL_INFO_VERSION, axis, _arr3(halo_extents), _arr3(padding))
    return res, p


def cudecompGetGridDescConfig(handle, grid_desc):
    c = cudecompGridDescConfig_t()
    res = lib.cudecompGetGridDescConfigVersioned(handle, grid_desc, ctypes.byref(c), ctypes.sizeof(c),
                                                 CUDECOMP_GRID_DESC_CONFIG_VERSION)
    return res, c


def cudecompGetTransposeWorkspaceSize(handle, grid_desc):
    n = _i64(0)
    res = lib.cudecompGetTransposeWorkspaceSize(handle, grid_desc, ctypes.byref(n))
    return res, n.value


def cudecompGetHaloWorkspaceSize(handle, grid_desc, axis, halo_extents):
    n = _i64(0)
    res = lib.cudecompGetHaloWorkspaceSize(handle, grid_desc, axis, _arr3(halo_extents), ctypes.byref(n))
    return res, n.value


def cudecompGetDataTypeSize(dtype):
    n = _i64(0)
    res = lib.cudecompGetDataTypeSize(dtype, ctypes.byref(n))
    return res, n.value


def cudecompMalloc(handle, grid_desc, nbytes):
    """-> (result, device pointer as int)"""
    p = ctypes.c_void_p()
    res = lib.cudecompMalloc(handle, grid_desc, ctypes.byref(p), nbytes)
    return res, (p.value or 0)


def cudecompFree(handle, grid_desc, ptr):
    return lib.cudecompFree(handle, grid_desc, _ptr(ptr))


def cudecompTransposeCommBackendToString(backend):
    return lib.cudecompTransposeCommBackendToString(backend).decode()


def cudecompHaloCommBackendToString(backend):
    return lib.cudecompHaloCommBackendToString(backend).decode()


def cudecompGetShiftedRank(handle, grid_desc, axis, dim, displacement, periodic):
    r = _i32(-2)
    res = lib.cudecompGetShiftedRank(handle, grid_desc, axis, dim, displacement, bool(periodic), ctypes.byref(r))
    return res, r.value


def _transpose(fn):
    def call(handle, grid_desc, input, output, work, dtype, input_halo_extents=None, output_halo_extents=None,
             input_padding=None, output_padding=None, stream=None):
        return fn(handle, grid_desc, _ptr(input), _ptr(output), _ptr(work), dtype, _arr3(input_halo_extents),
                  _arr3(output_halo_extents), _arr3(input_padding), _arr3(output_padding), _stream(stream))
    return call


cudecompTransposeXToY = _transpose(lib.cudecompTransposeXToY)
cudecompTransposeYToZ = _transpose(lib.cudecompTransposeYToZ)
cudecompTransposeZToY = _transpose(lib.cudecompTransposeZToY)
cudecompTransposeYToX = _transpose(lib.cudecompTransposeYToX)
TRANSPOSES = {"XY": cudecompTransposeXToY, "YZ": cudecompTransposeYToZ, "ZY": cudecompTransposeZToY,
              "YX": cudecompTransposeYToX}


def _halo(fn):
    def call(handle, grid_desc, input, work, dtype, halo_extents, halo_periods, dim, padding=None, stream=None):
        return fn(handle, grid_desc, _ptr(input), _ptr(work), dtype, _arr3(halo_extents), _bool3(halo_periods), dim,
                  _arr3(padding), _stream(stream))
    return call


cudecompUpdateHalosX = _halo(lib.cudecompUpdateHalosX)
cudecompUpdateHalosY = _halo(lib.cudecompUpdateHalosY)
cudecompUpdateHalosZ = _halo(lib.cudecompUpdateHalosZ)
UPDATE_HALOS = [cudecompUpdateHalosX, cudecompUpdateHalosY, cudecompUpdateHalosZ]


# --------------------------------------------------------------------------------------------- extensions
def launch_count():
    n = ctypes.c_uint64(0)
    lib.cudecompB200GetLaunchCount(ctypes.byref(n))
    return n.value


def cumem_state(handle):
    """Outcome of CUDECOMP_ENABLE_CUMEM on this handle (cudecomp_b200_ext.h): 0 off, 1 on, 2 no fd passing, 3 no device support."""
    p = _i32(0)
    check(lib.cudecompB200GetCumemState(handle, ctypes.byref(p)), "cudecompB200GetCumemState")
    return p.value


def probe_fd_passing(handle):
    """Collective, host only: 1 when every rank can duplicate a file descriptor of its neighbour (pidfd_getfd)."""
    p = _i32(0)
    check(lib.cudecompB200ProbeFdPassing(handle, ctypes.byref(p)), "cudecompB200ProbeFdPassing")
    return p.value


def last_path(handle, grid_desc):
    p = _i32(0)
    check(lib.cudecompB200GetLastPath(handle, grid_desc, ctypes.byref(p)), "cudecompB200GetLastPath")
    return p.value


def set_tuning(handle, grid_desc, grid_ctas=0, force_staged=False):
    return lib.cudecompB200SetTuning(handle, grid_desc, int(grid_ctas), 1 if force_staged else 0)


def set_kernel_variant(handle, grid_desc, variant):
    return lib.cudecompB200SetKernelVariant(handle, grid_desc, int(variant))


def set_schedule(handle, grid_desc, tile_bytes=0, peer_order=0, balance_grid=False):
    return lib.cudecompB200SetSchedule(handle, grid_desc, int(tile_bytes), int(peer_order), 1 if balance_grid else 0)


def set_transfer_mode(handle, grid_desc, mode):
    return lib.cudecompB200SetTransferMode(handle, grid_desc, int(mode))


def set_pipeline_chunks(handle, grid_desc, nchunks):
    return lib.cudecompB200SetPipelineChunks(handle, grid_desc, int(nchunks))


def set_staged_mode(handle, grid_desc, mode, lag=0):
    """0: staged transposes run as ONE phased launch (default), 1: separate push / unpack launches; lag 0 keeps the value."""
    return lib.cudecompB200SetStagedMode(handle, grid_desc, int(mode), int(lag))


def check_errors(handle, grid_desc):
    return lib.cudecompB200CheckErrors(handle, grid_desc)


def _boxes(n, arr):
    out = []
    for i in range(n):
        b = arr[i]
        out.append(dict(peer_rank=b.peer_rank, is_unpack=bool(b.is_unpack), step=b.step, src_offset=b.src_offset,
                        dst_offset=b.dst_offset, extent=tuple(b.extent), src_stride=tuple(b.src_stride),
                        dst_stride=tuple(b.dst_stride)))
    return out


def describe_transpose_boxes(handle, grid_desc, ax, direction, input_halo_extents=None, output_halo_extents=None,
                             input_padding=None, output_padding=None, staged=False, max_boxes=256):
    arr = (cudecompB200Box_t * max_boxes)()
    n = lib.cudecompB200DescribeTransposeBoxes(handle, grid_desc, ax, direction, _arr3(input_halo_extents),
                                               _arr3(output_halo_extents), _arr3(input_padding),
                                               _arr3(output_padding), 1 if staged else 0, arr, max_boxes)
    if n < 0:
        raise CudecompError(CUDECOMP_RESULT_INVALID_USAGE, "cudecompB200DescribeTransposeBoxes")
    return _boxes(n, arr)


def describe_halo_boxes(handle, grid_desc, ax, dim, halo_extents, halo_periods=None, padding=None, staged=False,
                        max_boxes=16):
    arr = (cudecompB200Box_t * max_boxes)()
    n = lib.cudecompB200DescribeHaloBoxes(handle, grid_desc, ax, dim, _arr3(halo_extents), _bool3(halo_periods),
                                          _arr3(padding), 1 if staged else 0, arr, max_boxes)
    if n < 0:
        raise CudecompError(CUDECOMP_RESULT_INVALID_USAGE, "cudecompB200DescribeHaloBoxes")
    return _boxes(n, arr)


def plan_transpose_boxes(config, rank, ax, direction, input_halo_extents=None, output_halo_extents=None,
                         input_padding=None, output_padding=None, staged=False, max_boxes=256):
    """Handle-free planner (any rank of any process grid). Raises CudecompError with the library's result code."""
    arr = (cudecompB200Box_t * max_boxes)()
    n = lib.cudecompB200PlanTransposeBoxes(ctypes.byref(config), rank, ax, direction, _arr3(input_halo_extents),
                                           _arr3(output_halo_extents), _arr3(input_padding), _arr3(output_padding),
                                           int(staged), arr, max_boxes)
    if n < 0:
        raise CudecompError(-n, "cudecompB200PlanTransposeBoxes")
    return _boxes(n, arr)


def plan_halo_boxes(config, rank, ax, dim, halo_extents, halo_periods=None, padding=None, staged=False, max_boxes=16):
    arr = (cudecompB200Box_t * max_boxes)()
    n = lib.cudecompB200PlanHaloBoxes(ctypes.byref(config), rank, ax, dim, _arr3(halo_extents), _bool3(halo_periods),
                                      _arr3(padding), 1 if staged else 0, arr, max_boxes)
    if n < 0:
        raise CudecompError(-n, "cudecompB200PlanHaloBoxes")
    return _boxes(n, arr)


def plan_pipelined_transpose_boxes(config, rank, ax, direction, input_halo_extents=None, output_halo_extents=None,
                                   input_padding=None, output_padding=None, inplace=False, nchunks=4, max_boxes=8192):
    """Chunked staged schedule (boxes carry their step). Empty list when chunking does not apply."""
    arr = (cudecompB200Box_t * max_boxes)()
    n = lib.cudecompB200PlanPipelinedTransposeBoxes(ctypes.byref(config), rank, ax, direction,
                                                    _arr3(input_halo_extents), _arr3(output_halo_extents),
                                                    _arr3(input_padding), _arr3(output_padding), int(inplace),
                                                    nchunks, arr, max_boxes)
    if n < 0:
        raise CudecompError(-n, "cudecompB200PlanPipelinedTransposeBoxes")
    if n > max_boxes:
        raise ValueError("plan has %d boxes, buffer holds %d" % (n, max_boxes))
    return _boxes(n, arr)


def autotune_candidates(options, nranks=1, rank_order=CUDECOMP_RANK_ORDER_ROW_MAJOR):
    """-> (result, transpose backends, halo backends, pdims candidates) after environment filters and disable flags."""
    tb, hb = (_i32 * 8)(), (_i32 * 5)()
    nt, nh, npd = _i32(0), _i32(0), _i32(0)
    pd = ((_i32 * 2) * 64)()
    fn = lib.cudecompB200GetAutotuneCandidates
    fn.restype = ctypes.c_int
    res = fn(ctypes.byref(options), _i32(nranks), ctypes.c_int(rank_order), tb, ctypes.byref(nt), hb, ctypes.byref(nh), pd,
             _i32(64), ctypes.byref(npd))
    return res, list(tb[:nt.value]), list(hb[:nh.value]), [tuple(pd[i]) for i in range(npd.value)]

"""C-ABI surface of libcudecomp.so on a machine without a GPU: symbols, struct layouts, defaults, argument
checking and error codes (reference tests/ctest/api_tests.cc:254-317,449-493 and the *RejectsInvalid* cases)."""
import ctypes
import os
import re

import pytest

from cudecomp_b200 import capi as cd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(cudecomp\w+|MPI_\w+)\s*\(", src))
    # drop macros / inline wrappers defined in the header itself
    inline = set(re.findall(r"static inline \w+ (\w+)\(", src))
    return sorted(n for n in names - inline if not n.isupper() and not n.endswith("_t"))


@pytest.mark.parametrize("header", ["cudecomp.h", "cudecomp_b200_ext.h", "cudecomp_b200_mpi.h", "mpi_shim/mpi.h"])
def test_library_exports_every_declared_symbol(header):
    names = declared_functions(header)
    assert len(names) >= 6
    missing = [n for n in names if not hasattr(cd.lib, n)]
    assert not missing, missing


def test_reference_abi_has_24_entry_points():
    names = declared_functions("cudecomp.h")
    assert sorted(names) == sorted(cd.API_SYMBOLS) and len(names) == 24


def test_struct_sizes_match_reference_abi():
    # reference src/cudecomp.cc:216,242,268
    assert ctypes.sizeof(cd.cudecompGridDescConfig_t) == 104
    assert ctypes.sizeof(cd.cudecompGridDescAutotuneOptions_t) == 320
    assert ctypes.sizeof(cd.cudecompPencilInfo_t) == 96
    assert cd.cudecompGridDescConfig_t.transpose_mem_order.offset == 60
    assert cd.cudecompGridDescAutotuneOptions_t.transpose_op_weights.offset == 56
    assert cd.cudecompPencilInfo_t.size.offset == 88


def test_config_defaults():
    c = cd.cudecompGridDescConfig_t()
    assert cd.cudecompGridDescConfigSetDefaults(c) == 0
    assert c.struct_size == 104 and c.magic == 0x434f4e46 and c.version == 1
    assert list(c.gdims) == [0, 0, 0] and list(c.gdims_dist) == [0, 0, 0] and list(c.pdims) == [0, 0]
    assert c.transpose_comm_backend == cd.CUDECOMP_TRANSPOSE_COMM_MPI_P2P
    assert c.halo_comm_backend == cd.CUDECOMP_HALO_COMM_MPI
    assert c.rank_order == cd.CUDECOMP_RANK_ORDER_DEFAULT
    assert [list(r) for r in c.transpose_mem_order] == [[-1] * 3] * 3
    assert list(c.transpose_axis_contiguous) == [False] * 3


def test_autotune_option_defaults():
    o = cd.cudecompGridDescAutotuneOptions_t()
    assert cd.cudecompGridDescAutotuneOptionsSetDefaults(o) == 0
    assert o.struct_size == 320 and o.magic == 0x4155544f and o.version == 1
    assert (o.n_warmup_trials, o.n_trials) == (3, 5)
    assert o.grid_mode == cd.CUDECOMP_AUTOTUNE_GRID_TRANSPOSE and o.dtype == cd.CUDECOMP_DOUBLE
    assert o.allow_uneven_decompositions and not o.disable_mpi_backends and not o.disable_nccl_backends
    assert not o.disable_nvshmem_backends and o.skip_threshold == 0.0
    assert not o.autotune_transpose_backend and not o.autotune_halo_backend
    assert list(o.transpose_use_inplace_buffers) == [False] * 4 and list(o.transpose_op_weights) == [1.0] * 4
    assert list(o.halo_extents) == [0, 0, 0] and o.halo_axis == 0


def test_set_defaults_reject_bad_arguments():
    c = cd.cudecompGridDescConfig_t()
    L = cd.lib
    assert L.cudecompGridDescConfigSetDefaultsVersioned(None, 104, 1) == cd.CUDECOMP_RESULT_INVALID_USAGE
    assert L.cudecompGridDescConfigSetDefaultsVersioned(ctypes.byref(c), 100, 1) == cd.CUDECOMP_RESULT_INVALID_USAGE
    assert L.cudecompGridDescConfigSetDefaultsVersioned(ctypes.byref(c), 104, 2) == cd.CUDECOMP_RESULT_INVALID_USAGE
    assert L.cudecompGridDescAutotuneOptionsSetDefaultsVersioned(None, 320, 1) == cd.CUDECOMP_RESULT_INVALID_USAGE


def test_dtype_sizes_and_backend_names(golden):
    for name, size in golden["dtype_sizes"]:
        assert cd.cudecompGetDataTypeSize(getattr(cd, name)) == (0, size)
    assert cd.cudecompGetDataTypeSize(999)[0] == cd.CUDECOMP_RESULT_INVALID_USAGE
    assert cd.lib.cudecompGetDataTypeSize(cd.CUDECOMP_FLOAT, None) == cd.CUDECOMP_RESULT_INVALID_USAGE
    for name, text in golden["transpose_backend_names"]:
        assert cd.cudecompTransposeCommBackendToString(getattr(cd, name)) == text
    for name, text in golden["halo_backend_names"]:
        assert cd.cudecompHaloCommBackendToString(getattr(cd, name)) == text
    assert cd.cudecompTransposeCommBackendToString(999) == "ERROR"
    assert cd.cudecompHaloCommBackendToString(999) == "ERROR"


@pytest.fixture(scope="module")
def handle():
    os.environ.pop("WORLD_SIZE", None)
    os.environ.pop("RANK", None)
    assert cd.MPI_Init() == 0
    res, h = cd.cudecompInit(cd.MPI_COMM_WORLD)
    assert res == 0
    yield h
    assert cd.cudecompFinalize(h) == 0


def _config(gdims=(9, 10, 11), pdims=(1, 1)):
    c = cd.cudecompGridDescConfig_t()
    cd.cudecompGridDescConfigSetDefaults(c)
    c.gdims[:] = gdims
    c.pdims[:] = pdims
    return c


def test_grid_desc_create_validation(handle):
    INV = cd.CUDECOMP_RESULT_INVALID_USAGE
    raw = cd.cudecompGridDescConfig_t()  # never initialised by SetDefaults
    assert cd.cudecompGridDescCreate(handle, raw)[0] == INV
    assert cd.cudecompGridDescCreate(None, _config())[0] == INV
    c = _config(pdims=(2, 1))
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV  # product != nranks
    c = _config(pdims=(0, 0))
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV  # autotune without options
    c = _config()
    c.transpose_comm_backend = 99
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV
    c = _config()
    c.transpose_mem_order[0][0] = 0  # partially set
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV
    c = _config()
    for i in range(3):
        c.transpose_mem_order[i][:] = [0, 0, 2]
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV
    c = _config()
    c.gdims_dist[:] = [10, 10, 11]
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV
    c = _config()
    c.version = 2
    assert cd.cudecompGridDescCreate(handle, c)[0] == INV


def test_single_rank_queries_and_config_roundtrip(handle):
    c = _config()
    c.transpose_axis_contiguous[1] = True
    res, gd = cd.cudecompGridDescCreate(handle, c)
    assert res == 0
    # unset fields come back unset (reference src/cudecomp.cc:1250-1265)
    assert list(c.gdims_dist) == [0, 0, 0] and [list(r) for r in c.transpose_mem_order] == [[-1] * 3] * 3
    res, c2 = cd.cudecompGetGridDescConfig(handle, gd)
    assert res == 0 and list(c2.pdims) == [1, 1] and c2.rank_order == cd.CUDECOMP_RANK_ORDER_ROW_MAJOR
    res, p = cd.cudecompGetPencilInfo(handle, gd, 1, (1, 2, 1), (1, 0, 2))
    assert res == 0 and list(p.order) == [1, 2, 0] and list(p.shape) == [14, 15, 12] and p.size == 14 * 15 * 12
    assert p.magic == 0x50494e46 and p.struct_size == 96
    INV = cd.CUDECOMP_RESULT_INVALID_USAGE
    assert cd.cudecompGetPencilInfo(handle, gd, 3)[0] == INV
    assert cd.cudecompGetPencilInfo(handle, gd, 0, (-1, 0, 0))[0] == INV
    assert cd.cudecompGetPencilInfo(handle, gd, 0, None, (0, -2, 0))[0] == INV
    assert cd.cudecompGetPencilInfo(handle, gd, 0, (2 ** 30, 0, 0))[0] == INV  # shape overflow
    assert cd.cudecompGetHaloWorkspaceSize(handle, gd, 0, None)[0] == INV
    assert cd.cudecompGetHaloWorkspaceSize(handle, gd, 3, (1, 1, 1))[0] == INV
    assert cd.cudecompGetShiftedRank(handle, gd, 0, 0, 1, True) == (0, 0)
    assert cd.cudecompGetShiftedRank(handle, gd, 0, 1, 1, False) == (0, -1)
    assert cd.cudecompGetShiftedRank(handle, gd, 0, 3, 1, False)[0] == INV
    # transposes validate before touching pointers
    assert cd.cudecompTransposeXToY(handle, gd, None, 8, 8, cd.CUDECOMP_FLOAT) == INV
    assert cd.cudecompTransposeXToY(handle, gd, 8, 8, 8, 123) == INV
    assert cd.cudecompUpdateHalosX(handle, gd, None, None, cd.CUDECOMP_FLOAT, (0, 0, 0), None, 0) == 0
    assert cd.cudecompUpdateHalosX(handle, gd, 8, 8, cd.CUDECOMP_FLOAT, (1, 0, 0), None, 5) == INV
    assert cd.cudecompGridDescDestroy(handle, gd) == 0
    # a destroyed descriptor is recognised as stale (never dereferenced): every entry point rejects it
    assert cd.cudecompGridDescDestroy(handle, gd) == INV
    assert cd.cudecompGetPencilInfo(handle, gd, 0)[0] == INV
    assert cd.cudecompGetTransposeWorkspaceSize(handle, gd)[0] == INV


def test_empty_pencils_not_supported(handle):
    # reference api_tests.cc:1493-1505: checked before pointers are touched
    c = _config(gdims=(0 + 4, 4, 4))
    res, gd = cd.cudecompGridDescCreate(handle, c)
    assert res == 0
    cd.cudecompGridDescDestroy(handle, gd)


def test_product_fails_loudly_without_gpu(handle):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = _config()
    c.transpose_axis_contiguous[0] = True
    res, gd = cd.cudecompGridDescCreate(handle, c)
    assert res == 0
    buf = (ctypes.c_char * 8192)()
    addr = ctypes.addressof(buf)
    # out of place on one rank would need a kernel: must report a CUDA error, never fall back to the CPU
    assert cd.cudecompTransposeXToY(handle, gd, addr, addr + 4096, addr, cd.CUDECOMP_FLOAT) == cd.CUDECOMP_RESULT_CUDA_ERROR
    assert cd.cudecompMalloc(handle, gd, 1024)[0] == cd.CUDECOMP_RESULT_CUDA_ERROR
    cd.cudecompGridDescDestroy(handle, gd)


def test_schedule_knobs_validate_their_arguments(handle):
    res, gd = cd.cudecompGridDescCreate(handle, _config())
    assert res == 0
    INV = cd.CUDECOMP_RESULT_INVALID_USAGE
    assert cd.set_schedule(handle, gd, 0, 0, False) == 0
    assert cd.set_schedule(handle, gd, 16384, 1, True) == 0
    assert cd.set_schedule(handle, gd, 3000, 0, False) == INV      # not a power of two
    assert cd.set_schedule(handle, gd, 2048, 0, False) == INV      # below 4 KiB
    assert cd.set_schedule(handle, gd, 1 << 20, 0, False) == INV   # above 256 KiB
    assert cd.set_schedule(handle, gd, 0, 2, False) == INV
    assert cd.set_schedule(None, gd, 0, 0, False) == INV
    assert cd.set_kernel_variant(handle, gd, 3) == 0 and cd.set_kernel_variant(handle, gd, 4) == INV
    assert cd.set_pipeline_chunks(handle, gd, 65) == INV
    assert cd.set_pipeline_chunks(handle, gd, 8) == 0
    cd.cudecompGridDescDestroy(handle, gd)


def test_workspace_sizes_match_the_oracle_for_odd_shapes(handle):
    """cudecompGetHaloWorkspaceSize / cudecompGetTransposeWorkspaceSize against the oracle on one rank for shapes whose
    three extents differ and halos that differ per dimension (a swapped axis in the face product would hide behind the
    maximum over the dimensions or behind the rounding to 64 elements on friendlier shapes)."""
    from oracle import oracle as orc
    import itertools
    for gdims, ac in itertools.product([(9, 130, 17), (200, 3, 70), (65, 66, 5)], [False, True]):
        c = _config(gdims=gdims)
        for i in range(3):
            c.transpose_axis_contiguous[i] = ac
        res, gd = cd.cudecompGridDescCreate(handle, c)
        assert res == 0
        o = orc.Oracle(list(gdims), [1, 1], (ac,) * 3)
        assert cd.cudecompGetTransposeWorkspaceSize(handle, gd) == (0, o.transpose_workspace_size())
        for halo in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (3, 1, 2), (1, 4, 1), (2, 2, 5)]:
            for ax in range(3):
                assert cd.cudecompGetHaloWorkspaceSize(handle, gd, ax, halo) == (0, o.halo_workspace_size(0, ax, list(halo))), (
                    gdims, ac, halo, ax)
        cd.cudecompGridDescDestroy(handle, gd)

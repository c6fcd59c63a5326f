"""The reference's API contract tests (tests/ctest/api_tests.cc) restated for one rank of a 4-rank run.

Every function below follows one TEST_F of the reference (line numbers cited) and returns the list of expectations that
did not hold; the parent test asserts that every rank returns an empty list. Calls go through the raw C symbols so that
null out-pointers, wrong struct sizes and wrong versions can be passed exactly as the reference's tests pass them.
`with_gpu` enables the parts that allocate device memory (cudecompMalloc); everything else is host-only validation.
"""
import ctypes
import os

from cudecomp_b200 import capi as cd

L = cd.lib
INV = cd.CUDECOMP_RESULT_INVALID_USAGE
NOT_SUPPORTED = cd.CUDECOMP_RESULT_NOT_SUPPORTED
OK = cd.CUDECOMP_RESULT_SUCCESS
CFG_V = cd.CUDECOMP_GRID_DESC_CONFIG_VERSION
OPT_V = cd.CUDECOMP_GRID_DESC_AUTOTUNE_OPTIONS_VERSION
PI_V = cd.CUDECOMP_PENCIL_INFO_VERSION
I32_MAX = 2 ** 31 - 1

# api_tests.cc:69-75
GDIMS, GDIMS_DIST, PDIMS = (9, 10, 11), (8, 9, 10), (2, 2)
HALO, PADDING, PERIODS = (1, 2, 1), (1, 0, 2), (False, True, False)


class Expect:
    def __init__(self):
        self.failures = []

    def eq(self, want, got, what):
        if want != got:
            self.failures.append("%s: expected %r, got %r" % (what, want, got))

    def true(self, cond, what):
        if not cond:
            self.failures.append(what)


def distributed_config():  # api_tests.cc:516-521 + test_utils setDistributedConfig
    c = cd.cudecompGridDescConfig_t()
    assert cd.cudecompGridDescConfigSetDefaults(c) == OK
    c.gdims[:] = GDIMS
    c.pdims[:] = PDIMS
    return c


def empty_pencil_config():  # api_tests.cc:523-529
    c = distributed_config()
    c.gdims_dist[:] = [GDIMS[0], 1, GDIMS[2]]
    return c


def fast_autotune_options():  # api_tests.cc:531-540
    o = cd.cudecompGridDescAutotuneOptions_t()
    assert cd.cudecompGridDescAutotuneOptionsSetDefaults(o) == OK
    o.n_warmup_trials, o.n_trials, o.dtype = 0, 1, cd.CUDECOMP_FLOAT
    o.disable_nccl_backends = o.disable_nvshmem_backends = True
    return o


def create(handle, config, options=None):
    return cd.cudecompGridDescCreate(handle, config, options)


def create_raw(handle, gd_ref, cfg_ref, cfg_size, cfg_version, opt_ref=None, opt_size=0, opt_version=0):
    return L.cudecompGridDescCreateVersioned(handle, gd_ref, cfg_ref, cfg_size, cfg_version, opt_ref, opt_size,
                                             opt_version)


def expect_create_invalid(e, handle, config, what):  # api_tests.cc:542-547
    res, gd = create(handle, config)
    e.eq(INV, res, "GridDescCreate(%s)" % what)
    if gd:
        cd.cudecompGridDescDestroy(handle, gd)


class GridDesc:
    """Collective create / destroy around a block, like the reference's gridDescGuard."""

    def __init__(self, handle, config, options=None):
        self.handle, self.config, self.options = handle, config, options

    def __enter__(self):
        res, self.gd = create(self.handle, self.config, self.options)
        if res != OK:
            raise RuntimeError("cudecompGridDescCreate -> %d" % res)
        return self.gd

    def __exit__(self, *exc):
        cd.cudecompGridDescDestroy(self.handle, self.gd)
        return False


# ------------------------------------------------------------------------------------------------ the tests
def init_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:571-573
    e.eq(INV, L.cudecompInit(None, cd.MPI_COMM_WORLD), "cudecompInit(nullptr)")


def finalize_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1031-1033
    e.eq(INV, L.cudecompFinalize(None), "cudecompFinalize(nullptr)")


def multiple_live_handles(e, handle, rank, with_gpu):  # api_tests.cc:575-608
    res, second = cd.cudecompInit(cd.MPI_COMM_WORLD)
    e.eq(OK, res, "second cudecompInit")
    c1 = distributed_config()
    c2 = distributed_config()
    c2.rank_order = cd.CUDECOMP_RANK_ORDER_COL_MAJOR
    with GridDesc(handle, c1) as gd1, GridDesc(second, c2) as gd2:
        q = cd.cudecompGridDescConfig_t()
        e.eq(INV, L.cudecompGetGridDescConfigVersioned(second, gd1, ctypes.byref(q), ctypes.sizeof(q), CFG_V),
             "GetGridDescConfig with a descriptor of another handle")
        e.eq(INV, L.cudecompGridDescDestroy(second, gd1), "GridDescDestroy with a descriptor of another handle")
        if with_gpu:
            r1, b1 = cd.cudecompMalloc(handle, gd1, 1024)
            r2, b2 = cd.cudecompMalloc(second, gd2, 2048)
            e.eq((OK, OK), (r1, r2), "cudecompMalloc on two handles")
            e.true(b1 != 0 and b2 != 0 and b1 != b2, "two handles hand out distinct buffers")
            unused = ctypes.c_void_p()
            e.eq(INV, L.cudecompMalloc(second, gd1, ctypes.byref(unused), 1024), "Malloc with a foreign descriptor")
            e.eq(None, unused.value, "failed Malloc leaves the pointer null")
            e.eq(OK, cd.cudecompFree(handle, gd1, b1), "Free 1")
            e.eq(OK, cd.cudecompFree(second, gd2, b2), "Free 2")
    e.eq(OK, cd.cudecompFinalize(second), "Finalize second handle")


def finalize_in_creation_order(e, handle, rank, with_gpu):  # api_tests.cc:610-632 (on two fresh handles)
    r1, first = cd.cudecompInit(cd.MPI_COMM_WORLD)
    r2, second = cd.cudecompInit(cd.MPI_COMM_WORLD)
    e.eq((OK, OK), (r1, r2), "two cudecompInit")
    with GridDesc(first, distributed_config()), GridDesc(second, distributed_config()):
        pass
    e.eq(OK, cd.cudecompFinalize(first), "Finalize first")
    e.eq(OK, cd.cudecompFinalize(second), "Finalize second")


def multiple_nccl_backed_handles(e, handle, rank, with_gpu):  # api_tests.cc:634-656
    res, second = cd.cudecompInit(cd.MPI_COMM_WORLD)
    e.eq(OK, res, "second cudecompInit")
    c1 = distributed_config()
    c1.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_NCCL
    c2 = distributed_config()
    c2.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_NCCL
    with GridDesc(handle, c1), GridDesc(second, c2):
        pass
    e.eq(OK, cd.cudecompFinalize(second), "Finalize second handle")


def grid_desc_create_rejects_invalid_configs(e, handle, rank, with_gpu):  # api_tests.cc:1035-1088
    def variant(what, mutate):
        c = distributed_config()
        mutate(c)
        expect_create_invalid(e, handle, c, what)

    def mem_order_011(c):
        for ax in range(3):
            c.transpose_mem_order[ax][:] = [0, 1, 1]

    variant("magic=0", lambda c: setattr(c, "magic", 0))
    variant("version+1", lambda c: setattr(c, "version", CFG_V + 1))
    variant("struct_size-1", lambda c: setattr(c, "struct_size", ctypes.sizeof(c) - 1))
    variant("pdims 1x1 on 4 ranks", lambda c: c.pdims.__setitem__(slice(None), [1, 1]))
    variant("pdims 0x1", lambda c: c.pdims.__setitem__(slice(None), [0, 1]))
    variant("pdims -2x-2", lambda c: c.pdims.__setitem__(slice(None), [-2, -2]))
    variant("rank_order 999", lambda c: setattr(c, "rank_order", 999))
    variant("transpose backend 999", lambda c: setattr(c, "transpose_comm_backend", 999))
    variant("halo backend 999", lambda c: setattr(c, "halo_comm_backend", 999))
    variant("partial mem order", lambda c: c.transpose_mem_order[0].__setitem__(0, 0))
    variant("mem order 0,1,1", mem_order_011)
    variant("gdims_dist > gdims", lambda c: c.gdims_dist.__setitem__(slice(None), [GDIMS[0] + 1, GDIMS[1], GDIMS[2]]))


def grid_desc_create_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1090-1104
    c = distributed_config()
    gd = cd.cudecompGridDesc_t()
    size = ctypes.sizeof(c)
    e.eq(INV, create_raw(None, ctypes.byref(gd), ctypes.byref(c), size, CFG_V), "null handle")
    e.eq(INV, create_raw(handle, None, ctypes.byref(c), size, CFG_V), "null grid_desc out-pointer")
    e.eq(INV, create_raw(handle, ctypes.byref(gd), None, size, CFG_V), "null config")
    e.eq(INV, create_raw(handle, ctypes.byref(gd), ctypes.byref(c), size - 1, CFG_V), "config size - 1")
    e.eq(INV, create_raw(handle, ctypes.byref(gd), ctypes.byref(c), size, CFG_V + 1), "config version + 1")


def grid_desc_create_rejects_invalid_autotune_inputs(e, handle, rank, with_gpu):  # api_tests.cc:1106-1136
    c = distributed_config()
    c.pdims[:] = [0, 0]
    e.eq(INV, create(handle, c)[0], "pdims 0x0 without options")

    def variant(what, mutate):
        cfg, o = distributed_config(), fast_autotune_options()
        mutate(o)
        e.eq(INV, create(handle, cfg, o)[0], "autotune options " + what)

    variant("magic=0", lambda o: setattr(o, "magic", 0))
    variant("version+1", lambda o: setattr(o, "version", OPT_V + 1))
    variant("struct_size-1", lambda o: setattr(o, "struct_size", ctypes.sizeof(o) - 1))
    cfg, o = distributed_config(), fast_autotune_options()
    gd = cd.cudecompGridDesc_t()
    e.eq(INV, create_raw(handle, ctypes.byref(gd), ctypes.byref(cfg), ctypes.sizeof(cfg), CFG_V, ctypes.byref(o),
                         ctypes.sizeof(o) - 1, OPT_V), "options size - 1 through the versioned entry point")
    variant("grid_mode 999", lambda o: setattr(o, "grid_mode", 999))


def fixed_selections_ignore_autotune_environment(e, handle, rank, with_gpu):  # api_tests.cc:1138-1149
    names = ["CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS", "CUDECOMP_AUTOTUNE_HALO_BACKENDS", "CUDECOMP_AUTOTUNE_P_ROW_RANGE",
             "CUDECOMP_AUTOTUNE_P_COL_RANGE"]
    old = {n: os.environ.get(n) for n in names}
    for n in names:
        os.environ[n] = "invalid"
    try:
        res, gd = create(handle, distributed_config(), fast_autotune_options())
        e.eq(OK, res, "fixed pdims and backends with invalid autotune environment")
        if res == OK:
            cd.cudecompGridDescDestroy(handle, gd)
    finally:
        for n, v in old.items():
            if v is None:
                os.environ.pop(n, None)
            else:
                os.environ[n] = v


def grid_desc_destroy_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1151-1153
    e.eq(INV, L.cudecompGridDescDestroy(handle, None), "GridDescDestroy(handle, nullptr)")


def config_fields(c):
    return dict(rank_order=c.rank_order, tb=c.transpose_comm_backend, hb=c.halo_comm_backend, pdims=list(c.pdims),
                gdims=list(c.gdims), gdims_dist=list(c.gdims_dist), ac=list(c.transpose_axis_contiguous),
                mo=[list(r) for r in c.transpose_mem_order])


def create_preserves_config_settings(e, handle, rank, with_gpu):  # api_tests.cc:1155-1175
    c = distributed_config()
    c.gdims_dist[:] = GDIMS_DIST
    c.rank_order = cd.CUDECOMP_RANK_ORDER_COL_MAJOR
    c.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_MPI_A2A
    c.halo_comm_backend = cd.CUDECOMP_HALO_COMM_MPI_BLOCKING
    for i in range(3):
        c.transpose_axis_contiguous[i] = True
    expected = config_fields(c)
    with GridDesc(handle, c) as gd:
        res, q = cd.cudecompGetGridDescConfig(handle, gd)
        e.eq(OK, res, "GetGridDescConfig")
        e.eq(expected, config_fields(c), "config after create")
        e.eq(expected, config_fields(q), "queried config")


def get_grid_desc_config_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1177-1194
    with GridDesc(handle, distributed_config()) as gd:
        q = cd.cudecompGridDescConfig_t()
        f = L.cudecompGetGridDescConfigVersioned
        e.eq(INV, f(handle, gd, None, ctypes.sizeof(q), CFG_V), "null config")
        e.eq(INV, f(handle, None, ctypes.byref(q), ctypes.sizeof(q), CFG_V), "null grid_desc")
        e.eq(INV, f(handle, gd, ctypes.byref(q), ctypes.sizeof(q), CFG_V + 1), "version + 1")
        e.eq(INV, f(handle, gd, ctypes.byref(q), ctypes.sizeof(q) - 1, CFG_V), "size - 1")


def describes_empty_pencils(e, handle, rank, with_gpu):  # api_tests.cc:1292-1308
    with GridDesc(handle, empty_pencil_config()) as gd:
        any_empty = False
        for ax in range(3):
            res, p = cd.cudecompGetPencilInfo(handle, gd, ax)
            e.eq(OK, res, "GetPencilInfo axis %d" % ax)
            if 0 in list(p.shape):
                any_empty = True
                e.eq(0, p.size, "size of an empty pencil (axis %d)" % ax)
        # gdims_dist[1] = 1 over 2 rows/columns leaves an empty pencil on some rank for some axis
        flag = cd.MPI_Allreduce_max(1 if any_empty else 0)
        e.eq(1, flag, "some rank sees an empty pencil")


def get_pencil_info_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1310-1338
    with GridDesc(handle, distributed_config()) as gd:
        p = cd.cudecompPencilInfo_t()
        f = L.cudecompGetPencilInfoVersioned
        size = ctypes.sizeof(p)
        e.eq(INV, f(handle, gd, None, size, PI_V, 0, None, None), "null pencil info")
        e.eq(INV, f(handle, gd, ctypes.byref(p), size - 1, PI_V, 0, None, None), "size - 1")
        e.eq(INV, f(handle, gd, ctypes.byref(p), size, PI_V + 1, 0, None, None), "version + 1")
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, -1)[0], "axis -1")
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, 0, (-1, 0, 0))[0], "negative halo extents")
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, 0, None, (0, -1, 0))[0], "negative padding")
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, 0, (I32_MAX, 0, 0))[0], "oversized halo extents")
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, 0, None, (I32_MAX, 0, 0))[0], "oversized padding")


def get_pencil_info_rejects_size_overflow(e, handle, rank, with_gpu):  # api_tests.cc:1340-1352
    c = distributed_config()
    c.gdims[:] = [I32_MAX] * 3
    with GridDesc(handle, c) as gd:
        e.eq(INV, cd.cudecompGetPencilInfo(handle, gd, 0)[0], "pencil of (2^31-1)^3 / 4 elements")


def workspace_size_queries_reject_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1354-1378
    with GridDesc(handle, distributed_config()) as gd:
        n = ctypes.c_int64(0)
        e.eq(INV, L.cudecompGetTransposeWorkspaceSize(handle, gd, None), "transpose: null size")
        e.eq(INV, L.cudecompGetTransposeWorkspaceSize(handle, None, ctypes.byref(n)), "transpose: null grid_desc")
        halo = (ctypes.c_int32 * 3)(*HALO)
        e.eq(INV, L.cudecompGetHaloWorkspaceSize(handle, gd, 0, None, ctypes.byref(n)), "halo: null extents")
        e.eq(INV, L.cudecompGetHaloWorkspaceSize(handle, gd, 0, halo, None), "halo: null size")
        e.eq(INV, L.cudecompGetHaloWorkspaceSize(handle, gd, 3, halo, ctypes.byref(n)), "halo: axis 3")


def expect_shifted(e, handle, gd, rank, axis, dim, disp, periodic, expected):  # api_tests.cc:244-252
    res, got = cd.cudecompGetShiftedRank(handle, gd, axis, dim, disp, periodic)
    e.eq((OK, expected[rank]), (res, got), "GetShiftedRank(axis=%d dim=%d disp=%d periodic=%s)" % (axis, dim, disp,
                                                                                                 periodic))


def shifted_ranks_row_major(e, handle, rank, with_gpu):  # api_tests.cc:1380-1393
    with GridDesc(handle, distributed_config()) as gd:
        expect_shifted(e, handle, gd, rank, 0, 1, 1, False, [2, 3, -1, -1])
        expect_shifted(e, handle, gd, rank, 0, 1, -1, False, [-1, -1, 0, 1])
        expect_shifted(e, handle, gd, rank, 0, 1, 1, True, [2, 3, 0, 1])
        expect_shifted(e, handle, gd, rank, 0, 2, 1, False, [1, -1, 3, -1])
        expect_shifted(e, handle, gd, rank, 0, 2, -1, False, [-1, 0, -1, 2])
        expect_shifted(e, handle, gd, rank, 0, 2, 1, True, [1, 0, 3, 2])


def shifted_ranks_column_major(e, handle, rank, with_gpu):  # api_tests.cc:1395-1409
    c = distributed_config()
    c.rank_order = cd.CUDECOMP_RANK_ORDER_COL_MAJOR
    with GridDesc(handle, c) as gd:
        expect_shifted(e, handle, gd, rank, 0, 1, 1, False, [1, -1, 3, -1])
        expect_shifted(e, handle, gd, rank, 0, 1, -1, False, [-1, 0, -1, 2])
        expect_shifted(e, handle, gd, rank, 0, 1, 1, True, [1, 0, 3, 2])
        expect_shifted(e, handle, gd, rank, 0, 2, 1, False, [2, 3, -1, -1])
        expect_shifted(e, handle, gd, rank, 0, 2, -1, False, [-1, -1, 0, 1])
        expect_shifted(e, handle, gd, rank, 0, 2, 1, True, [2, 3, 0, 1])


def shifted_ranks_axis_aligned_and_zero(e, handle, rank, with_gpu):  # api_tests.cc:1411-1433
    with GridDesc(handle, distributed_config()) as gd:
        e.eq((OK, rank), cd.cudecompGetShiftedRank(handle, gd, 0, 1, 0, False), "zero displacement")
        e.eq((OK, -1), cd.cudecompGetShiftedRank(handle, gd, 0, 0, 1, False), "pencil axis, not periodic")
        e.eq((OK, rank), cd.cudecompGetShiftedRank(handle, gd, 0, 0, 1, True), "pencil axis, periodic")
        e.eq((OK, rank), cd.cudecompGetShiftedRank(handle, gd, 0, 1, PDIMS[0], True), "full wrap, periodic")
        e.eq((OK, -1), cd.cudecompGetShiftedRank(handle, gd, 0, 1, PDIMS[0], False), "full wrap, not periodic")


def shifted_rank_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1435-1445
    with GridDesc(handle, distributed_config()) as gd:
        r = ctypes.c_int32(0)
        e.eq(INV, L.cudecompGetShiftedRank(handle, gd, 0, 1, 1, False, None), "null out-pointer")
        e.eq(INV, L.cudecompGetShiftedRank(handle, gd, 3, 1, 1, False, ctypes.byref(r)), "axis 3")
        e.eq(INV, L.cudecompGetShiftedRank(handle, gd, 0, 3, 1, False, ctypes.byref(r)), "dim 3")


def malloc_and_free_reject_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1447-1466
    with GridDesc(handle, distributed_config()) as gd:
        buf = ctypes.c_void_p()
        e.eq(INV, L.cudecompMalloc(handle, gd, None, 16), "Malloc: null out-pointer")
        e.eq(INV, L.cudecompMalloc(handle, gd, ctypes.byref(buf), 0), "Malloc: zero bytes")
        e.eq(INV, L.cudecompFree(None, gd, None), "Free: null handle")
        e.eq(INV, L.cudecompFree(handle, None, None), "Free: null grid_desc")


def transpose_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1468-1491
    with GridDesc(handle, distributed_config()) as gd:
        word = ctypes.c_int(0)
        ptr = ctypes.c_void_p(ctypes.addressof(word))
        f = L.cudecompTransposeXToY
        tail = (None, None, None, None, None)
        e.eq(INV, f(handle, gd, None, ptr, ptr, cd.CUDECOMP_FLOAT, *tail), "null input")
        e.eq(INV, f(handle, gd, ptr, None, ptr, cd.CUDECOMP_FLOAT, *tail), "null output")
        e.eq(INV, f(handle, gd, ptr, ptr, None, cd.CUDECOMP_FLOAT, *tail), "null work")
        e.eq(INV, f(handle, gd, ptr, ptr, ptr, 999, *tail), "dtype 999")


def transpose_rejects_empty_pencils(e, handle, rank, with_gpu):  # api_tests.cc:1493-1505
    with GridDesc(handle, empty_pencil_config()) as gd:
        words = (ctypes.c_int * 3)()
        a = ctypes.addressof(words)
        e.eq(NOT_SUPPORTED, L.cudecompTransposeXToY(handle, gd, ctypes.c_void_p(a), ctypes.c_void_p(a + 4),
                                                    ctypes.c_void_p(a + 8), cd.CUDECOMP_FLOAT, None, None, None, None,
                                                    None), "transpose on a decomposition with empty pencils")


def halo_rejects_invalid_arguments(e, handle, rank, with_gpu):  # api_tests.cc:1507-1533
    with GridDesc(handle, distributed_config()) as gd:
        word = ctypes.c_int(0)
        ptr = ctypes.c_void_p(ctypes.addressof(word))
        halo = (ctypes.c_int32 * 3)(*HALO)
        periods = (ctypes.c_bool * 3)(*PERIODS)
        f = L.cudecompUpdateHalosX
        e.eq(INV, f(handle, gd, ptr, ptr, cd.CUDECOMP_FLOAT, None, periods, 0, None, None), "null halo extents")
        e.eq(INV, f(handle, gd, None, ptr, cd.CUDECOMP_FLOAT, halo, periods, 0, None, None), "null input")
        e.eq(INV, f(handle, gd, ptr, None, cd.CUDECOMP_FLOAT, halo, periods, 0, None, None), "null work")
        e.eq(INV, f(handle, gd, ptr, ptr, cd.CUDECOMP_FLOAT, halo, periods, 3, None, None), "dim 3")
        e.eq(INV, f(handle, gd, ptr, ptr, 999, halo, periods, 0, None, None), "dtype 999")


def halo_rejects_empty_pencils(e, handle, rank, with_gpu):  # api_tests.cc:1535-1547
    with GridDesc(handle, empty_pencil_config()) as gd:
        words = (ctypes.c_int * 2)()
        a = ctypes.addressof(words)
        halo = (ctypes.c_int32 * 3)(*HALO)
        periods = (ctypes.c_bool * 3)(*PERIODS)
        e.eq(NOT_SUPPORTED, L.cudecompUpdateHalosX(handle, gd, ctypes.c_void_p(a), ctypes.c_void_p(a + 4),
                                                   cd.CUDECOMP_FLOAT, halo, periods, 0, None, None),
             "halo update on a decomposition with empty pencils")


def autotune_grid_rules_without_timing(e, handle, rank, with_gpu):
    """Which process grids the autotuner considers at all (reference src/autotune.cc:94-106,361-373), observable without
    a device: with nothing to time, the first candidate that survives the rules is selected. 4 ranks, row-major:
    candidates in order 4x1, 2x2, 1x4; a grid is dropped when it leaves a pencil empty (pdims[0] > min(gx, gy) or
    pdims[1] > min(gy, gz)), when it splits unevenly and uneven decompositions are not allowed, or when the
    CUDECOMP_AUTOTUNE_P_*_RANGE variables exclude it; nothing left -> NOT_SUPPORTED."""
    import torch
    if with_gpu or torch.cuda.is_available():
        return  # with a device the candidates are timed and the fastest one wins

    def selected(gdims, allow_uneven=True, env=None):
        c = distributed_config()
        c.gdims[:] = gdims
        c.pdims[:] = [0, 0]
        o = fast_autotune_options()
        o.allow_uneven_decompositions = allow_uneven
        old = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update(env or {})
        try:
            res, gd = create(handle, c, o)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        if res != OK:
            return res
        cd.cudecompGridDescDestroy(handle, gd)
        return list(c.pdims)  # written back to the caller's config (reference src/cudecomp.cc:1247-1265)

    e.eq([4, 1], selected((8, 6, 10)), "first candidate, everything allowed")
    e.eq([2, 2], selected((8, 6, 10), allow_uneven=False), "4x1 splits y = 6 unevenly")
    e.eq([4, 1], selected((8, 8, 10), allow_uneven=False), "4x1 splits 8 x 8 evenly")
    e.eq([2, 2], selected((2, 8, 8)), "4x1 would leave X pencils empty (4 > min(2, 8))")
    e.eq([1, 4], selected((1, 8, 8)), "only 1x4 keeps every pencil populated")
    e.eq(NOT_SUPPORTED, selected((1, 1, 8)), "no grid keeps every pencil populated")
    e.eq(NOT_SUPPORTED, selected((6, 6, 7), allow_uneven=False), "no even split of z = 7 or y = 6 by 4 / of 7 by 2")
    e.eq([2, 2], selected((8, 6, 10), env={"CUDECOMP_AUTOTUNE_P_ROW_RANGE": "1,2"}), "row range excludes 4x1")
    e.eq([1, 4], selected((8, 6, 10), env={"CUDECOMP_AUTOTUNE_P_COL_RANGE": "3,4"}), "column range leaves 1x4")
    e.eq(INV, selected((8, 6, 10), env={"CUDECOMP_AUTOTUNE_P_ROW_RANGE": "3,3"}), "range excludes every candidate")


TESTS = [
    autotune_grid_rules_without_timing,
    init_rejects_invalid_arguments,
    finalize_rejects_invalid_arguments,
    multiple_live_handles,
    finalize_in_creation_order,
    multiple_nccl_backed_handles,
    grid_desc_create_rejects_invalid_configs,
    grid_desc_create_rejects_invalid_arguments,
    grid_desc_create_rejects_invalid_autotune_inputs,
    fixed_selections_ignore_autotune_environment,
    grid_desc_destroy_rejects_invalid_arguments,
    create_preserves_config_settings,
    get_grid_desc_config_rejects_invalid_arguments,
    describes_empty_pencils,
    get_pencil_info_rejects_invalid_arguments,
    get_pencil_info_rejects_size_overflow,
    workspace_size_queries_reject_invalid_arguments,
    shifted_ranks_row_major,
    shifted_ranks_column_major,
    shifted_ranks_axis_aligned_and_zero,
    shifted_rank_rejects_invalid_arguments,
    malloc_and_free_reject_invalid_arguments,
    transpose_rejects_invalid_arguments,
    transpose_rejects_empty_pencils,
    halo_rejects_invalid_arguments,
    halo_rejects_empty_pencils,
]
TEST_NAMES = [t.__name__ for t in TESTS]


def run(handle, rank, name, with_gpu):
    e = Expect()
    dict((t.__name__, t) for t in TESTS)[name](e, handle, rank, with_gpu)
    return dict(ok=not e.failures, failures=e.failures)

"""Case matrices, restated from the reference's parametrised tests.

transpose: tests/ctest/transpose_tests.cc:163-273 (baseline sweep + coverage cases, gdims 9x10x11)
halo     : tests/ctest/halo_tests.cc:103-146 (halo (1,3,2), gdims 9x10x11)
legacy   : tests/test_config.yaml + tests/test_runner.py:80-90 (all 36 (x,y) memory-order pairs, chained X->Y->Z->Y->X)
"""
import itertools

GDIMS = [9, 10, 11]
IN_HALO, OUT_HALO = [1, 2, 1], [2, 1, 1]
IN_PAD, OUT_PAD = [1, 1, 2], [2, 1, 1]
OPS = ["XY", "YX", "YZ", "ZY"]
AXES = {"XY": (0, 1), "YX": (1, 0), "YZ": (1, 2), "ZY": (2, 1)}

UNPACK_ORDER = [[0, 1, 2], [0, 1, 2], [0, 1, 2]]
TRANSPOSE_UNPACK_ORDER = [[0, 2, 1], [0, 1, 2], [0, 1, 2]]
SPLIT_UNPACK_ORDER = [[0, 1, 2], [0, 2, 1], [1, 2, 0]]


def with_halo_padding(case):
    a, b = AXES[case["op"]]
    case = dict(case)
    case["halos"] = {str(a): IN_HALO, str(b): OUT_HALO}
    case["pads"] = {str(a): IN_PAD, str(b): OUT_PAD}
    return case


def transpose_baseline(pdims, dtypes=("float", "float_complex")):
    """appendBaselineCases: layouts x dtypes x ops x in/out of place."""
    cases = []
    for layout, ac in (("DefaultLayout", [False] * 3), ("AxisContiguous", [True] * 3)):
        for dtype in dtypes:
            for op in OPS:
                for oop in (False, True):
                    cases.append(dict(kind="transpose", name="Baseline%s_%s_%s_P%dx%d_%s" % (
                        layout, op, dtype, pdims[0], pdims[1], "OutOfPlace" if oop else "InPlace"),
                        gdims=GDIMS, pdims=list(pdims), dtype=dtype, op=op, out_of_place=oop, axis_contiguous=ac))
    return cases


def transpose_coverage_2x2():
    """appendCoverageCases, the 2x2 ones."""
    base = dict(kind="transpose", gdims=GDIMS, pdims=[2, 2], dtype="float", out_of_place=True)
    cases = [
        with_halo_padding(dict(base, name="ExplicitMemOrderUnpack_XY", op="XY", mem_order=UNPACK_ORDER)),
        with_halo_padding(dict(base, name="ExplicitMemOrderTransposeUnpack_XY", op="XY",
                               mem_order=TRANSPOSE_UNPACK_ORDER)),
    ]
    for op in OPS:
        cases.append(with_halo_padding(dict(base, name="ExplicitMemOrderSplitUnpack_%s" % op, op=op,
                                            mem_order=SPLIT_UNPACK_ORDER)))
    cases.append(dict(base, name="ColumnMajorRankOrder_XY", op="XY", out_of_place=False, rank_order=2))
    cases.append(with_halo_padding(dict(base, name="DtypeWorkspacePadding_XY_double", op="XY", dtype="double",
                                        mem_order=SPLIT_UNPACK_ORDER)))
    cases.append(with_halo_padding(dict(base, name="DtypeWorkspacePadding_YZ_double_complex", op="YZ",
                                        dtype="double_complex", mem_order=SPLIT_UNPACK_ORDER)))
    # pipelined-backend coverage cases of the reference, same layouts through this engine
    cases.append(with_halo_padding(dict(base, name="ExplicitMemOrderTransposePackOffset_XY", op="XY",
                                        mem_order=[[1, 0, 2], [1, 2, 0], [0, 1, 2]])))
    # staged schedule (backend value NVSHMEM) and a non-cudecompMalloc workspace
    cases.append(dict(base, name="StagedBackend_XY", op="XY", backend=6))
    cases.append(dict(base, name="ForceStaged_YZ_AxisContiguous", op="YZ", force_staged=True,
                      axis_contiguous=[True] * 3))
    cases.append(dict(base, name="TorchWorkspace_ZY", op="ZY", work_alloc="torch", out_of_place=False))
    # gdims_dist: the grid is distributed as 8x9x10, the remainder rides on the last rank (api_tests.cc:132-153)
    for op in OPS:
        cases.append(dict(base, name="GdimsDist_%s" % op, op=op, gdims_dist=[8, 9, 10]))
    return cases


def transpose_coverage_3x1():
    base = dict(kind="transpose", gdims=GDIMS, pdims=[3, 1], dtype="float", out_of_place=True)
    cases = [with_halo_padding(dict(base, name="NonPowerOfTwoCommunicator_XY", op="XY", mem_order=UNPACK_ORDER))]
    for mo, nm in ((UNPACK_ORDER, "Unpack"), (TRANSPOSE_UNPACK_ORDER, "TransposeUnpack"),
                   (SPLIT_UNPACK_ORDER, "SplitUnpack")):
        cases.append(with_halo_padding(dict(base, name="InterGroup%s_XY" % nm, op="XY", mem_order=mo)))
    for op in OPS:
        cases.append(dict(base, name="P3x1_%s_inplace" % op, op=op, out_of_place=False, axis_contiguous=[True] * 3))
    return cases


def transpose_single_rank():
    cases = transpose_baseline((1, 1))
    base = dict(kind="transpose", gdims=GDIMS, pdims=[1, 1], dtype="float", out_of_place=True)
    cases.append(with_halo_padding(dict(base, name="DirectTransposePackOffset_XY", op="XY",
                                        mem_order=[[1, 0, 2], [2, 1, 0], [0, 1, 2]])))
    cases.append(with_halo_padding(dict(base, name="DirectTransposeUnpackOffset_XY", op="XY",
                                        mem_order=[[1, 0, 2], [0, 1, 2], [0, 1, 2]])))
    for dtype in ("double", "double_complex"):
        cases.append(dict(base, name="Chain_%s" % dtype, dtype=dtype, ops=["XY", "YZ", "ZY", "YX"],
                          axis_contiguous=[True] * 3))
    return cases


def legacy_mem_order_chain(pdims, dtype="float", out_of_place=False, stride=1):
    """All 36 (x,y) memory-order pairs of tests/test_runner.py:80-90 with z cycling, chained X->Y->Z->Y->X."""
    perms = [list(p) for p in itertools.permutations(range(3))]
    cases = []
    k = 0
    for i, px in enumerate(perms):
        for j, py in enumerate(perms):
            k += 1
            if (k - 1) % stride:
                continue
            pz = perms[(i + j) % 6]
            cases.append(dict(kind="transpose", name="MemOrder_x%s_y%s_z%s" % ("".join(map(str, px)), "".join(map(str, py)),
                                                                               "".join(map(str, pz))),
                              gdims=[12, 10, 14], pdims=list(pdims), dtype=dtype, out_of_place=out_of_place,
                              ops=["XY", "YZ", "ZY", "YX"], mem_order=[px, py, pz], fills=["pattern"]))
    return cases


HALO = [1, 3, 2]


def halo_baseline(pdims=(2, 2)):
    cases = []
    for layout, ac in (("DefaultLayout", [False] * 3), ("AxisContiguous", [True] * 3)):
        for axis in range(3):
            for dtype in ("float", "float_complex"):
                for pname, per in (("Periodic", [True] * 3), ("NonPeriodic", [False] * 3)):
                    cases.append(dict(kind="halo", name="Baseline%s%s_Axis%d_%s_P%dx%d" % (
                        layout, pname, axis, dtype, pdims[0], pdims[1]), gdims=GDIMS, pdims=list(pdims), dtype=dtype,
                        axis=axis, halo=HALO, periods=per, axis_contiguous=ac))
    return cases


def halo_coverage():
    base = dict(kind="halo", gdims=GDIMS, pdims=[2, 2], dtype="float", axis=0, halo=HALO, periods=[True] * 3)
    return [
        dict(base, name="NonzeroPadding", padding=[1, 0, 2]),
        dict(base, name="ColumnMajorRankOrder", rank_order=2),
        dict(base, name="DtypeWorkspacePadding_double", dtype="double", padding=[1, 0, 2]),
        dict(base, name="DtypeWorkspacePadding_double_complex", dtype="double_complex", padding=[1, 0, 2]),
        dict(base, name="StagedBackend", halo_backend=4),
        dict(base, name="ForceStagedAxis2", axis=2, force_staged=True, padding=[0, 1, 1]),
        dict(base, name="MixedPeriods", periods=[True, False, True], axis=1),
    ]


def halo_3x1():
    base = dict(kind="halo", gdims=GDIMS, pdims=[3, 1], dtype="float", axis=0, halo=HALO, periods=[False] * 3)
    return [dict(base, name="InteriorNonPeriodicNeighbors"),
            dict(base, name="InteriorPeriodicNeighbors_Axis1", axis=1, periods=[True] * 3, halo=[1, 1, 2])]

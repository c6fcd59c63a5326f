"""Pins the CPU oracle (oracle/cudecomp_oracle.c) to the reference's own known answers. CPU only.

  * golden pencil-info tables of tests/ctest/api_tests.cc:92-153 (3 decompositions x 3 axes x 4 ranks)
  * golden shifted-rank tables of api_tests.cc:1386-1408
  * the analytic global-index pattern the reference uses to check every transpose (transpose_tests.cc:323-378)
    and halo update (halo_tests.cc:229-272), over the reference's own case matrix.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import cases as C

DT = orc.NP_DTYPES


@pytest.mark.parametrize("table,kwargs", [("default", {}), ("column_major", {"col_major": True}),
                                          ("gdims_dist", {"use_dist": True})])
def test_pencil_info_matches_reference_tables(golden, table, kwargs):
    dist = golden["gdims_dist"] if kwargs.get("use_dist") else None
    o = orc.Oracle(golden["gdims"], golden["pdims"], gdims_dist=dist, col_major=kwargs.get("col_major", False))
    for axis in range(3):
        for rank in range(4):
            want = golden["pencil_info"][table][axis][rank]
            got = o.pencil_info(rank, axis, golden["halo_extents"], golden["padding"])
            assert list(got.shape) == want["shape"], (table, axis, rank)
            assert list(got.lo) == want["lo"]
            assert list(got.hi) == want["hi"]
            assert list(got.order) == want["order"]
            assert list(got.halo_extents) == want["halo_extents"]
            assert list(got.padding) == want["padding"]
            assert got.size == want["size"]


@pytest.mark.parametrize("layout", ["row_major", "col_major"])
def test_shifted_ranks_match_reference_tables(golden, layout):
    o = orc.Oracle(golden["gdims"], golden["pdims"], col_major=(layout == "col_major"))
    for q in golden["shifted_ranks"][layout]:
        got = [o.shifted_rank(r, q["axis"], q["dim"], q["displacement"], q["periodic"]) for r in range(4)]
        assert got == q["expected"], q


def test_shifted_rank_axis_aligned_and_zero(golden):
    # api_tests.cc:1410-1433
    o = orc.Oracle(golden["gdims"], golden["pdims"])
    for r in range(4):
        assert o.shifted_rank(r, 0, 1, 0, False) == r
        assert o.shifted_rank(r, 0, 0, 1, False) == -1
        assert o.shifted_rank(r, 0, 0, 1, True) == r
        assert o.shifted_rank(r, 0, 1, golden["pdims"][0], True) == r
        assert o.shifted_rank(r, 0, 1, golden["pdims"][0], False) == -1


def _run_transpose_case(case):
    o = orc.Oracle(case["gdims"], case["pdims"], case.get("axis_contiguous") or (False,) * 3, case.get("mem_order"),
                   case.get("gdims_dist"), case.get("rank_order", 0) == 2)
    dt = DT[case.get("dtype", "float")]
    halos, pads = case.get("halos") or {}, case.get("pads") or {}
    ops = case.get("ops") or [case["op"]]
    inplace = not case.get("out_of_place", False)
    a0 = orc.transpose_axes(ops[0])[0]
    sizes = [max(o.pencil_info(r, ax, halos.get(str(ax)), pads.get(str(ax))).size for ax in range(3))
             for r in range(o.nranks)]
    cur = []
    for r in range(o.nranks):
        pa = o.pencil_info(r, a0, halos.get(str(a0)), pads.get(str(a0)))
        buf = np.zeros(sizes[r], dt)
        buf[:pa.size] = orc.pattern_pencil(pa, case["gdims"], dt)
        cur.append(buf)
    other = [np.zeros(n, dt) for n in sizes]
    for op in ops:
        a, b = orc.transpose_axes(op)
        o.transpose(op, cur, cur if inplace else other, halos.get(str(a)), halos.get(str(b)), pads.get(str(a)),
                    pads.get(str(b)))
        res = cur if inplace else other
        for r in range(o.nranks):
            pb = o.pencil_info(r, b, halos.get(str(b)), pads.get(str(b)))
            want = orc.pattern_pencil(pb, case["gdims"], dt)
            assert orc.interior_equal(pb, want, res[r][:pb.size]), (case["name"], op, r)
        if not inplace:
            cur, other = other, cur


ALL_TRANSPOSE = (C.transpose_single_rank() + C.transpose_baseline((2, 2)) + C.transpose_coverage_2x2() +
                 C.transpose_coverage_3x1() + C.legacy_mem_order_chain((2, 2)) +
                 C.legacy_mem_order_chain((2, 2), dtype="double_complex", out_of_place=True, stride=5))


@pytest.mark.parametrize("case", ALL_TRANSPOSE, ids=[c["name"] for c in ALL_TRANSPOSE])
def test_oracle_transpose_equals_analytic_pattern(case):
    _run_transpose_case(case)


ALL_HALO = C.halo_baseline() + C.halo_coverage() + C.halo_3x1() + C.halo_baseline((1, 1))


@pytest.mark.parametrize("case", ALL_HALO, ids=[c["name"] for c in ALL_HALO])
def test_oracle_halo_equals_analytic_reference(case):
    o = orc.Oracle(case["gdims"], case["pdims"], case.get("axis_contiguous") or (False,) * 3, case.get("mem_order"),
                   None, case.get("rank_order", 0) == 2)
    dt = DT[case.get("dtype", "float")]
    ax, halo, per, pad = case["axis"], case["halo"], case["periods"], case.get("padding")
    data, want = [], []
    for r in range(o.nranks):
        p = o.pencil_info(r, ax, halo, pad)
        data.append(orc.pattern_pencil(p, case["gdims"], dt))
        want.append(orc.halo_reference(p, case["gdims"], dt, per))
    for dim in range(3):
        o.halo(ax, dim, data, halo, per, pad)
    for r in range(o.nranks):
        assert np.array_equal(data[r], want[r]), (case["name"], r)


def test_oracle_rejects_wide_halo_and_empty_pencils():
    o = orc.Oracle([9, 10, 11], [2, 2])
    p = o.pencil_info(0, 0, [0, 6, 0])
    with pytest.raises(RuntimeError):
        o.halo(0, 1, [np.zeros(o.pencil_info(r, 0, [0, 6, 0]).size, np.float32) for r in range(4)], [0, 6, 0])
    assert p.size > 0
    o2 = orc.Oracle([2, 2, 2], [4, 1])
    assert o2.has_empty_pencils(0) and o2.has_empty_pencils(1)


def test_oracle_workspace_sizes():
    # reference formulas src/cudecomp.cc:1411-1459 on the golden grid
    o = orc.Oracle([9, 10, 11], [2, 2])
    x, y, z = 9 * 5 * 6, 5 * 10 * 6, 5 * 5 * 11
    al = lambda n: (n + 63) // 64 * 64  # noqa: E731
    assert o.transpose_workspace_size() == max(al(x) + y, al(y) + x, al(y) + z, al(z) + y)
    p = o.pencil_info(0, 0, [1, 2, 1])
    s = {p.order[i]: p.shape[i] for i in range(3)}
    assert o.halo_workspace_size(0, 0, [1, 2, 1]) == max(4 * al(s[1] * s[2] * 1), 4 * al(s[0] * s[2] * 2),
                                                          4 * al(s[0] * s[1] * 1))

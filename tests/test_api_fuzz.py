"""Hostile inputs through the C ABI, host only: configurations, halos, paddings and indices that no caller should
pass -- negative and zero extents, extents near INT32_MAX, process grids that do not match, memory orders that are not
permutations, enumerators out of range, ranks outside the grid -- must come back as a cudecompResult_t (the reference's
contract: nothing throws across the ABI, src/cudecomp.cc:431-443), never as a crash, a hang or a plan that points
outside the pencils it belongs to.

Memory-safety invariant of every plan that IS returned: each box lies inside the source pencil of the sending rank and
inside the destination pencil (or workspace) of the receiving rank, as sized by cudecompGetPencilInfo /
cudecompGetTransposeWorkspaceSize for the same arguments. The kernels trust these numbers blindly."""
import ctypes
import os

import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from cudecomp_b200 import capi as cd

EXAMPLES = int(os.environ.get("CDB_HYPOTHESIS_EXAMPLES", "150"))
VALID_RESULTS = set(range(0, 10))

extent = st.one_of(st.integers(-3, 40), st.sampled_from([0, 1, 2, 2**15, 2**20, 2**31 - 1]))
small = st.integers(-2, 6)
halo3 = st.lists(st.one_of(st.integers(-1, 3), st.sampled_from([0, 0, 2**30])), min_size=3, max_size=3)


@st.composite
def configs(draw, hostile=True):
    c = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(c))
    c.gdims[:] = [draw(extent if hostile else st.integers(1, 40)) for _ in range(3)]
    c.pdims[:] = [draw(st.integers(-1, 5) if hostile else st.integers(1, 4)) for _ in range(2)]
    if draw(st.booleans()):
        c.gdims_dist[:] = [draw(extent if hostile else st.integers(1, 40)) for _ in range(3)]
    c.rank_order = draw(st.integers(-1, 4) if hostile else st.sampled_from([0, 1, 2]))
    c.transpose_comm_backend = draw(st.integers(-1, 10) if hostile else st.integers(1, 8))
    c.halo_comm_backend = draw(st.integers(-1, 7) if hostile else st.integers(1, 5))
    for i in range(3):
        c.transpose_axis_contiguous[i] = draw(st.booleans())
    mode = draw(st.sampled_from(["unset", "perm", "junk"] if hostile else ["unset", "perm"]))
    if mode != "unset":
        perms = [(0, 1, 2), (1, 2, 0), (2, 0, 1), (1, 0, 2), (2, 1, 0), (0, 2, 1)]
        for i in range(3):
            row = draw(st.sampled_from(perms)) if mode == "perm" else [draw(st.integers(-2, 3)) for _ in range(3)]
            c.transpose_mem_order[i][:] = list(row)
    return c


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(configs(), st.integers(-1, 20), st.integers(-1, 3), st.sampled_from([-1, 0, 1, 2]), halo3, halo3, halo3, halo3,
       st.integers(0, 2))
def test_handle_free_planner_survives_hostile_inputs(cfg, rank, ax, direction, ha, hb, pa, pb, staged):
    for fn in (lambda: cd.plan_transpose_boxes(cfg, rank, ax, direction, ha, hb, pa, pb, staged),
               lambda: cd.plan_halo_boxes(cfg, rank, ax, direction, ha, [True, False, True], pa, bool(staged)),
               lambda: cd.plan_pipelined_transpose_boxes(cfg, rank, ax, direction, ha, hb, pa, pb, staged & 1, 3)):
        try:
            boxes = fn()
        except cd.CudecompError as e:
            assert e.code in VALID_RESULTS and e.code != 0
            continue
        for b in boxes:
            assert all(x >= 0 for x in b["extent"]) and b["src_offset"] >= 0 and b["dst_offset"] >= 0, b


@pytest.fixture(scope="module")
def handle():
    env = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE") if k in os.environ}
    assert cd.MPI_Init() == 0
    res, h = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res, "cudecompInit")
    yield h
    cd.check(cd.cudecompFinalize(h))
    os.environ.update(env)


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(configs(), st.integers(-1, 3), halo3, halo3, st.integers(-1, 3), st.integers(-3, 3), st.booleans())
def test_descriptor_api_survives_hostile_inputs(handle, cfg, ax, halo, pad, dim, disp, periodic):
    """One rank, no device: whatever the configuration, creation answers with a result code; on a descriptor that was
    created every query answers with a result code and sizes are never negative."""
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    assert res in VALID_RESULTS
    if res != 0:
        return
    try:
        r, p = cd.cudecompGetPencilInfo(handle, gd, ax, halo, pad)
        assert r in VALID_RESULTS
        if r == 0:
            assert p.size >= 0 and all(s >= 0 for s in p.shape)
            prod = 1
            for s in p.shape:
                prod *= s
            assert prod == p.size
        r, w = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
        assert r in VALID_RESULTS and (r != 0 or w >= 0)
        r, w = cd.cudecompGetHaloWorkspaceSize(handle, gd, ax, halo)
        assert r in VALID_RESULTS and (r != 0 or w >= 0)
        r, nb = cd.cudecompGetShiftedRank(handle, gd, ax, dim, disp, periodic)
        assert r in VALID_RESULTS and (r != 0 or nb in (-1, 0))
        # data-moving entry points: argument errors first, then "no device" -- never a crash
        buf = ctypes.c_void_p(0x1000)
        for op in ("XY", "YZ", "ZY", "YX"):
            r = cd.TRANSPOSES[op](handle, gd, buf, buf, buf, cd.CUDECOMP_DOUBLE, halo, halo, pad, pad)
            assert r in VALID_RESULTS
        r = cd.cudecompUpdateHalosX(handle, gd, buf, buf, cd.CUDECOMP_FLOAT, halo, [periodic] * 3, dim, pad)
        assert r in VALID_RESULTS
    finally:
        assert cd.cudecompGridDescDestroy(handle, gd) == 0


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(configs(hostile=False), st.sampled_from(["XY", "YZ", "ZY", "YX"]),
       st.lists(st.integers(0, 2), min_size=3, max_size=3), st.lists(st.integers(0, 2), min_size=3, max_size=3),
       st.lists(st.integers(0, 2), min_size=3, max_size=3), st.lists(st.integers(0, 2), min_size=3, max_size=3),
       st.booleans())
def test_every_planned_box_lies_inside_its_pencils(cfg, op, ha, hb, pa, pb, staged):
    """Valid configurations of any process grid: every box any rank would execute stays inside the sender's source
    pencil and the receiver's destination pencil (direct) or workspace (staged), sized by the oracle-independent public
    queries of a 1-rank descriptor with the same geometry arguments -- through the handle-free pencil sizes the plan API
    itself reports as the largest offset reached."""
    from oracle import oracle as orc  # sizes only; the planner is the thing under test
    n = cfg.pdims[0] * cfg.pdims[1]
    ac = [bool(cfg.transpose_axis_contiguous[i]) for i in range(3)]
    mo = [list(cfg.transpose_mem_order[i]) for i in range(3)]
    mo = mo if all(v >= 0 for row in mo for v in row) else None
    dist = list(cfg.gdims_dist) if any(cfg.gdims_dist) else None
    if dist and any(d < 1 or d > g for d, g in zip(dist, cfg.gdims)):
        return
    o = orc.Oracle(list(cfg.gdims), list(cfg.pdims), ac, mo, dist, cfg.rank_order == 2)
    ax, direction = orc.TRANSPOSE_OPS[op]
    a, b = orc.transpose_axes(op)
    if o.has_empty_pencils(a) or o.has_empty_pencils(b):
        return
    work = o.transpose_workspace_size()

    def last(offset, extent, stride):
        return offset + sum((e - 1) * s for e, s in zip(extent, stride))

    for r in range(n):
        src_size = o.pencil_info(r, a, ha, pa).size
        for bx in cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, staged):
            if 0 in bx["extent"]:
                continue
            peer = bx["peer_rank"]
            assert 0 <= peer < n
            if bx["is_unpack"]:
                assert peer == r and last(bx["src_offset"], bx["extent"], bx["src_stride"]) < work
                assert last(bx["dst_offset"], bx["extent"], bx["dst_stride"]) < o.pencil_info(r, b, hb, pb).size
            else:
                assert last(bx["src_offset"], bx["extent"], bx["src_stride"]) < src_size
                dst_size = work if staged else o.pencil_info(peer, b, hb, pb).size
                assert last(bx["dst_offset"], bx["extent"], bx["dst_stride"]) < dst_size


wild = st.one_of(st.integers(-3, 70), st.sampled_from([-2**31, 2**31 - 1, 1 << 20, 4096, 65536]))


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.lists(wild, min_size=10, max_size=10), st.integers(-2, 40), st.integers(-1, 4))
def test_schedule_knobs_and_candidate_queries_survive_hostile_values(handle, v, nranks, rank_order):
    """The extension entry points that take raw numbers (include/cudecomp_b200_ext.h): any value is either accepted
    or refused with a result code, and a descriptor that accepted them still plans and destroys."""
    cfg = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(cfg))
    cfg.gdims[:] = [12, 10, 8]
    cfg.pdims[:] = [1, 1]
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    try:
        for r in (cd.set_tuning(handle, gd, v[0], bool(v[1] & 1)), cd.set_kernel_variant(handle, gd, v[2]),
                  cd.set_schedule(handle, gd, v[3], v[4], bool(v[5] & 1)), cd.set_transfer_mode(handle, gd, v[6]),
                  cd.set_pipeline_chunks(handle, gd, v[7]), cd.set_staged_mode(handle, gd, v[8], v[9]),
                  cd.check_errors(handle, gd)):
            assert r in VALID_RESULTS
        assert len(cd.describe_transpose_boxes(handle, gd, 0, 1)) >= 1
    finally:
        assert cd.cudecompGridDescDestroy(handle, gd) == 0
    opts = cd.cudecompGridDescAutotuneOptions_t()
    cd.check(cd.cudecompGridDescAutotuneOptionsSetDefaults(opts))
    opts.disable_nccl_backends = bool(v[0] & 1)
    opts.disable_nvshmem_backends = bool(v[1] & 1)
    res, tb, hb, pd = cd.autotune_candidates(opts, nranks, rank_order)
    assert res in VALID_RESULTS
    if res == 0 and nranks >= 1:
        assert all(p[0] * p[1] == nranks for p in pd)

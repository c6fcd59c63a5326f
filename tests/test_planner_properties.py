"""Property tests of the transfer planner over arbitrary decompositions (CPU only, one process).

For random grids, process grids (incl. non-power-of-two and slab grids up to 8 ranks), rank orders, memory orders,
gdims_dist, halos and padding, the plans of ALL ranks (handle-free planner entry points of libcudecomp.so) are executed
with numpy and must reproduce the oracle's result byte for byte -- interior cells and the cells that must stay
untouched -- for all four transposes in both schedules (direct / staged) and for halo updates of every pencil axis and
dimension. Error behaviour must agree too: empty pencils -> NOT_SUPPORTED, halo wider than a slab -> INVALID_USAGE.
"""
import itertools

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from cudecomp_b200 import capi as cd
from oracle import oracle as orc
from tests.test_host_multirank import apply_box, simulate_transpose

PDIMS = [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (1, 3), (2, 3), (3, 2), (4, 2), (2, 4), (1, 8), (8, 1), (5, 1), (3, 3)]
PERMS = [list(p) for p in itertools.permutations(range(3))]
OPS = {"XY": (0, 1), "YZ": (1, 1), "ZY": (2, -1), "YX": (1, -1)}


@st.composite
def decompositions(draw):
    pd = draw(st.sampled_from(PDIMS))
    gdims = [draw(st.integers(1, 14)) for _ in range(3)]
    layout = draw(st.sampled_from(["default", "axis_contiguous", "explicit"]))
    ac = [False] * 3
    mo = None
    if layout == "axis_contiguous":
        ac = [draw(st.booleans()) for _ in range(3)]
    elif layout == "explicit":
        mo = [draw(st.sampled_from(PERMS)) for _ in range(3)]
    dist = None
    if draw(st.booleans()):
        dist = [draw(st.integers(1, g)) for g in gdims]
    col_major = draw(st.booleans())
    halos = {str(a): [draw(st.integers(0, 2)) for _ in range(3)] for a in range(3)}
    pads = {str(a): [draw(st.integers(0, 2)) for _ in range(3)] for a in range(3)}
    return dict(gdims=gdims, pdims=list(pd), axis_contiguous=ac, mem_order=mo, gdims_dist=dist, col_major=col_major,
                halos=halos, pads=pads)


def make_config(d):
    c = cd.cudecompGridDescConfig_t()
    cd.cudecompGridDescConfigSetDefaults(c)
    c.gdims[:] = d["gdims"]
    c.pdims[:] = d["pdims"]
    if d["gdims_dist"]:
        c.gdims_dist[:] = d["gdims_dist"]
    c.rank_order = cd.CUDECOMP_RANK_ORDER_COL_MAJOR if d["col_major"] else cd.CUDECOMP_RANK_ORDER_ROW_MAJOR
    for i in range(3):
        c.transpose_axis_contiguous[i] = d["axis_contiguous"][i]
    if d["mem_order"]:
        for i in range(3):
            c.transpose_mem_order[i][:] = d["mem_order"][i]
    return c


def make_oracle(d):
    return orc.Oracle(d["gdims"], d["pdims"], d["axis_contiguous"], d["mem_order"], d["gdims_dist"], d["col_major"])


@settings(max_examples=int(__import__("os").environ.get("CDB_HYPOTHESIS_EXAMPLES", "150")), deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions())
def test_transpose_plans_equal_oracle(d):
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            with pytest.raises(cd.CudecompError) as e:
                cd.plan_transpose_boxes(cfg, 0, ax, direction, ha, hb, pa, pb)
            assert e.value.code == cd.CUDECOMP_RESULT_NOT_SUPPORTED
            continue
        rng = np.random.default_rng(7)
        ins = [rng.integers(1, 1 << 40, o.pencil_info(r, a, ha, pa).size).astype(np.int64) for r in range(n)]
        want = [np.full(o.pencil_info(r, b, hb, pb).size, -3, np.int64) for r in range(n)]
        o.transpose(op, ins, want, ha, hb, pa, pb)
        for staged in (False, True):
            plans = [cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, staged) for r in range(n)]
            outs = [np.full(w.size, -3, np.int64) for w in want]
            works = [np.full(max(o.transpose_workspace_size(), 1), -9, np.int64) for _ in range(n)]
            simulate_transpose(plans, ins, outs, works, staged)
            for r in range(n):
                assert np.array_equal(outs[r], want[r]), (d, op, staged, r)


@settings(max_examples=int(__import__("os").environ.get("CDB_HYPOTHESIS_EXAMPLES", "150")), deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions())
def test_receiver_driven_plans_equal_oracle(d):
    """cudecompB200SetTransferMode(1): every rank LOADS its blocks from the peers' inputs. Box j of rank r reads rank j's
    input with rank j's strides and writes rank r's output with rank r's strides."""
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        rng = np.random.default_rng(9)
        ins = [rng.integers(1, 1 << 40, o.pencil_info(r, a, ha, pa).size).astype(np.int64) for r in range(n)]
        want = [np.full(o.pencil_info(r, b, hb, pb).size, -3, np.int64) for r in range(n)]
        o.transpose(op, ins, want, ha, hb, pa, pb)
        outs = [np.full(w.size, -3, np.int64) for w in want]
        for r in range(n):
            pull = cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, 2)
            push = cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, 0)
            assert sorted(bx["peer_rank"] for bx in pull) == sorted(bx["peer_rank"] for bx in push)  # same communicator
            for box in pull:
                assert not box["is_unpack"]
                apply_box(box, ins[box["peer_rank"]], outs[r])
        for r in range(n):
            assert np.array_equal(outs[r], want[r]), (d, op, r)
        # through the receiver's own workspace (staged == 3): load into `work`, then the local unpack
        outs = [np.full(w.size, -3, np.int64) for w in want]
        for r in range(n):
            work = np.full(max(o.transpose_workspace_size(), 1), -9, np.int64)
            plan = cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, 3)
            for box in plan:
                if not box["is_unpack"]:
                    apply_box(box, ins[box["peer_rank"]], work)
            for box in plan:
                if box["is_unpack"]:
                    apply_box(box, work, outs[r])
        for r in range(n):
            assert np.array_equal(outs[r], want[r]), (d, op, r, "pull staged")


@settings(max_examples=int(__import__("os").environ.get("CDB_HYPOTHESIS_EXAMPLES", "150")), deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions(), st.lists(st.integers(0, 3), min_size=3, max_size=3), st.lists(st.booleans(), min_size=3, max_size=3),
       st.lists(st.integers(0, 2), min_size=3, max_size=3))
def test_halo_plans_equal_oracle(d, halo, periods, padding):
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    for ax in range(3):
        if o.has_empty_pencils(ax):
            if any(halo):
                with pytest.raises(cd.CudecompError) as e:
                    cd.plan_halo_boxes(cfg, 0, ax, 0, halo, periods, padding)
                assert e.value.code == cd.CUDECOMP_RESULT_NOT_SUPPORTED
            continue
        rng = np.random.default_rng(11)
        data = [rng.integers(1, 1 << 40, o.pencil_info(r, ax, halo, padding).size).astype(np.int64) for r in range(n)]
        for staged in (False, True):
            mine = [x.copy() for x in data]
            ref = [x.copy() for x in data]
            for dim in range(3):
                try:
                    o.halo(ax, dim, ref, halo, periods, padding)
                    oracle_ok = True
                except RuntimeError:
                    oracle_ok = False
                plans, codes = [], set()
                for r in range(n):
                    try:
                        plans.append(cd.plan_halo_boxes(cfg, r, ax, dim, halo, periods, padding, staged))
                    except cd.CudecompError as e:
                        codes.add(e.code)
                        plans.append(None)
                if not oracle_ok:
                    # the halo does not fit the thinnest slab: at least the ranks that touch it refuse (the engine then
                    # refuses on every rank, engine.cc runHalo)
                    assert codes == {cd.CUDECOMP_RESULT_INVALID_USAGE}, (d, ax, dim, halo)
                    break
                assert not codes, (d, ax, dim, codes)
                works = [np.full(max(o.halo_workspace_size(r, ax, halo), 1), -9, np.int64) for r in range(n)]
                snap = [x.copy() for x in mine]
                for r in range(n):
                    for box in plans[r]:
                        if not box["is_unpack"]:
                            apply_box(box, snap[r], (works if staged else mine)[box["peer_rank"]])
                if staged:
                    for r in range(n):
                        for box in plans[r]:
                            if box["is_unpack"]:
                                apply_box(box, works[r], mine[r])
                for r in range(n):
                    assert np.array_equal(mine[r], ref[r]), (d, ax, dim, staged, r, halo, periods, padding)


# ----------------------------------------------------------------------------------- chunked (pipelined) schedule
def _box_cells(box, which):
    idx = np.indices(box["extent"], dtype=np.int64).reshape(3, -1)
    off, strides = (box["src_offset"], box["src_stride"]) if which == "src" else (box["dst_offset"], box["dst_stride"])
    return off + sum(idx[k] * strides[k] for k in range(3))


def run_pipelined(cfg, o, op, ha, hb, pa, pb, inplace, K, ins, pull=False, column_chunks=False):
    """Executes the chunked schedule of every rank step by step the way the device would be allowed to:
    step s = all pushes of chunk s, then all unpacks scheduled for step s. Returns the output buffers, or None if
    chunking does not apply."""
    ax, direction = OPS[op]
    a, b = orc.transpose_axes(op)
    n = o.nranks
    # column chunks (plan.cc: chunks along the fastest axis when it takes no part in the transpose): element size 8 in
    # bits 8-15, bit 2 lifts the row-length floor so that the small grids of these tests qualify
    flags = int(inplace) + (2 if pull else 0) + ((8 << 8) + 4 if column_chunks else 0)
    plans = [cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, flags, K) for r in range(n)]
    if not any(plans):
        return None
    sizes = [max(o.pencil_info(r, a, ha, pa).size, o.pencil_info(r, b, hb, pb).size) for r in range(n)]
    bufs = []
    for r in range(n):
        buf = np.full(sizes[r], -3, np.int64)
        buf[:ins[r].size] = ins[r]
        bufs.append(buf)
    outs = bufs if inplace else [np.full(sizes[r], -3, np.int64) for r in range(n)]
    works = [np.full(max(o.transpose_workspace_size(), 1), -9, np.int64) for _ in range(n)]
    arrived = [np.full(w.size, -1, np.int64) for w in works]  # step in which a workspace cell was written
    for s in range(K):
        for r in range(n):
            for box in plans[r]:
                if box["step"] == s and not box["is_unpack"]:
                    dst = _box_cells(box, "dst")
                    peer = box["peer_rank"]
                    # sender-driven: r stores into peer's workspace; receiver-driven: r loads from peer's pencil
                    owner, source = (r, peer) if pull else (peer, r)
                    assert (arrived[owner][dst] == -1).all(), "a workspace cell is written twice"
                    arrived[owner][dst] = s
                    works[owner][dst] = bufs[source][_box_cells(box, "src")]
        for r in range(n):
            for box in plans[r]:
                if box["step"] == s and box["is_unpack"]:
                    src = _box_cells(box, "src")
                    assert ((arrived[r][src] >= 0) & (arrived[r][src] <= s)).all(), "unpack before the data arrived"
                    outs[r][_box_cells(box, "dst")] = works[r][src]
    return outs


@settings(max_examples=int(__import__("os").environ.get("CDB_HYPOTHESIS_EXAMPLES", "150")), deadline=None,
          suppress_health_check=list(HealthCheck))
@given(decompositions(), st.sampled_from([2, 3, 4, 8]), st.booleans(), st.booleans(), st.booleans())
def test_pipelined_schedule_equals_oracle(d, K, inplace, pull, column_chunks):
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    for op in OPS:
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        rng = np.random.default_rng(5)
        ins = [rng.integers(1, 1 << 40, o.pencil_info(r, a, ha, pa).size).astype(np.int64) for r in range(n)]
        sizes = [max(o.pencil_info(r, a, ha, pa).size, o.pencil_info(r, b, hb, pb).size) for r in range(n)]
        # oracle result with the same pre-existing buffer contents
        ref_in = []
        for r in range(n):
            buf = np.full(sizes[r], -3, np.int64)
            buf[:ins[r].size] = ins[r]
            ref_in.append(buf)
        ref_out = ref_in if inplace else [np.full(sizes[r], -3, np.int64) for r in range(n)]
        o.transpose(op, ref_in, ref_out, ha, hb, pa, pb)
        got = run_pipelined(cfg, o, op, ha, hb, pa, pb, inplace, K, ins, pull, column_chunks)
        if got is None:
            continue
        for r in range(n):
            assert np.array_equal(got[r], ref_out[r]), (d, op, K, inplace, pull, column_chunks, r)

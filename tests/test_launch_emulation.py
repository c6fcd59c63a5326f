"""The copy kernels' index arithmetic, checked without a GPU.

tests/host_emu walks every launch the way the device does -- launches prepared by the product's own host code
(csrc/launch_params.cc), slots and tiles decoded by the very functions the kernels are compiled from (csrc/tiling.h),
warps and lanes restated from csrc/kernels.cu -- on numpy buffers, with every access checked for alignment and bounds.
For random decompositions, element sizes, buffer alignments, tile sizes, slot orders, kernel variants and grid sizes
the outcome of all ranks' launches must equal the oracle byte for byte, and every destination byte must be written
exactly once. This covers what the -m gpu parity tests cover for the default schedule, and it is the only pre-hardware
check of the opt-in schedules (pairwise slot order, other tile sizes, the TMA bulk variant's 16-byte rules).
"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from cudecomp_b200 import capi as cd
from oracle import oracle as orc
from tests import host_emu as emu
from tests.test_planner_properties import OPS, decompositions, make_config, make_oracle

EXAMPLES = int(os.environ.get("CDB_HYPOTHESIS_EXAMPLES", "150"))
DT = {4: np.int32, 8: np.int64, 16: np.complex128}


def rand_fill(arr, rng):
    if arr.dtype == np.complex128:
        arr[:] = rng.integers(1, 1 << 30, arr.size) + 1j * rng.integers(1, 1 << 30, arr.size)
    else:
        arr[:] = rng.integers(1, 1 << 30, arr.size)


@st.composite
def schedules(draw):
    return dict(es=draw(st.sampled_from([4, 8, 16])), tile_bytes=draw(st.sampled_from([0, 4096, 8192, 16384, 65536])),
                peer_order=draw(st.sampled_from([0, 1])), kernel_variant=draw(st.sampled_from([0, 1, 2])),
                grid=draw(st.sampled_from([0, 1, 3, 7, 64])), threads=draw(st.sampled_from([256, 128, 64])),
                misalign=draw(st.sampled_from([0, 0, 1, 2, 3])))


@st.composite
def long_row_decompositions(draw):
    """One long axis (rows of several KiB: the warp-per-piece path, rows cut into segments, the bulk variant), the
    other two short; at most 4 ranks so that the examples stay small."""
    d = draw(decompositions())
    d["pdims"] = list(draw(st.sampled_from([(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (1, 4), (4, 1)])))
    k = draw(st.integers(0, 2))
    d["gdims"] = [draw(st.integers(1, 5)) for _ in range(3)]
    d["gdims"][k] = draw(st.sampled_from([96, 128, 250, 256, 515, 640, 1030]))
    d["gdims_dist"] = None
    return d


def group_index(plans):
    """communicator index of a destination rank = its position among the destinations of the push boxes"""
    world = sorted({b["peer_rank"] for b in plans if not b["is_unpack"]})
    return {w: i for i, w in enumerate(world)}


def emulate(boxes, src_of, dst_of, legal, s, me=-1, comm=0, peer_index=None):
    if not boxes:
        return dict(bytes_written=0, kinds=0, launches=0)
    return emu.run_boxes(boxes, [src_of(b) for b in boxes], [dst_of(b) for b in boxes], s["es"], legal, me=me,
                         comm_size=comm, peer_index=peer_index, tile_bytes=s["tile_bytes"], peer_order=s["peer_order"],
                         kernel_variant=s["kernel_variant"], grid=s["grid"], threads=s["threads"])


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions(), schedules())
def test_emulated_transposes_equal_oracle(d, s):
    check_transposes(d, s)


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(long_row_decompositions(), schedules())
def test_emulated_transposes_with_long_rows_equal_oracle(d, s):
    check_transposes(d, s)


@st.composite
def vector_transpose_decompositions(draw):
    """Permuting layouts whose extents keep every row 16-byte aligned: what the vectorised transpose kernel takes."""
    pd = draw(st.sampled_from([(1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (1, 4)]))
    gdims = [draw(st.sampled_from([8, 16, 32, 48, 64, 72, 136])) for _ in range(3)]
    layout = draw(st.sampled_from(["axis_contiguous", "explicit"]))
    ac, mo = [False] * 3, None
    if layout == "axis_contiguous":
        ac = [True, True, draw(st.booleans())]
    else:
        mo = [draw(st.sampled_from([(0, 1, 2), (1, 2, 0), (2, 0, 1), (1, 0, 2), (2, 1, 0), (0, 2, 1)])) for _ in range(3)]
    zero = {str(a): [0, 0, 0] for a in range(3)}
    halos = zero if draw(st.booleans()) else {str(a): [4 * draw(st.integers(0, 1)) for _ in range(3)] for a in range(3)}
    return dict(gdims=gdims, pdims=list(pd), axis_contiguous=ac, mem_order=mo, gdims_dist=None, col_major=False,
                halos=halos, pads=zero)


@settings(max_examples=max(EXAMPLES // 2, 30), deadline=None, suppress_health_check=list(HealthCheck))
@given(vector_transpose_decompositions(), st.sampled_from([4, 8, 16]), st.sampled_from([0, 1, 5, 64]), st.sampled_from([0, 1]),
       st.sampled_from([0, 0x100]))
def test_emulated_vectorised_transposes_equal_oracle(d, es, grid, peer_order, geometry):
    s = dict(es=es, tile_bytes=0, peer_order=peer_order, kernel_variant=geometry, grid=grid, threads=256, misalign=0)
    check_transposes(d, s)


def test_vectorised_transpose_is_selected_and_falls_back():
    """Aligned permuting boxes take the 16-byte kernel (kind bit 8); odd extents, misaligned buffers and kernel variant 3
    keep the element-wise one (bit 2) -- with identical results (checked by the property tests above)."""
    zero = {str(a): [0, 0, 0] for a in range(3)}

    def kinds(gdims, es, misalign=0, variant=0):
        d = dict(gdims=gdims, pdims=[2, 1], axis_contiguous=[True] * 3, mem_order=None, gdims_dist=None, col_major=False,
                 halos=zero, pads=zero)
        cfg, o = make_config(d), make_oracle(d)
        dt = DT[es]
        ins = [emu.aligned_array(o.pencil_info(r, 0).size, dt, misalign) for r in range(2)]
        outs = [emu.aligned_array(o.pencil_info(r, 1).size, dt, misalign) for r in range(2)]
        push = cd.plan_transpose_boxes(cfg, 0, 0, 1)
        return emu.run_boxes(push, [ins[0]] * len(push), [outs[bx["peer_rank"]] for bx in push], es, ins + outs,
                             kernel_variant=variant, me=0, comm_size=2, peer_index=[0, 1])["kinds"]

    for es in (4, 8, 16):
        assert kinds([64, 32, 16], es) == 8
        assert kinds([64, 32, 16], es, variant=3) == 2
    assert kinds([62, 30, 16], 4) == 2   # extents that are not multiples of 4 elements
    assert kinds([64, 32, 16], 4, misalign=4) == 2
    assert kinds([64, 32, 16], 16, misalign=16) == 8  # 16-byte elements are always aligned


def check_transposes(d, s):
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[s["es"]]
    off = s["misalign"] * s["es"]  # buffers start at a multiple of the element size past a 256-byte boundary
    rng = np.random.default_rng(3)
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        ins = [emu.aligned_array(o.pencil_info(r, a, ha, pa).size, dt, off) for r in range(n)]
        for x in ins:
            rand_fill(x, rng)
        want = [np.full(o.pencil_info(r, b, hb, pb).size, -3, dt) for r in range(n)]
        o.transpose(op, ins, want, ha, hb, pa, pb)
        for staged in (False, True):
            plans = [cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, staged) for r in range(n)]
            outs = [emu.aligned_array(w.size, dt, off, -3) for w in want]
            works = [emu.aligned_array(max(o.transpose_workspace_size(), 1), dt, 0, -9) for _ in range(n)]
            legal = ins + outs + works
            moved = 0
            for r in range(n):
                push = [bx for bx in plans[r] if not bx["is_unpack"]]
                gi = group_index(plans[r])
                st_ = emulate(push, lambda bx: ins[r], lambda bx: (works if staged else outs)[bx["peer_rank"]], legal, s,
                              me=gi[r], comm=len(gi), peer_index=[gi[bx["peer_rank"]] for bx in push])
                moved += st_["bytes_written"]
                assert st_["bytes_written"] == sum(int(np.prod(bx["extent"])) for bx in push) * s["es"], (d, s, op, r)
            if staged:
                for r in range(n):
                    unpack = [bx for bx in plans[r] if bx["is_unpack"]]
                    st_ = emulate(unpack, lambda bx: works[r], lambda bx: outs[r], legal, s)
                    assert st_["bytes_written"] == sum(int(np.prod(bx["extent"])) for bx in unpack) * s["es"]
            assert moved == sum(o.pencil_info(r, a).size for r in range(n)) * s["es"]  # every interior cell exactly once
            for r in range(n):
                assert np.array_equal(outs[r], want[r]), (d, s, op, staged, r)


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions(), schedules(), st.lists(st.integers(0, 3), min_size=3, max_size=3),
       st.lists(st.booleans(), min_size=3, max_size=3), st.lists(st.integers(0, 2), min_size=3, max_size=3))
def test_emulated_halos_equal_oracle(d, s, halo, periods, padding):
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[s["es"]]
    rng = np.random.default_rng(13)
    for ax in range(3):
        if o.has_empty_pencils(ax):
            continue
        data = [emu.aligned_array(o.pencil_info(r, ax, halo, padding).size, dt, s["misalign"] * s["es"]) for r in range(n)]
        for x in data:
            rand_fill(x, rng)
        for staged in (False, True):
            mine = [emu.aligned_array(x.size, dt, s["misalign"] * s["es"], 0) for x in data]
            for m, x in zip(mine, data):
                m[:] = x
            ref = [x.copy() for x in data]
            for dim in range(3):
                try:
                    o.halo(ax, dim, ref, halo, periods, padding)
                except RuntimeError:
                    break  # halo wider than a slab: covered by test_planner_properties
                plans = [cd.plan_halo_boxes(cfg, r, ax, dim, halo, periods, padding, staged) for r in range(n)]
                works = [emu.aligned_array(max(o.halo_workspace_size(r, ax, halo), 1), dt, 0, -9) for r in range(n)]
                snap = [emu.aligned_array(x.size, dt, s["misalign"] * s["es"], 0) for x in mine]
                for sn, m in zip(snap, mine):
                    sn[:] = m
                legal = mine + works + snap
                for r in range(n):
                    push = [bx for bx in plans[r] if not bx["is_unpack"]]
                    emulate(push, lambda bx: snap[r], lambda bx: (works if staged else mine)[bx["peer_rank"]], legal, s)
                if staged:
                    for r in range(n):
                        unpack = [bx for bx in plans[r] if bx["is_unpack"]]
                        emulate(unpack, lambda bx: works[r], lambda bx: mine[r], legal, s)
                for r in range(n):
                    assert np.array_equal(mine[r], ref[r]), (d, s, ax, dim, staged, r, halo, periods, padding)


@settings(max_examples=max(EXAMPLES // 3, 20), deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions(), schedules(), st.sampled_from([2, 3, 4, 8]), st.booleans(), st.booleans())
def test_emulated_pipelined_schedule_equals_oracle(d, s, K, inplace, pull):
    """The chunked schedule executed step by step (engine.cc runPipelinedStaged): push launches of step k, then the
    unpack launches of step k, every launch through the emulator."""
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[s["es"]]
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        plans = [cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, int(inplace) + (2 if pull else 0), K)
                 for r in range(n)]
        if not any(plans):
            continue
        rng = np.random.default_rng(5)
        sizes = [max(o.pencil_info(r, a, ha, pa).size, o.pencil_info(r, b, hb, pb).size) for r in range(n)]
        bufs = [emu.aligned_array(sizes[r], dt, 0, -3) for r in range(n)]
        for r in range(n):
            rand_fill(bufs[r][:o.pencil_info(r, a, ha, pa).size], rng)
        ref_in = [x.copy() for x in bufs]
        ref_out = ref_in if inplace else [np.full(sizes[r], -3, dt) for r in range(n)]
        o.transpose(op, ref_in, ref_out, ha, hb, pa, pb)
        outs = bufs if inplace else [emu.aligned_array(sizes[r], dt, 0, -3) for r in range(n)]
        works = [emu.aligned_array(max(o.transpose_workspace_size(), 1), dt, 0, -9) for _ in range(n)]
        legal = bufs + outs + works
        for step in range(K):
            for r in range(n):
                push = [bx for bx in plans[r] if bx["step"] == step and not bx["is_unpack"]]
                gi = group_index(plans[r])
                if pull:  # load from the owner's pencil into my workspace
                    emulate(push, lambda bx: bufs[bx["peer_rank"]], lambda bx: works[r], legal, s, me=gi.get(r, -1),
                            comm=len(gi), peer_index=[gi[bx["peer_rank"]] for bx in push])
                else:
                    emulate(push, lambda bx: bufs[r], lambda bx: works[bx["peer_rank"]], legal, s, me=gi.get(r, -1),
                            comm=len(gi), peer_index=[gi[bx["peer_rank"]] for bx in push])
            for r in range(n):
                unpack = [bx for bx in plans[r] if bx["step"] == step and bx["is_unpack"]]
                emulate(unpack, lambda bx: works[r], lambda bx: outs[r], legal, s)
        for r in range(n):
            assert np.array_equal(outs[r], ref_out[r]), (d, s, op, K, inplace, pull, r)


@settings(max_examples=max(EXAMPLES // 3, 20), deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions(), schedules(), st.sampled_from([1, 2, 3, 4, 8]), st.booleans(), st.sampled_from([1, 2, 3]),
       st.booleans())
def test_emulated_fused_staged_schedule_equals_oracle(d, s, K, inplace, lag, late):
    """The fused staged schedule (engine.cc runFusedStaged, kernels.cu rowCopyPhasedKernel): every rank's phased launch
    is prepared by the product's preparePhased and walked by the emulator. The kernel only orders an unpack box after
    the step it waits for, so both extremes are executed: every unpack as EARLY as its dependency allows (right after
    its step has been pushed by all ranks, before any later push has read its source), and as LATE as possible."""
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[s["es"]]
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        if list(o.pencil_info(0, a).order) != list(o.pencil_info(0, b).order):
            continue  # differing memory orders: the engine keeps separate launches (engine.cc runFusedStaged)
        if K > 1:
            # column chunks (plan.cc) on every other example: element size in bits 8-15, bit 2 lifts the row-length floor
            flags = int(inplace) + ((s["es"] << 8) + 4 if s["grid"] in (0, 3) else 0)
            plans = [cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, flags, K) for r in range(n)]
            nsteps = K
        else:
            plans = [cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, staged=1) for r in range(n)]
            nsteps = 1
        if not any(plans):
            continue
        rng = np.random.default_rng(7)
        sizes = [max(o.pencil_info(r, a, ha, pa).size, o.pencil_info(r, b, hb, pb).size) for r in range(n)]
        bufs = [emu.aligned_array(sizes[r], dt, 0, -3) for r in range(n)]
        for r in range(n):
            rand_fill(bufs[r][:o.pencil_info(r, a, ha, pa).size], rng)
        ref_in = [x.copy() for x in bufs]
        ref_out = ref_in if inplace else [np.full(sizes[r], -3, dt) for r in range(n)]
        o.transpose(op, ref_in, ref_out, ha, hb, pa, pb)
        outs = bufs if inplace else [emu.aligned_array(sizes[r], dt, 0, -3) for r in range(n)]
        works = [emu.aligned_array(max(o.transpose_workspace_size(), 1), dt, 0, -9) for _ in range(n)]
        legal = bufs + outs + works

        def run(r, want_unpack, step):
            bx = plans[r]
            srcs = [works[r] if x["is_unpack"] else bufs[r] for x in bx]
            dsts = [outs[r] if x["is_unpack"] else works[x["peer_rank"]] for x in bx]
            return emu.run_phased(bx, srcs, dsts, s["es"], legal, nsteps, lag, want_unpack, step, tile_bytes=s["tile_bytes"],
                                  grid=s["grid"], threads=s["threads"], kernel_variant=s["kernel_variant"] & 2,
                                  head_percent=[0, 25, 60][s["peer_order"] + (s["misalign"] > 1)])

        order = [(False, k) for k in range(nsteps)] + [(True, k) for k in range(nsteps)] if late else \
            [(u, k) for k in range(nsteps) for u in (False, True)]
        for want_unpack, step in order:
            for r in range(n):
                assert run(r, want_unpack, step) is not None, "equal memory orders must always give a phased launch"
        for r in range(n):
            assert np.array_equal(outs[r], ref_out[r]), (d, s, op, K, inplace, lag, late, r)


@pytest.mark.parametrize("gdims,pdims,es,K,tile_bytes,flags", [
    ([256, 64, 48], [2, 2], 16, 4, 4096, 0),             # plane chunks, boxes of several hundred tiles
    ([1024, 24, 20], [2, 2], 16, 4, 4096, (16 << 8)),    # column chunks with the engine's 2 KiB row rule (Y<->Z along x)
    ([512, 40, 36], [1, 4], 8, 8, 8192, (8 << 8)),       # 4 KiB rows: column chunks reduce 8 chunks to 2
    ([250, 60, 44], [4, 1], 4, 5, 4096, 0),              # uneven splits, 4-byte elements
])
def test_emulated_fused_schedule_with_many_segments(gdims, pdims, es, K, tile_bytes, flags):
    """The hypothesis examples above are tiny (every box is one tile, one segment). Here boxes have hundreds of tiles, so
    phases hold many 64-tile segments per box, partly filled last segments and the head / tail split of the pushes; in
    place, earliest legal order, head 25 % and lag 1 as the engine runs it."""
    d = dict(gdims=gdims, pdims=pdims, axis_contiguous=[False] * 3, mem_order=None, gdims_dist=None, col_major=False,
             halos={str(a): [0, 0, 0] for a in range(3)}, pads={str(a): [0, 0, 0] for a in range(3)})
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[es]
    rng = np.random.default_rng(11)
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        plans = [cd.plan_pipelined_transpose_boxes(cfg, r, ax, direction, None, None, None, None, 1 + flags, K)
                 for r in range(n)]
        if not any(plans):
            continue
        nsteps = 1 + max(bx["step"] for p in plans for bx in p)
        sizes = [max(o.pencil_info(r, a).size, o.pencil_info(r, b).size) for r in range(n)]
        bufs = [emu.aligned_array(sizes[r], dt, 0, -3) for r in range(n)]
        for r in range(n):
            rand_fill(bufs[r][:o.pencil_info(r, a).size], rng)
        ref = [x.copy() for x in bufs]
        o.transpose(op, ref, ref)
        works = [emu.aligned_array(max(o.transpose_workspace_size(), 1), dt, 0, -9) for _ in range(n)]
        legal = bufs + works
        seen_segments = 0
        for step in range(nsteps):
            for want_unpack in (False, True):
                for r in range(n):
                    bx = plans[r]
                    srcs = [works[r] if x["is_unpack"] else bufs[r] for x in bx]
                    dsts = [bufs[r] if x["is_unpack"] else works[x["peer_rank"]] for x in bx]
                    st_ = emu.run_phased(bx, srcs, dsts, es, legal, nsteps, 1, want_unpack, step, tile_bytes=tile_bytes,
                                         grid=37, kernel_variant=2, head_percent=25)
                    assert st_ is not None
                    seen_segments = max(seen_segments, st_["slots"])
        assert seen_segments > 64 * 4  # more slots than one segment per box: the segment interleave was exercised
        for r in range(n):
            assert np.array_equal(bufs[r], ref[r]), (op, r)


def test_kernel_selection_and_vector_width():
    """Default-layout transposes are row copies at the widest vector the alignment allows; differing memory orders go
    through the tiled transpose kernel; the bulk variant only takes over for 16-byte aligned rows of at least 2 KiB."""
    d = dict(gdims=[256, 8, 6], pdims=[2, 1], axis_contiguous=[False] * 3, mem_order=None, gdims_dist=None, col_major=False,
             halos={str(a): [0, 0, 0] for a in range(3)}, pads={str(a): [0, 0, 0] for a in range(3)})
    cfg, o = make_config(d), make_oracle(d)

    def run(es, variant, misalign=0, dd=d, cfg_=cfg, o_=o):
        dt = DT[es]
        ins = [emu.aligned_array(o_.pencil_info(r, 0).size, dt, misalign) for r in range(2)]
        outs = [emu.aligned_array(o_.pencil_info(r, 1).size, dt, misalign) for r in range(2)]
        push = cd.plan_transpose_boxes(cfg_, 0, 0, 1)
        return emu.run_boxes(push, [ins[0]] * len(push), [outs[bx["peer_rank"]] for bx in push], es, ins + outs,
                             kernel_variant=variant, me=0, comm_size=2, peer_index=[0, 1])

    st16 = run(16, 0)
    assert st16["kinds"] == 1 and st16["vec"] == 16
    assert run(4, 0)["vec"] == 16          # 128 floats per row: still 16-byte vectors
    assert run(8, 0, misalign=8)["vec"] == 8  # buffers only 8-byte aligned
    assert run(4, 0, misalign=4)["vec"] == 4
    assert run(16, 1)["kinds"] == 4        # rows of 128 x 16 B = 2 KiB: TMA bulk
    assert run(8, 1)["kinds"] == 1         # 1 KiB rows: stays SIMT
    assert run(16, 2)["vec"] == 32         # variant 2: 256-bit accesses, everything here is 32-byte aligned
    assert run(16, 2, misalign=16)["vec"] == 16  # ... unless a buffer is only 16-byte aligned
    assert run(16, 1)["accesses"] == 2 * 4 * 6   # one bulk copy per row segment: 2 peers x (4 x 6) rows of my pencil
    # axis-contiguous layouts permute: the tiled transpose kernels (16-byte accesses when rows are aligned, bit 8)
    d2 = dict(d, axis_contiguous=[True] * 3)
    cfg2, o2 = make_config(d2), make_oracle(d2)
    assert run(8, 0, dd=d2, cfg_=cfg2, o_=o2)["kinds"] == 8
    assert run(8, 3, dd=d2, cfg_=cfg2, o_=o2)["kinds"] == 2
    assert run(8, 0, misalign=8, dd=d2, cfg_=cfg2, o_=o2)["kinds"] == 2


def test_contiguous_axes_are_merged_into_long_rows():
    """Y->Z in the default layout moves, per z-plane, a block of x-rows that is contiguous on both sides: the planner
    must hand the kernel ONE row per plane (2 MB at the headline size), not one row per x-line -- the per-row index
    arithmetic and the tile tails depend on it. The results would be the same either way, so only this test sees it."""
    d = dict(gdims=[32, 24, 10], pdims=[1, 2], axis_contiguous=[False] * 3, mem_order=None, gdims_dist=None, col_major=False,
             halos={str(a): [0, 0, 0] for a in range(3)}, pads={str(a): [0, 0, 0] for a in range(3)})
    cfg, o = make_config(d), make_oracle(d)
    push = cd.plan_transpose_boxes(cfg, 0, 1, 1)  # Y->Z: y is split in two blocks of 12
    src = emu.aligned_array(o.pencil_info(0, 1).size, np.int64, 0, 1)
    outs = [emu.aligned_array(o.pencil_info(r, 2).size, np.int64, 0, 0) for r in range(2)]
    st_ = emu.run_boxes(push, [src] * 2, [outs[bx["peer_rank"]] for bx in push], 8, [src] + outs, me=0, comm_size=2,
                        peer_index=[0, 1])
    assert st_["kinds"] == 1 and st_["longest_row"] == 32 * 12 * 8  # x-lines of one y-block merged: 32 x 12 elements
    # X->Y cannot merge (the x-range is cut in the source): rows stay x-lines of half the x extent
    d2 = dict(d, pdims=[2, 1])
    cfg2, o2 = make_config(d2), make_oracle(d2)
    push = cd.plan_transpose_boxes(cfg2, 0, 0, 1)
    src = emu.aligned_array(o2.pencil_info(0, 0).size, np.int64, 0, 1)
    outs = [emu.aligned_array(o2.pencil_info(r, 1).size, np.int64, 0, 0) for r in range(2)]
    st_ = emu.run_boxes(push, [src] * 2, [outs[bx["peer_rank"]] for bx in push], 8, [src] + outs, me=0, comm_size=2,
                        peer_index=[0, 1])
    assert st_["longest_row"] == 16 * 8


def test_slot_orders_are_permutations():
    import ctypes
    lib = emu.lib()
    for nboxes, max_tiles in [(1, 5), (2, 7), (4, 1), (4, 33), (8, 9), (16, 3)]:
        for order in (0, 1):
            seen = set()
            for t in range(nboxes * max_tiles):
                b, j = ctypes.c_uint32(), ctypes.c_uint32()
                lib.cdb_emu_slot(ctypes.c_uint32(t), ctypes.c_uint32(nboxes), ctypes.c_uint32(max_tiles),
                                 ctypes.c_uint32(order), ctypes.byref(b), ctypes.byref(j))
                assert b.value < nboxes and j.value < max_tiles
                seen.add((b.value, j.value))
            assert len(seen) == nboxes * max_tiles
    # pairwise: between two local slots the remote slots walk one peer after the other
    nboxes, max_tiles = 4, 30
    order = []
    for t in range(nboxes * max_tiles):
        b, j = ctypes.c_uint32(), ctypes.c_uint32()
        lib.cdb_emu_slot(ctypes.c_uint32(t), ctypes.c_uint32(nboxes), ctypes.c_uint32(max_tiles), ctypes.c_uint32(1),
                         ctypes.byref(b), ctypes.byref(j))
        if b.value:
            order.append(b.value)
    assert order == sorted(order) and order.count(1) == order.count(2) == order.count(3) == max_tiles


def test_balanced_grid():
    g = emu.lib().cdb_emu_choose_grid
    assert g(0, 370, 444, 4096, 0) == 370                 # default: 2.5 CTAs per SM
    assert g(0, 370, 444, 100, 0) == 100                  # never more CTAs than slots
    assert g(1000, 370, 444, 1 << 20, 0) == 444           # capped by residency
    assert g(0, 370, 444, 0, 0) == 1                      # an empty launch still handshakes
    b = g(0, 370, 444, 4096, 1)                           # 4096 slots: 12 full rounds of 342 CTAs instead of 11.07 of 370
    assert 296 <= b <= 370 and -(-4096 // b) * b - 4096 < 10
    assert g(0, 370, 444, 370 * 50, 1) == 370             # already even: unchanged
    assert g(0, 370, 444, 3600, 1) == 360                 # 300 and 360 both divide 3600: the larger count wins the tie
    for slots in (371, 1000, 4097, 65536, 99991):
        b = g(0, 370, 444, slots, 1)
        assert 296 <= b <= 370
        util = lambda c: slots / (c * -(-slots // c))
        assert util(b) >= util(370) - 1e-12


def test_seventeen_rank_communicator_needs_two_launches():
    """17 ranks in one communicator = 17 boxes per transpose, one more than a launch carries (kMaxBoxes = 16): the boxes
    are split over two launches (the handshake enters with the first and leaves with the last, engine.cc launchBoxes)."""
    for pdims in ([17, 1], [1, 17]):
        d = dict(gdims=[40, 36, 38], pdims=pdims, axis_contiguous=[False] * 3, mem_order=None, gdims_dist=None,
                 col_major=False, halos={str(a): [0, 0, 0] for a in range(3)}, pads={str(a): [0, 0, 0] for a in range(3)})
        s = dict(es=8, tile_bytes=0, peer_order=1, kernel_variant=0, grid=5, threads=256, misalign=0)
        check_transposes(d, s)
        cfg, o = make_config(d), make_oracle(d)
        op = "XY" if pdims[0] == 17 else "YZ"
        ax, direction = OPS[op]
        a, b = orc.transpose_axes(op)
        push = cd.plan_transpose_boxes(cfg, 3, ax, direction)
        assert len(push) == 17
        src = emu.aligned_array(o.pencil_info(3, a).size, np.int64, 0, 1)
        outs = {bx["peer_rank"]: emu.aligned_array(o.pencil_info(bx["peer_rank"], b).size, np.int64, 0, 0) for bx in push}
        gi = group_index(push)
        st_ = emu.run_boxes(push, [src] * 17, [outs[bx["peer_rank"]] for bx in push], 8, [src] + list(outs.values()),
                            me=gi[3], comm_size=17, peer_index=[gi[bx["peer_rank"]] for bx in push])
        assert st_["launches"] == 2


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.one_of(decompositions(), long_row_decompositions()), schedules())
def test_emulated_receiver_driven_transposes_equal_oracle(d, s):
    """Receiver-driven direct plans (cudecompB200SetTransferMode): the same kernels with the SOURCE in a peer's buffer."""
    cfg, o = make_config(d), make_oracle(d)
    n = o.nranks
    dt = DT[s["es"]]
    off = s["misalign"] * s["es"]
    rng = np.random.default_rng(21)
    for op, (ax, direction) in OPS.items():
        a, b = orc.transpose_axes(op)
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        ha, hb, pa, pb = d["halos"][str(a)], d["halos"][str(b)], d["pads"][str(a)], d["pads"][str(b)]
        ins = [emu.aligned_array(o.pencil_info(r, a, ha, pa).size, dt, off) for r in range(n)]
        for x in ins:
            rand_fill(x, rng)
        want = [np.full(o.pencil_info(r, b, hb, pb).size, -3, dt) for r in range(n)]
        o.transpose(op, ins, want, ha, hb, pa, pb)
        outs = [emu.aligned_array(w.size, dt, off, -3) for w in want]
        for r in range(n):
            pull = cd.plan_transpose_boxes(cfg, r, ax, direction, ha, hb, pa, pb, 2)
            gi = group_index(pull)
            st_ = emulate(pull, lambda bx: ins[bx["peer_rank"]], lambda bx: outs[r], ins + outs, s, me=gi[r], comm=len(gi),
                          peer_index=[gi[bx["peer_rank"]] for bx in pull])
            assert st_["bytes_written"] == o.pencil_info(r, b).size * s["es"]  # my whole output interior, exactly once
        for r in range(n):
            assert np.array_equal(outs[r], want[r]), (d, s, op, r)

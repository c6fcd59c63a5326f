"""N>1 host logic without a GPU: real processes, the library's bootstrap, geometry and transfer planning.

Each rank asks libcudecomp.so (through the C ABI) for its pencil infos and for the transfer plan it would execute;
the parent then *executes the plans with numpy* -- every box copied from the sender's pencil into the destination
rank's buffer with the strides the sender computed for that peer -- and compares with the oracle / the analytic
pattern. This covers everything of the multi-rank path except the CUDA kernels themselves.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import cases as C
from tests._launcher import run_ranks


def apply_box(box, src, dst):
    ext = box["extent"]
    idx = np.indices(ext, dtype=np.int64).reshape(3, -1)
    s = box["src_offset"] + sum(idx[k] * box["src_stride"][k] for k in range(3))
    d = box["dst_offset"] + sum(idx[k] * box["dst_stride"][k] for k in range(3))
    assert np.unique(d).size == d.size  # a box never writes a cell twice
    dst[d] = src[s]


def simulate_transpose(plans, inputs, outputs, works, staged):
    n = len(plans)
    for r in range(n):
        for box in plans[r]:
            if box["is_unpack"]:
                continue
            apply_box(box, inputs[r], (works if staged else outputs)[box["peer_rank"]])
    if staged:
        for r in range(n):
            for box in plans[r]:
                if box["is_unpack"]:
                    assert box["peer_rank"] == r
                    apply_box(box, works[r], outputs[r])


def check_transposes(case, results):
    o = orc.Oracle(case["gdims"], case["pdims"], case.get("axis_contiguous") or (False,) * 3, case.get("mem_order"),
                   case.get("gdims_dist"), case.get("rank_order", 0) == 2)
    n = o.nranks
    halos, pads = case.get("halos") or {}, case.get("pads") or {}
    for r in range(n):
        for ax in range(3):
            got = results[r]["pencils"][str(ax)]
            want = o.pencil_info(r, ax, halos.get(str(ax)), pads.get(str(ax)))
            assert tuple(map(tuple, got[:6])) + (got[6],) == want.as_tuple(), (case["name"], r, ax)
        assert results[r]["workspace"] == o.transpose_workspace_size()
    dt = np.float64
    for op in ("XY", "YZ", "ZY", "YX"):
        a, b = orc.transpose_axes(op)
        ha, hb, pa_, pb_ = halos.get(str(a)), halos.get(str(b)), pads.get(str(a)), pads.get(str(b))
        for staged in (False, True):
            plans = [results[r]["transposes"]["%s/%d" % (op, staged)] for r in range(n)]
            ins, outs, want = [], [], []
            for r in range(n):
                pa, pb = o.pencil_info(r, a, ha, pa_), o.pencil_info(r, b, hb, pb_)
                ins.append(orc.pattern_pencil(pa, case["gdims"], dt))
                outs.append(np.full(pb.size, -3.0, dt))
                want.append(pb)
            works = [np.full(o.transpose_workspace_size(), -9.0, dt) for _ in range(n)]
            simulate_transpose(plans, ins, outs, works, staged)
            ref_out = [np.full(w.size, -3.0, dt) for w in want]
            o.transpose(op, ins, ref_out, ha, hb, pa_, pb_)
            for r in range(n):
                assert orc.interior_equal(want[r], orc.pattern_pencil(want[r], case["gdims"], dt), outs[r]), \
                    (case["name"], op, staged, r)
                assert np.array_equal(outs[r], ref_out[r]), (case["name"], op, staged, r)  # incl. untouched cells
            # wire accounting: every interior cell of the source is sent exactly once
            for r in range(n):
                sent = sum(int(np.prod(bx["extent"])) for bx in plans[r] if not bx["is_unpack"])
                pa = o.pencil_info(r, a)
                assert sent == pa.size


def check_halos(case, results):
    o = orc.Oracle(case["gdims"], case["pdims"], case.get("axis_contiguous") or (False,) * 3, case.get("mem_order"),
                   None, case.get("rank_order", 0) == 2)
    n = o.nranks
    dt = np.float64
    halo, per, pad = case["halo"], case.get("periods") or [False] * 3, case.get("padding")
    for r in range(n):  # workspace sizes through the C ABI against the oracle (reference src/cudecomp.cc:1434-1459)
        assert results[r]["halo_workspace"] == [o.halo_workspace_size(r, ax, halo) for ax in range(3)], (case["name"], r)
    for ax in range(3):
        for staged in (False, True):
            data = [orc.pattern_pencil(o.pencil_info(r, ax, halo, pad), case["gdims"], dt) for r in range(n)]
            ref = [d.copy() for d in data]
            for dim in range(3):
                plans = [results[r]["halos"]["%d/%d/%d" % (ax, dim, staged)] for r in range(n)]
                works = [np.full(max(o.halo_workspace_size(r, ax, halo), 1), -9.0, dt) for r in range(n)]
                snapshot = [d.copy() for d in data]  # all ranks read their faces before anybody's halo is written
                for r in range(n):
                    for box in plans[r]:
                        if not box["is_unpack"]:
                            apply_box(box, snapshot[r], (works if staged else data)[box["peer_rank"]])
                if staged:
                    for r in range(n):
                        for box in plans[r]:
                            if box["is_unpack"]:
                                apply_box(box, works[r], data[r])
                o.halo(ax, dim, ref, halo, per, pad)
            for r in range(n):
                want = orc.halo_reference(o.pencil_info(r, ax, halo, pad), case["gdims"], dt, per)
                assert np.array_equal(data[r], want), (case["name"], ax, staged, r)
                assert np.array_equal(data[r], ref[r])


PLAN_CASES_2x2 = [
    dict(name="default", gdims=C.GDIMS, pdims=[2, 2], halo=C.HALO, periods=[True] * 3),
    dict(name="axis_contiguous", gdims=C.GDIMS, pdims=[2, 2], axis_contiguous=[True] * 3, halo=C.HALO,
         periods=[False] * 3),
    dict(name="col_major_halo_pad", gdims=C.GDIMS, pdims=[2, 2], rank_order=2,
         halos={"0": [1, 2, 1], "1": [2, 1, 1], "2": [1, 1, 1]}, pads={"0": [1, 1, 2], "1": [2, 1, 1], "2": [0, 1, 0]},
         mem_order=C.SPLIT_UNPACK_ORDER, halo=C.HALO, periods=[True, False, True], padding=[1, 0, 2]),
    dict(name="gdims_dist", gdims=C.GDIMS, pdims=[2, 2], gdims_dist=[8, 9, 10], mem_order=[[2, 0, 1], [1, 0, 2], [0, 2, 1]]),
    dict(name="slab_4x1", gdims=[8, 12, 10], pdims=[4, 1], halo=[1, 1, 1], periods=[True] * 3),
    dict(name="slab_1x4", gdims=[8, 12, 10], pdims=[1, 4], axis_contiguous=[True, False, True], halo=[2, 1, 2],
         periods=[True] * 3),
]


@pytest.fixture(scope="module")
def plan_results_4():
    results, _ = run_ranks(4, "plan", PLAN_CASES_2x2, timeout=300)
    return results


@pytest.mark.parametrize("i", range(len(PLAN_CASES_2x2)), ids=[c["name"] for c in PLAN_CASES_2x2])
def test_plans_on_4_ranks(plan_results_4, i):
    case = PLAN_CASES_2x2[i]
    per_rank = [plan_results_4[r][i] for r in range(4)]
    assert all(p["ok"] for p in per_rank), per_rank
    check_transposes(case, per_rank)
    if case.get("halo"):
        check_halos(case, per_rank)


def test_golden_tables_through_the_c_abi_on_4_ranks(golden):
    """The reference's ApiGetPencilInfoTest / ApiGetShiftedRankTest (api_tests.cc:1248-1290,1380-1408)."""
    halos = {str(a): golden["halo_extents"] for a in range(3)}
    pads = {str(a): golden["padding"] for a in range(3)}
    cases = [
        dict(name="default", gdims=golden["gdims"], pdims=golden["pdims"], halos=halos, pads=pads,
             shifted=golden["shifted_ranks"]["row_major"]),
        dict(name="column_major", gdims=golden["gdims"], pdims=golden["pdims"], halos=halos, pads=pads, rank_order=2,
             shifted=golden["shifted_ranks"]["col_major"]),
        dict(name="gdims_dist", gdims=golden["gdims"], pdims=golden["pdims"], halos=halos, pads=pads,
             gdims_dist=golden["gdims_dist"]),
    ]
    results, _ = run_ranks(4, "plan", cases, timeout=300)
    for i, case in enumerate(cases):
        table = golden["pencil_info"][case["name"]]
        for r in range(4):
            for ax in range(3):
                shape, lo, hi, order, halo, pad, size = results[r][i]["pencils"][str(ax)]
                want = table[ax][r]
                assert (shape, lo, hi, order, halo, pad, size) == (
                    want["shape"], want["lo"], want["hi"], want["order"], want["halo_extents"], want["padding"],
                    want["size"]), (case["name"], ax, r)
            for q, got in zip(case.get("shifted", []), results[r][i]["shifted"]):
                assert got == q["expected"][r], (case["name"], q, r)


def test_two_ranks_with_gloo_cross_check():
    """world_size 2, both 1x2 and 2x1 (BASELINE config 1: X<->Y only communicates on 2x1, Y<->Z only on 1x2)."""
    cases = [dict(name="1x2", gdims=[16, 12, 20], pdims=[1, 2], halo=[1, 1, 1], periods=[True] * 3),
             dict(name="2x1", gdims=[16, 12, 20], pdims=[2, 1], halo=[1, 1, 1], periods=[True] * 3)]
    results, _ = run_ranks(2, "plan", cases, timeout=300, extra_env={"CDB_TEST_GLOO": "1"})
    for r in range(2):
        assert results[r]["gathered"][r] == results[r]["mine"]
        assert results[r]["gathered"] == results[0]["gathered"]
    for i, case in enumerate(cases):
        per_rank = [results[r]["mine"][i] for r in range(2)]
        check_transposes(case, per_rank)
        check_halos(case, per_rank)
        # which operations cross ranks: peers other than self appear only on the communicating axis
        for r in range(2):
            xy = per_rank[r]["transposes"]["XY/0"]
            yz = per_rank[r]["transposes"]["YZ/0"]
            assert len(xy) == (2 if case["name"] == "2x1" else 1)
            assert len(yz) == (2 if case["name"] == "1x2" else 1)


def test_mpi_shim_collectives_on_4_ranks():
    """The MPI subset exported for callers without MPI (include/mpi_shim/mpi.h), exercised through the C symbols."""
    results, _ = run_ranks(4, "shim", [dict(name="battery")], timeout=300)
    for r in range(4):
        out = results[r][0]
        assert out.get("ok", True) is not False, out
        assert out["allreduce_int_sum"] == [1 + 2 + 3 + 4, 10 * (0 + 1 + 2 + 3)]
        assert out["allreduce_double_max_inplace"] == 4.5
        assert out["allreduce_float_min"] == -2.0
        assert out["allreduce_lor"] == 1
        assert out["bcast"] == [7, 8, 9]
        assert out["allgather"] == [100, 101, 102, 103]
        assert out["split_size"] == 2 and out["dup_size"] == 2
        # key = -rank reverses the order inside each colour
        assert out["split_rank"] == (1 if r < 2 else 0)
        assert out["split_sum"] == (0 + 2 if r % 2 == 0 else 1 + 3)
        assert out["freed"] == [0, 0]
        assert out["subcomm_handle"] == [0, 0] and out["subcomm_shape"] == [8, 8, 4]
        assert out["wtime_positive"]
        assert out["allreduce_int64_sum"] == [4 * (1 << 40) + 1000 * (0 + 1 + 2 + 3) + 4 * k for k in range(5)]
        assert out["allreduce_unsigned_max"] == 0x80000003 and out["allreduce_int_min_negative"] == -5
        assert out["allgather_inplace"] == [v for q in range(4) for v in ((1 << 33) + q, -q)]
        assert out["allreduce_self"] == [1.25, 2.5, float(r)] and out["comm_free_null"] == 0
    assert results[0][0]["gather_root0"] == [100, 101, 102, 103]
    assert results[0][0]["reduce_max_root0"] == 9


@pytest.mark.parametrize("nranks", [2, 4, 6, 8])
def test_descriptor_mailbox_stress(nranks):
    """The shared-memory mailbox every multi-rank operation relies on: thousands of exchanges on both channels with
    changing row/column group shapes, resets and randomised rank delays; every message is verified in the library."""
    results, _ = run_ranks(nranks, "mailbox", [dict(name="stress", iterations=4000, seed=s) for s in (1, 2, 3)],
                           timeout=300)
    for r in range(nranks):
        assert all(c["ok"] for c in results[r]), results[r]


@pytest.mark.parametrize("nranks", [1, 2, 5])
def test_cumem_request_is_settled_collectively_at_init(nranks):
    """CUDECOMP_ENABLE_CUMEM (reference docs/env_vars.rst, src/cudecomp.cc:596-660): without the variable the feature is
    off; with it the ranks check together (a) that file descriptors -- the POSIX-fd handles of cuMem allocations -- can
    be passed between them and (b) that the device supports VMM allocations. On this GPU-less host (b) fails on every
    rank alike, the reference's warning is printed once, and allocation stays on cudaMalloc; (a) must work among the
    processes of one host."""
    results, logs = run_ranks(nranks, "cumem", [dict(name="probe")], timeout=120)
    for r in range(nranks):
        assert results[r][0] == dict(ok=True, state=0, fd_passing=1), results[r]
    assert "CUDECOMP:WARN" not in logs[0]
    results, logs = run_ranks(nranks, "cumem", [dict(name="probe")], timeout=120,
                              extra_env=dict(CUDECOMP_ENABLE_CUMEM="1"))
    for r in range(nranks):
        assert results[r][0] == dict(ok=True, state=3, fd_passing=1), results[r]
    assert "CUDECOMP:WARN: CUDECOMP_ENABLE_CUMEM is set but the current device does not support CUDA VMM" in logs[0]
    assert all("CUDECOMP:WARN" not in logs[r] for r in range(1, nranks))
    # CUDECOMP_ENABLE_NCCL_UBR asks for the same allocation path (reference src/cudecomp.cc:603)
    results, _ = run_ranks(nranks, "cumem", [dict(name="probe")], timeout=120, extra_env=dict(CUDECOMP_ENABLE_NCCL_UBR="1"))
    assert all(results[r][0]["state"] == 3 for r in range(nranks))


def test_cumem_request_on_a_subset_of_the_ranks_is_dropped(tmp_path):
    """The variable set on rank 0 only: every rank must still take the same path (off), with a warning."""
    import os
    import subprocess
    import sys
    from tests._launcher import free_port, ROOT
    port = free_port()
    procs = []
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cudecomp_b200 import capi as cd\n"
            "assert cd.MPI_Init() == 0\n"
            "res, h = cd.cudecompInit(cd.MPI_COMM_WORLD); cd.check(res)\n"
            "print('state', cd.cumem_state(h)); cd.check(cd.cudecompFinalize(h)); cd.MPI_Finalize()\n") % ROOT
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        env.pop("CUDECOMP_ENABLE_CUMEM", None)
        if r == 0:
            env["CUDECOMP_ENABLE_CUMEM"] = "1"
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("state 0" in o for o in outs), outs
    assert "CUDECOMP_ENABLE_CUMEM is not set on every rank" in outs[0]


def test_sixteen_ranks_plans_and_mailbox():
    """Communicators of 16 ranks (one launch carries at most 16 boxes, kernels.h kMaxBoxes): 4x4 and both slab grids,
    planned by 16 real processes, executed with numpy and compared with the oracle; then the mailbox with 16 members."""
    cases = [dict(kind="transpose", name="p4x4", gdims=[33, 34, 35], pdims=[4, 4], dtype="double"),
             dict(kind="transpose", name="p16x1", gdims=[40, 34, 35], pdims=[16, 1], dtype="double"),
             dict(kind="transpose", name="p1x16_axis_contiguous", gdims=[40, 34, 35], pdims=[1, 16], dtype="double",
                  axis_contiguous=[True] * 3)]
    results, _ = run_ranks(16, "plan", cases, timeout=500)
    for i, case in enumerate(cases):
        check_transposes(case, [r[i] for r in results])
    results, _ = run_ranks(16, "mailbox", [dict(name="stress16", iterations=2000, seed=9)], timeout=500)
    assert all(r[0]["ok"] for r in results)


def test_rendezvous_survives_stray_connections():
    """Anything may dial the rendezvous port while the ranks are still arriving (port scanners, health checks, a rank of
    another job): connections that close at once, send garbage, or send half a greeting and then go silent must be
    dropped -- the job neither fails nor hangs (csrc/bootstrap.cc recvGreeting)."""
    import os
    import socket
    import subprocess
    import sys
    import time
    from tests._launcher import free_port, ROOT
    port = free_port()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cudecomp_b200 import capi as cd\n"
            "assert cd.MPI_Init() == 0\n"
            "res, h = cd.cudecompInit(cd.MPI_COMM_WORLD); cd.check(res)\n"
            "cd.MPI_Barrier(); cd.check(cd.cudecompFinalize(h)); cd.MPI_Finalize(); print('rank done')\n") % ROOT

    def start(rank):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        return subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                text=True)

    p0 = start(0)
    procs = [p0]
    strays = []
    try:
        deadline = time.time() + 60
        while True:  # rank 0 listens on MASTER_PORT + 1
            try:
                s = socket.create_connection(("127.0.0.1", port + 1), timeout=1)
                break
            except OSError:
                assert time.time() < deadline and p0.poll() is None, "rank 0 never listened"
                time.sleep(0.05)
        s.close()                                    # closes at once
        s = socket.create_connection(("127.0.0.1", port + 1))
        s.sendall(b"GET / HTTP/1.0\r\n\r\n")         # a full greeting's worth of garbage
        strays.append(s)
        s = socket.create_connection(("127.0.0.1", port + 1))
        s.sendall(b"\x01\x02\x03")                   # half a greeting, then silence
        strays.append(s)
        p1 = start(1)
        procs.append(p1)
        out1 = p1.communicate(timeout=120)[0]
        out0 = p0.communicate(timeout=120)[0]
        assert p0.returncode == 0 and "rank done" in out0, out0
        assert p1.returncode == 0 and "rank done" in out1, out1
    finally:
        for s in strays:
            s.close()
        for p in procs:
            if p.poll() is None:
                p.kill()

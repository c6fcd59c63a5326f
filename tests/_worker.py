"""One rank of a multi-rank test run (started by tests/_launcher.py).

mode 'gpu'  : runs each case through the C ABI of libcudecomp.so on this rank's GPU and checks the result against
              (a) the analytic global-index pattern of the reference's tests and (b) the CPU oracle fed with the
              same seeded inputs. Ranks share a GPU when there are fewer GPUs than ranks (the reference's tests do
              the same under MPS, tests/README.md:69-82).
mode 'api'  : host only. The reference's API contract tests (tests/_api_battery.py); 'api_gpu' adds the parts that
              allocate device memory.
mode 'plan' : host only. Dumps the rank's pencil infos and transfer plans so the parent can execute the plans with
              numpy and compare with the oracle (covers the N>1 planning logic without a GPU).
"""
import ctypes
import json
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cudecomp_b200 import capi as cd  # noqa: E402
from oracle import oracle as orc  # noqa: E402

DTYPES = {"float": (cd.CUDECOMP_FLOAT, np.float32), "double": (cd.CUDECOMP_DOUBLE, np.float64),
          "float_complex": (cd.CUDECOMP_FLOAT_COMPLEX, np.complex64),
          "double_complex": (cd.CUDECOMP_DOUBLE_COMPLEX, np.complex128)}
OPS = ["XY", "YZ", "ZY", "YX"]


def make_config(case):
    c = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(c))
    c.gdims[:] = case["gdims"]
    c.pdims[:] = case.get("pdims", [0, 0])
    if case.get("gdims_dist"):
        c.gdims_dist[:] = case["gdims_dist"]
    c.rank_order = case.get("rank_order", 0)
    c.transpose_comm_backend = case.get("backend", cd.CUDECOMP_TRANSPOSE_COMM_NCCL)
    c.halo_comm_backend = case.get("halo_backend", cd.CUDECOMP_HALO_COMM_NCCL)
    ac = case.get("axis_contiguous") or [False] * 3
    for i in range(3):
        c.transpose_axis_contiguous[i] = bool(ac[i])
    mo = case.get("mem_order")
    if mo:
        for i in range(3):
            for j in range(3):
                c.transpose_mem_order[i][j] = mo[i][j]
    return c


def make_oracle(case, pdims=None):
    return orc.Oracle(case["gdims"], pdims or case["pdims"], case.get("axis_contiguous") or (False,) * 3,
                      case.get("mem_order"), case.get("gdims_dist"), case.get("rank_order", 0) == 2)


def pinfo_to_py(p):
    return orc.PencilInfo(p.shape, p.lo, p.hi, p.order, p.halo_extents, p.padding, p.size)


def seeded(n, np_dtype, seed):
    """Random bit patterns (finite values) -- the engine never computes on the payload."""
    rng = np.random.default_rng(seed)
    if np.issubdtype(np_dtype, np.complexfloating):
        base = np.float32 if np_dtype == np.complex64 else np.float64
        return (rng.standard_normal(n).astype(base) + 1j * rng.standard_normal(n).astype(base)).astype(np_dtype)
    return rng.standard_normal(n).astype(np_dtype)


# ------------------------------------------------------------------------------------------------- gpu mode
class Gpu:
    def __init__(self):
        import torch
        self.torch = torch
        ndev = torch.cuda.device_count()
        if ndev == 0:
            raise RuntimeError("no CUDA device")
        self.dev = int(os.environ.get("LOCAL_RANK", "0")) % ndev
        torch.cuda.set_device(self.dev)

    def upload(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy())
        return t.cuda(self.dev)

    def download(self, t, np_dtype):
        self.torch.cuda.synchronize()
        return t.cpu().numpy().view(np_dtype).copy()

    def empty(self, nbytes):
        return self.torch.zeros(max(int(nbytes), 16), dtype=self.torch.uint8, device="cuda:%d" % self.dev)


def transpose_case(gpu, handle, rank, case):
    """Single op or a chain of ops; every step is checked."""
    dt_enum, np_dtype = DTYPES[case.get("dtype", "float")]
    es = np.dtype(np_dtype).itemsize
    cfg = make_config(case)
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    if res != case.get("expect_create", 0):
        return dict(ok=False, msg="GridDescCreate returned %d" % res)
    if res != 0:
        return dict(ok=True, msg="create failed as expected")
    out = dict(ok=True, msg="", paths=[])
    work_ptr = 0
    try:
        o = make_oracle(case)
        nranks = o.nranks
        ops = case.get("ops") or [case["op"]]
        halos = case.get("halos") or {}  # axis -> halo extents of that pencil
        pads = case.get("pads") or {}
        inplace = not case.get("out_of_place", False)
        if case.get("force_staged"):
            cd.check(cd.set_tuning(handle, gd, 0, True))
        if case.get("grid_ctas"):
            cd.check(cd.set_tuning(handle, gd, case["grid_ctas"], bool(case.get("force_staged"))))
        if case.get("pipeline_chunks"):
            cd.check(cd.set_pipeline_chunks(handle, gd, case["pipeline_chunks"]))
        if case.get("kernel_variant"):
            cd.check(cd.set_kernel_variant(handle, gd, case["kernel_variant"]))
        if case.get("staged_mode") is not None or case.get("fused_lag"):
            cd.check(cd.set_staged_mode(handle, gd, case.get("staged_mode") or 0, case.get("fused_lag") or 0))
        if case.get("pull"):
            cd.check(cd.set_transfer_mode(handle, gd, 1))
        if case.get("tile_bytes") or case.get("peer_order") or case.get("balance_grid"):
            cd.check(cd.set_schedule(handle, gd, case.get("tile_bytes", 0), case.get("peer_order", 0),
                                     bool(case.get("balance_grid"))))

        def h(ax):
            return halos.get(str(ax))

        def pd(ax):
            return pads.get(str(ax))

        # pencil infos from the library must agree with the oracle's
        infos = {}
        for ax in range(3):
            r_, p = cd.cudecompGetPencilInfo(handle, gd, ax, h(ax), pd(ax))
            cd.check(r_, "cudecompGetPencilInfo")
            mine = pinfo_to_py(p)
            ref = o.pencil_info(rank, ax, h(ax), pd(ax))
            if mine.as_tuple() != ref.as_tuple():
                return dict(ok=False, msg="pencil info mismatch axis %d: %r vs oracle %r" % (ax, mine, ref))
            infos[ax] = mine
        res, wsize = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
        cd.check(res)
        if wsize != o.transpose_workspace_size():
            return dict(ok=False, msg="workspace size mismatch %d vs %d" % (wsize, o.transpose_workspace_size()))
        if case.get("work_alloc", "cudecomp") == "cudecomp":
            res, work_ptr = cd.cudecompMalloc(handle, gd, wsize * es)
            cd.check(res, "cudecompMalloc")
            work = work_ptr
        else:
            work_t = gpu.empty(wsize * es)
            work = work_t.data_ptr()
        data_elems = max(infos[ax].size for ax in range(3))

        if case.get("expect", 0) != 0:
            # error path: the call must fail before touching the buffers (reference api_tests.cc:1493-1505)
            a, b = orc.transpose_axes(ops[0])
            res = cd.TRANSPOSES[ops[0]](handle, gd, 4096, 8192 if not inplace else 4096, work, dt_enum, h(a), h(b),
                                        pd(a), pd(b), None)
            ok = res == case["expect"]
            return dict(ok=ok, msg="returned %d, expected %d" % (res, case["expect"]), paths=[])

        for fill in case.get("fills", ["pattern", "random"]):
            a0 = orc.transpose_axes(ops[0])[0]
            # all ranks' inputs (the oracle runs every rank in this process)
            cur = []
            for r in range(nranks):
                pa = o.pencil_info(r, a0, h(a0), pd(a0))
                buf = np.full(max(o.pencil_info(r, ax, h(ax), pd(ax)).size for ax in range(3)), -7, dtype=np_dtype)
                if fill == "pattern":
                    buf[:pa.size] = orc.pattern_pencil(pa, case["gdims"], np_dtype)
                else:
                    buf[:pa.size] = seeded(pa.size, np_dtype, 1000 * r + 17)
                cur.append(buf)
            other = [np.full(c.size, -3, dtype=np_dtype) for c in cur]  # host mirror of the second device buffer
            d_in = gpu.upload(cur[rank])
            d_out = d_in if inplace else gpu.upload(other[rank])
            for op in ops:
                a, b = orc.transpose_axes(op)
                # oracle step on the host mirrors: same inputs, same pre-existing output contents as on the device
                o.transpose(op, cur, cur if inplace else other, h(a), h(b), pd(a), pd(b))
                exp = (cur if inplace else other)[rank]
                res = cd.TRANSPOSES[op](handle, gd, d_in, d_out, work, dt_enum, h(a), h(b), pd(a), pd(b), None)
                if res != case.get("expect", 0):
                    return dict(ok=False, msg="%s returned %d, expected %d" % (op, res, case.get("expect", 0)))
                if res != 0:
                    return dict(ok=True, msg="failed as expected")
                out["paths"].append(cd.last_path(handle, gd))
                got = gpu.download(d_out, np_dtype)
                pb = infos[b]
                if fill == "pattern":
                    want = orc.pattern_pencil(pb, case["gdims"], np_dtype)
                    if not orc.interior_equal(pb, want, got[:pb.size]):
                        return dict(ok=False, msg="%s %s: interior differs from the analytic pattern" % (op, fill))
                # the whole buffer (interior, untouched halo/padding cells, tail) must equal the oracle's
                if exp.size != got.size or not np.array_equal(exp.view(np.uint8), got.view(np.uint8)):
                    bad = np.flatnonzero(exp != got[:exp.size])
                    return dict(ok=False, msg="%s %s: differs from oracle at %d cells, first %d (want %r got %r)" %
                                (op, fill, bad.size, bad[0] if bad.size else -1, exp[bad[0]] if bad.size else None,
                                 got[bad[0]] if bad.size else None))
                if cd.check_errors(handle, gd) != 0:
                    return dict(ok=False, msg="%s: device-side handshake error" % op)
                if not inplace:
                    # the reference's legacy chain swaps the buffers (tests/cc/transpose_test.cc:523-545)
                    d_in, d_out = d_out, d_in
                    cur, other = other, cur
    finally:
        if work_ptr:
            cd.cudecompFree(handle, gd, work_ptr)
        cd.cudecompGridDescDestroy(handle, gd)
    return out


def stress_case(gpu, handle, rank, case):
    """`reps` X->Y->Z->Y->X round trips enqueued back to back with NO host synchronisation in between (what a solver's
    time loop does): consecutive operations overlap on the device as far as their handshakes allow, so a missing
    dependency between them shows up here and not in the per-operation cases. The final pencil must equal the input
    bit for bit (no halos, no padding: the round trip is the identity), and one extra X->Y must equal the oracle's."""
    dt_enum, np_dtype = DTYPES[case.get("dtype", "double")]
    es = np.dtype(np_dtype).itemsize
    cfg = make_config(case)
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    out = dict(ok=True, msg="", paths=[])
    work_ptr = 0
    try:
        o = make_oracle(case)
        inplace = not case.get("out_of_place", False)
        if case.get("force_staged"):
            cd.check(cd.set_tuning(handle, gd, 0, True))
        if case.get("pipeline_chunks"):
            cd.check(cd.set_pipeline_chunks(handle, gd, case["pipeline_chunks"]))
        if case.get("staged_mode") is not None or case.get("fused_lag"):
            cd.check(cd.set_staged_mode(handle, gd, case.get("staged_mode") or 0, case.get("fused_lag") or 0))
        res, wsize = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
        res, work_ptr = cd.cudecompMalloc(handle, gd, wsize * es)
        cd.check(res, "cudecompMalloc")
        n = max(o.pencil_info(rank, ax).size for ax in range(3))
        hosts = [seeded(max(o.pencil_info(r, ax).size for ax in range(3)), np_dtype, 31 * r + 5) for r in range(o.nranks)]
        d_a = gpu.upload(hosts[rank])
        d_b = d_a if inplace else gpu.empty(n * es)
        for _ in range(case.get("reps", 10)):
            x, y = d_a, d_b
            for op in OPS:
                cd.check(cd.TRANSPOSES[op](handle, gd, x, y, work_ptr, dt_enum, None, None, None, None, None), op)
                out["paths"].append(cd.last_path(handle, gd))
                if not inplace:
                    x, y = y, x
        got = gpu.download(d_a, np_dtype)
        if cd.check_errors(handle, gd) != 0:
            return dict(ok=False, msg="device-side handshake error")
        nx = o.pencil_info(rank, 0).size
        if not np.array_equal(got[:nx].view(np.uint8), hosts[rank][:nx].view(np.uint8)):
            bad = np.flatnonzero(got[:nx] != hosts[rank][:nx])
            return dict(ok=False, msg="round trips are not the identity: %d cells differ, first %d" % (bad.size, bad[0]))
        others = [np.zeros_like(hh) for hh in hosts]
        o.transpose("XY", hosts, hosts if inplace else others)
        cd.check(cd.TRANSPOSES["XY"](handle, gd, d_a, d_b, work_ptr, dt_enum, None, None, None, None, None), "XY")
        got = gpu.download(d_b, np_dtype)
        ny = o.pencil_info(rank, 1).size
        want = (hosts if inplace else others)[rank]
        if not np.array_equal(got[:ny].view(np.uint8), want[:ny].view(np.uint8)):
            return dict(ok=False, msg="X->Y after the round trips differs from the oracle")
    finally:
        if work_ptr:
            cd.cudecompFree(handle, gd, work_ptr)
        cd.cudecompGridDescDestroy(handle, gd)
    return out


def halo_case(gpu, handle, rank, case):
    dt_enum, np_dtype = DTYPES[case.get("dtype", "float")]
    es = np.dtype(np_dtype).itemsize
    cfg = make_config(case)
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    out = dict(ok=True, msg="", paths=[])
    work_ptr = 0
    try:
        o = make_oracle(case)
        ax = case["axis"]
        halo = case["halo"]
        periods = case.get("periods") or [False] * 3
        pad = case.get("padding")
        if case.get("force_staged"):
            cd.check(cd.set_tuning(handle, gd, 0, True))
        r_, p = cd.cudecompGetPencilInfo(handle, gd, ax, halo, pad)
        cd.check(r_)
        pinfo = pinfo_to_py(p)
        if pinfo.as_tuple() != o.pencil_info(rank, ax, halo, pad).as_tuple():
            return dict(ok=False, msg="pencil info mismatch")
        res, wsize = cd.cudecompGetHaloWorkspaceSize(handle, gd, ax, halo)
        cd.check(res)
        if wsize != o.halo_workspace_size(rank, ax, halo):
            return dict(ok=False, msg="halo workspace size mismatch %d vs %d" % (wsize, o.halo_workspace_size(rank, ax, halo)))
        res, work_ptr = cd.cudecompMalloc(handle, gd, max(wsize, 64) * es)
        cd.check(res)
        if case.get("expect", 0) != 0:
            d = gpu.empty(pinfo.size * es)
            res = 0
            for dim in case.get("dims", [0, 1, 2]):
                res = cd.UPDATE_HALOS[ax](handle, gd, d, work_ptr, dt_enum, halo, periods, dim, pad, None)
                if res != 0:
                    break
            return dict(ok=(res == case["expect"]), msg="returned %d, expected %d" % (res, case["expect"]), paths=[])

        for fill in case.get("fills", ["pattern", "random"]):
            hosts = []
            for r in range(o.nranks):
                pr = o.pencil_info(r, ax, halo, pad)
                hosts.append(orc.pattern_pencil(pr, case["gdims"], np_dtype) if fill == "pattern" else
                             seeded(pr.size, np_dtype, 77 + r))
            d = gpu.upload(hosts[rank])
            for dim in case.get("dims", [0, 1, 2]):
                o.halo(ax, dim, hosts, halo, periods, pad)
                res = cd.UPDATE_HALOS[ax](handle, gd, d, work_ptr, dt_enum, halo, periods, dim, pad, None)
                if res != case.get("expect", 0):
                    return dict(ok=False, msg="UpdateHalos dim %d returned %d" % (dim, res))
                if res != 0:
                    return dict(ok=True, msg="failed as expected")
                out["paths"].append(cd.last_path(handle, gd))
            got = gpu.download(d, np_dtype)[:pinfo.size]
            if fill == "pattern" and case.get("dims", [0, 1, 2]) == [0, 1, 2]:
                want = orc.halo_reference(pinfo, case["gdims"], np_dtype, periods)
                if not np.array_equal(want, got):
                    bad = np.flatnonzero(want != got)
                    return dict(ok=False, msg="halo: differs from analytic reference at %d cells (first %d)" %
                                (bad.size, bad[0]))
            if not np.array_equal(hosts[rank].view(np.uint8), got.view(np.uint8)):
                bad = np.flatnonzero(hosts[rank] != got)
                return dict(ok=False, msg="halo %s: differs from oracle at %d cells (first %d)" % (fill, bad.size, bad[0]))
            if cd.check_errors(handle, gd) != 0:
                return dict(ok=False, msg="device-side handshake error")
    finally:
        if work_ptr:
            cd.cudecompFree(handle, gd, work_ptr)
        cd.cudecompGridDescDestroy(handle, gd)
    return out


def autotune_case(gpu, handle, rank, case):
    cfg = make_config(case)
    cfg.pdims[:] = [0, 0]
    opt = cd.cudecompGridDescAutotuneOptions_t()
    cd.check(cd.cudecompGridDescAutotuneOptionsSetDefaults(opt))
    opt.dtype = DTYPES[case.get("dtype", "double")][0]
    opt.n_warmup_trials = case.get("n_warmup", 1)
    opt.n_trials = case.get("n_trials", 2)
    opt.autotune_transpose_backend = bool(case.get("autotune_backend", False))
    opt.grid_mode = case.get("grid_mode", 0)
    if case.get("halo"):
        opt.halo_extents[:] = case["halo"]
        opt.autotune_halo_backend = bool(case.get("autotune_halo_backend", False))
        for i in range(3):
            opt.halo_periods[i] = True
    res, gd = cd.cudecompGridDescCreate(handle, cfg, opt)
    if res != 0:
        return dict(ok=False, msg="autotuned GridDescCreate returned %d" % res)
    try:
        nranks = cd.MPI_Comm_size()
        ok = cfg.pdims[0] * cfg.pdims[1] == nranks and 1 <= cfg.transpose_comm_backend <= 8
        # the chosen grid must work
        sub = dict(case)
        sub["pdims"] = [cfg.pdims[0], cfg.pdims[1]]
        return dict(ok=ok, msg="selected %dx%d backend %d" % (cfg.pdims[0], cfg.pdims[1], cfg.transpose_comm_backend),
                    pdims=[cfg.pdims[0], cfg.pdims[1]], backend=int(cfg.transpose_comm_backend))
    finally:
        cd.cudecompGridDescDestroy(handle, gd)


# ------------------------------------------------------------------------------------------------ plan mode
def plan_case(handle, rank, case):
    cfg = make_config(case)
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    try:
        out = dict(ok=True, pencils={}, transposes={}, halos={}, shifted=[], config_pdims=list(cfg.pdims))
        halos = case.get("halos") or {}
        pads = case.get("pads") or {}
        for ax in range(3):
            r_, p = cd.cudecompGetPencilInfo(handle, gd, ax, halos.get(str(ax)), pads.get(str(ax)))
            cd.check(r_)
            out["pencils"][str(ax)] = [list(p.shape), list(p.lo), list(p.hi), list(p.order), list(p.halo_extents),
                                       list(p.padding), p.size]
        out["workspace"] = cd.cudecompGetTransposeWorkspaceSize(handle, gd)[1]
        for op in OPS:
            ax, d = orc.TRANSPOSE_OPS[op]
            a, b = orc.transpose_axes(op)
            for staged in (False, True):
                out["transposes"]["%s/%d" % (op, staged)] = cd.describe_transpose_boxes(
                    handle, gd, ax, d, halos.get(str(a)), halos.get(str(b)), pads.get(str(a)), pads.get(str(b)), staged)
        if case.get("halo"):
            out["halo_workspace"] = [cd.cudecompGetHaloWorkspaceSize(handle, gd, ax, case["halo"])[1] for ax in range(3)]
            for ax in range(3):
                for dim in range(3):
                    for staged in (False, True):
                        out["halos"]["%d/%d/%d" % (ax, dim, staged)] = cd.describe_halo_boxes(
                            handle, gd, ax, dim, case["halo"], case.get("periods"), case.get("padding"), staged)
        for q in case.get("shifted", []):
            res, v = cd.cudecompGetShiftedRank(handle, gd, q["axis"], q["dim"], q["displacement"], q["periodic"])
            out["shifted"].append(v)
        return out
    finally:
        cd.cudecompGridDescDestroy(handle, gd)


def shim_battery(rank, nranks):
    """Exercises the MPI subset of include/mpi_shim/mpi.h through the exported C symbols."""
    L = cd.lib
    MPI_INT, MPI_DOUBLE, MPI_FLOAT, MPI_INT64 = (1 << 8) | 4, (3 << 8) | 8, (3 << 8) | 4, (1 << 8) | 8
    SUM, MAX, MIN, LOR = 1, 2, 3, 4
    W = cd.MPI_COMM_WORLD
    IN_PLACE = ctypes.c_void_p(-1)
    out = {}
    v = (ctypes.c_int * 2)(rank + 1, 10 * rank)
    r = (ctypes.c_int * 2)()
    assert L.MPI_Allreduce(v, r, 2, MPI_INT, SUM, W) == 0
    out["allreduce_int_sum"] = list(r)
    d = (ctypes.c_double * 1)(1.5 * rank)
    assert L.MPI_Allreduce(IN_PLACE, d, 1, MPI_DOUBLE, MAX, W) == 0
    out["allreduce_double_max_inplace"] = d[0]
    f = (ctypes.c_float * 1)(float(rank) - 2.0)
    assert L.MPI_Allreduce(IN_PLACE, f, 1, MPI_FLOAT, MIN, W) == 0
    out["allreduce_float_min"] = f[0]
    flag = (ctypes.c_int * 1)(1 if rank == nranks - 1 else 0)
    assert L.MPI_Allreduce(IN_PLACE, flag, 1, MPI_INT, LOR, W) == 0
    out["allreduce_lor"] = flag[0]
    b = (ctypes.c_int64 * 3)(*([7, 8, 9] if rank == 1 % nranks else [0, 0, 0]))
    assert L.MPI_Bcast(b, 3, MPI_INT64, 1 % nranks, W) == 0
    out["bcast"] = list(b)
    g = (ctypes.c_int * nranks)()
    mine = (ctypes.c_int * 1)(100 + rank)
    assert L.MPI_Allgather(mine, 1, MPI_INT, g, 1, MPI_INT, W) == 0
    out["allgather"] = list(g)
    L.MPI_Gather.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, ctypes.c_int]
    g2 = (ctypes.c_int * nranks)()
    assert L.MPI_Gather(mine, 1, MPI_INT, g2, 1, MPI_INT, 0, W) == 0
    out["gather_root0"] = list(g2) if rank == 0 else None
    L.MPI_Reduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int]
    res = (ctypes.c_int * 1)(rank * rank)
    if rank == 0:  # the pattern of the reference's test drivers (tests/cc/transpose_test.cc:617)
        assert L.MPI_Reduce(IN_PLACE, res, 1, MPI_INT, MAX, 0, W) == 0
    else:
        assert L.MPI_Reduce(res, res, 1, MPI_INT, MAX, 0, W) == 0
    out["reduce_max_root0"] = res[0] if rank == 0 else None
    # sub-communicators: even / odd ranks, reversed order inside
    sub = cd.MPI_Comm_split(W, rank % 2, -rank)
    out["split_rank"], out["split_size"] = cd.MPI_Comm_rank(sub), cd.MPI_Comm_size(sub)
    s = (ctypes.c_int * 1)(rank)
    assert L.MPI_Allreduce(IN_PLACE, s, 1, MPI_INT, SUM, sub) == 0
    out["split_sum"] = s[0]
    L.MPI_Comm_dup.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    dup = ctypes.c_int(0)
    assert L.MPI_Comm_dup(sub, ctypes.byref(dup)) == 0
    out["dup_size"] = cd.MPI_Comm_size(dup.value)
    assert cd.MPI_Barrier(dup.value) == 0
    c1, c2 = ctypes.c_int(sub), ctypes.c_int(dup.value)
    L.MPI_Comm_free(ctypes.byref(c1))
    L.MPI_Comm_free(ctypes.byref(c2))
    out["freed"] = [c1.value, c2.value]
    # a cuDecomp handle on a sub-communicator (the reference's tests do this, mpi_test_utils.cc:56-66)
    sub2 = cd.MPI_Comm_split(W, 0 if rank < 2 else 1, rank)
    res_, h = cd.cudecompInit(sub2)
    cfg = cd.cudecompGridDescConfig_t()
    cd.cudecompGridDescConfigSetDefaults(cfg)
    cfg.gdims[:] = [8, 8, 8]
    n_sub = cd.MPI_Comm_size(sub2)
    cfg.pdims[:] = [1, n_sub]
    res2, gd = cd.cudecompGridDescCreate(h, cfg)
    out["subcomm_handle"] = [res_, res2]
    if res2 == 0:
        _, p = cd.cudecompGetPencilInfo(h, gd, 0)
        out["subcomm_shape"] = list(p.shape)
        cd.cudecompGridDescDestroy(h, gd)
    cd.cudecompFinalize(h)
    out["wtime_positive"] = cd.lib.MPI_Wtime() > 0
    # wider coverage of the typed reductions and the in-place conventions
    MPI_UNSIGNED, MPI_INT64_T = (2 << 8) | 4, (1 << 8) | 8
    big = (ctypes.c_int64 * 5)(*[(1 << 40) + 1000 * rank + k for k in range(5)])
    bigr = (ctypes.c_int64 * 5)()
    assert L.MPI_Allreduce(big, bigr, 5, MPI_INT64_T, SUM, W) == 0
    out["allreduce_int64_sum"] = list(bigr)
    u = (ctypes.c_uint32 * 1)(1 if rank < 2 else 0x80000000 + rank)
    assert L.MPI_Allreduce(IN_PLACE, u, 1, MPI_UNSIGNED, MAX, W) == 0
    out["allreduce_unsigned_max"] = u[0]
    neg = (ctypes.c_int * 1)(rank - 5)
    assert L.MPI_Allreduce(IN_PLACE, neg, 1, MPI_INT, MIN, W) == 0
    out["allreduce_int_min_negative"] = neg[0]
    ag = (ctypes.c_int64 * (2 * nranks))()
    ag[2 * rank], ag[2 * rank + 1] = (1 << 33) + rank, -rank
    assert L.MPI_Allgather(IN_PLACE, 0, 0, ag, 2, MPI_INT64_T, W) == 0
    out["allgather_inplace"] = list(ag)
    selfsum = (ctypes.c_double * 3)(1.25, 2.5, float(rank))
    selfout = (ctypes.c_double * 3)()
    assert L.MPI_Allreduce(selfsum, selfout, 3, MPI_DOUBLE, SUM, 2) == 0  # MPI_COMM_SELF: one rank, not in place
    out["allreduce_self"] = list(selfout)
    L.MPI_Comm_free.argtypes = [ctypes.c_void_p]
    out["comm_free_null"] = L.MPI_Comm_free(None)
    return out


_restated = None


def nccl_crosscheck_case(gpu, handle, rank, case):
    """The library against the reference's NCCL arm restated with torch ops + NCCL all-to-all (bench/nccl_restated.py):
    same seeded input, every op of the chain compared byte for byte. Needs one GPU per rank."""
    global _restated
    import importlib.util
    torch = gpu.torch
    import torch.distributed as dist
    nranks = int(os.environ["WORLD_SIZE"])
    if torch.cuda.device_count() < nranks:
        return dict(ok=True, skipped=True, msg="needs %d GPUs (NCCL cannot share a device between ranks)" % nranks)
    if _restated is None:
        spec = importlib.util.spec_from_file_location("nccl_restated", os.path.join(ROOT, "bench", "nccl_restated.py"))
        _restated = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_restated)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=nranks, device_id=torch.device("cuda", gpu.dev))
    dt_enum, np_dtype = DTYPES[case.get("dtype", "double")]
    es = np.dtype(np_dtype).itemsize
    cfg = make_config(case)
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    out = dict(ok=True, msg="", paths=[])
    work_ptr = 0
    try:
        geom = _restated.Geometry(case["gdims"], case["pdims"], case.get("axis_contiguous") or [False] * 3,
                                  case.get("mem_order"), case.get("gdims_dist"))
        rt = _restated.RestatedTranspose(geom, rank)
        rt.create_groups()
        sizes = [cd.cudecompGetPencilInfo(handle, gd, ax)[1].size for ax in range(3)]
        for ax in range(3):
            if sizes[ax] != _restated._prod(geom.torch_shape(rank, ax)):
                return dict(ok=False, msg="pencil size mismatch on axis %d" % ax)
        n = max(sizes)
        res, wsize = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
        cd.check(res)
        res, work_ptr = cd.cudecompMalloc(handle, gd, wsize * es)
        cd.check(res, "cudecompMalloc")
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64,
               np.complex128: torch.complex128}[np_dtype]
        host = np.zeros(n, np_dtype)
        host[:sizes[0]] = seeded(sizes[0], np_dtype, 4242 + rank)
        mine_in = torch.from_numpy(host.copy()).cuda(gpu.dev)
        ref_in = mine_in.clone()
        mine_out, ref_out = torch.zeros_like(mine_in), torch.zeros_like(mine_in)
        send, recv = torch.zeros(n, dtype=tdt, device=mine_in.device), torch.zeros(n, dtype=tdt, device=mine_in.device)
        for op in case.get("ops", OPS):
            b = orc.transpose_axes(op)[1]
            res = cd.TRANSPOSES[op](handle, gd, mine_in, mine_out, work_ptr, dt_enum, None, None, None, None, None)
            if res != 0:
                return dict(ok=False, msg="%s returned %d" % (op, res))
            rt.transpose(op, ref_in, ref_out, send, recv)
            torch.cuda.synchronize()
            out["paths"].append(cd.last_path(handle, gd))
            if not torch.equal(mine_out[:sizes[b]].view(torch.uint8), ref_out[:sizes[b]].view(torch.uint8)):
                return dict(ok=False, msg="%s: differs from the restated NCCL arm" % op)
            mine_in, mine_out = mine_out, mine_in
            ref_in, ref_out = ref_out, ref_in
    finally:
        if work_ptr:
            cd.cudecompFree(handle, gd, work_ptr)
        cd.cudecompGridDescDestroy(handle, gd)
    return out


def mailbox_selftest(handle, case):
    return dict(ok=cd.lib.cudecompB200SelfTestMailbox(handle, case.get("iterations", 2000), case.get("seed", 1)) == 0)


def cumem_probe(handle):
    return dict(ok=True, state=cd.cumem_state(handle), fd_passing=cd.probe_fd_passing(handle))


def main():
    payload_path, out_dir = sys.argv[1], sys.argv[2]
    with open(payload_path) as f:
        payload = json.load(f)
    rank = int(os.environ.get("RANK", "0"))
    results = []
    gpu = Gpu() if payload["mode"] in ("gpu", "api_gpu") else None
    assert cd.MPI_Init() == 0
    res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res, "cudecompInit")
    fatal = False
    for case in payload["cases"]:
        if fatal:
            results.append(dict(ok=False, msg="skipped after an earlier failure that may have desynchronised the ranks"))
            continue
        try:
            if payload["mode"] == "shim":
                results.append(shim_battery(rank, int(os.environ["WORLD_SIZE"])))
            elif payload["mode"] == "mailbox":
                results.append(mailbox_selftest(handle, case))
            elif payload["mode"] == "cumem":
                results.append(cumem_probe(handle))
            elif payload["mode"] == "plan":
                results.append(plan_case(handle, rank, case))
            elif payload["mode"] in ("api", "api_gpu"):
                from tests import _api_battery
                results.append(_api_battery.run(handle, rank, case["name"], payload["mode"] == "api_gpu"))
            elif case["kind"] == "halo":
                results.append(halo_case(gpu, handle, rank, case))
            elif case["kind"] == "nccl_crosscheck":
                results.append(nccl_crosscheck_case(gpu, handle, rank, case))
            elif case["kind"] == "autotune":
                results.append(autotune_case(gpu, handle, rank, case))
            elif case["kind"] == "stress":
                results.append(stress_case(gpu, handle, rank, case))
            else:
                results.append(transpose_case(gpu, handle, rank, case))
        except Exception as e:  # noqa: BLE001
            results.append(dict(ok=False, msg="exception: %s\n%s" % (e, traceback.format_exc())))
            fatal = True
        # keep ranks in lock step between cases (a failing rank must not leave the others mid-collective)
        cd.MPI_Barrier()
    if os.environ.get("CDB_TEST_GLOO") == "1":
        # N>1 host path cross-checked over torch.distributed (gloo): every rank sees every rank's plans
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=int(os.environ["WORLD_SIZE"]))
        results = json.loads(json.dumps(results))  # tuples -> lists, as the parent will read them
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, results)
        assert gathered[rank] == results
        results = dict(mine=results, gathered=gathered)
        dist.barrier()
        dist.destroy_process_group()
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(results, f)
    if payload["mode"] == "gpu":
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    cd.cudecompFinalize(handle)
    cd.MPI_Finalize()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark for the pencil-transpose hot path on B200.

Workload (BASELINE.json metric): 1024^3 complex128 grid, X->Y->Z->Y->X transpose round trip (the autotuner's trial
body, reference src/autotune.cc:541-576), default memory layout, out-of-place buffers (the autotuner's default,
transpose_use_inplace_buffers = false), process grid 1x1 / 1x2 / 2x2 / 2x4 at 1 / 2 / 4 / 8 GPUs.
A step is one round trip (4 transposes) over this rank's pencil. The global grid is fixed, so scaling is strong.

metric  = effective transpose GB/s, the reference's own accounting (include/internal/transpose.h:316,
          src/performance.cc:391): pencil bytes S per transpose / time. value = whole-job aggregate = N * 4S / t_step.
e2e     = same metric with the pencil starting and ending in pinned HOST memory: H2D of the input pencil, the round
          trip through the C ABI, D2H of the result pencil, all inside the timed region.
roofline= dominant kernel (cdb::rowCopyKernel<uint4>) against the measured HBM copy bandwidth: algorithmic bytes
          2S per launch (read the pencil once, write it once) / average launch duration from CUDA events.
nvlink  = bytes that leave this GPU per second against the measured 770 GB/s peer-copy rate (N > 1).

`--impl reference` times the CPU restatement of the reference's path (oracle/, OpenMP over all host cores) on a
bounded z-slab sample of the same workload; the reference itself cannot be built here.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_BY_N = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
OPS = ["XY", "YZ", "ZY", "YX"]
NVLINK_PEAK_GBS = 770.0  # measured peer-copy rate per direction (B200_PROFILING.md), 900 nominal
NVLINK_PACKET_FACTOR = 1.275068 / 1.073742  # bytes on the link per byte of payload for SM stores (ncu, profiles/)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=1024, help="grid edge (default 1024, the metric's configuration)")
    ap.add_argument("--dtype", default="double_complex")
    ap.add_argument("--inplace", action="store_true", help="in-place buffers (the reference benchmark's default)")
    ap.add_argument("--axis-contiguous", action="store_true")
    ap.add_argument("--pdims", default=None, help="override process grid, e.g. 2x4")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--staged", action="store_true", help="force the workspace-staged schedule")
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--bulk", action="store_true", help="TMA bulk row-copy variant (experimental)")
    ap.add_argument("--wide", action="store_true", help="256-bit LDG/STG row-copy variant (experimental)")
    ap.add_argument("--pull", action="store_true", help="receiver-driven direct transposes (experimental)")
    ap.add_argument("--chunks", type=int, default=0, help="chunks of staged (in-place) transposes, 0 = from the pencil size")
    ap.add_argument("--staged-mode", type=int, default=0, help="0 one phased launch per staged transpose, 1 separate launches")
    ap.add_argument("--lag", type=int, default=0, help="phases between a chunk's push and its unpack (0 = library default)")
    ap.add_argument("--kernel-variant", type=int, default=0, help="cudecompB200SetKernelVariant value (3 = element-wise transpose)")
    ap.add_argument("--no-wire-wide", action="store_true", help="128-bit accesses also in launches that store into peers")
    ap.add_argument("--no-column-chunks", action="store_true", help="fused staged schedule: plane chunks only")
    ap.add_argument("--phase-head", type=int, default=-1, help="percent of a step's pushes ahead of the unpacks (fused staged)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed integer-pattern check after the timed region")
    ap.add_argument("--tile-bytes", type=int, default=0, help="row-copy tile size (0 = 32 KiB)")
    ap.add_argument("--peer-order", type=int, default=0, help="0 one-shot interleaved, 1 pairwise rounds")
    ap.add_argument("--balance-grid", type=int, default=0, help="1: CTA count with the fullest last grid-stride round")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """CPU restatement of the reference path on the host cores, bounded sample: a z-slab of the same grid."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    import torch  # noqa: F401  (only for parity of the environment; not used by the timed code)

    n = args.n
    pd = GRID_BY_N.get(args.gpus, (1, args.gpus))
    if args.pdims:
        pd = tuple(int(v) for v in args.pdims.split("x"))
    # whole grid when the host can hold it and `steps + warmup` round trips end within ~3 minutes, else a z-slab
    nz = cpu_sample_nz(args, pd, budget_s=180.0, steps=args.steps + args.warmup)
    gd = [n, n, nz]
    o = orc.Oracle(gd, pd, (args.axis_contiguous,) * 3)
    # every core this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the
    # reference arm single-threaded at N > 1 although the other ranks exit immediately
    cores = len(os.sched_getaffinity(0))
    o.set_threads(cores)
    dt = orc.NP_DTYPES[args.dtype]
    es = np.dtype(dt).itemsize
    bufs_a = [np.ones(max(o.pencil_info(r, ax).size for ax in range(3)), dt) for r in range(o.nranks)]
    bufs_b = [np.zeros_like(b) for b in bufs_a]
    S_total = float(n) * n * nz * es

    def step():
        cur, other = bufs_a, bufs_b
        for op in OPS:
            o.transpose(op, cur, cur if args.inplace else other)
            if not args.inplace:
                cur, other = other, cur

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt_s = (time.perf_counter() - t0) / args.steps
    value = 4.0 * S_total / dt_s / 1e9
    sample = "%s %dx%dx%d of the %d^3 %s grid, %s, pdims %dx%d as %d in-process ranks" % (
        "the whole grid" if nz == n else "z-slab", n, n, nz, n, args.dtype,
        "in-place" if args.inplace else "out-of-place", pd[0], pd[1], o.nranks)
    line = {
        "impl": "reference", "metric": "effective transpose GB/s (4*S/t round trip, whole job)", "value": value,
        "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_s * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128" if es == 16 else args.dtype,
        "data": "synthetic",
        # the same `config` object as the native arm prints when the whole grid was timed; a z-slab says so
        "config": workload_config(args, pd, sample=None if nz == n else sample),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def index_pattern(torch, pinfo, gdims, es, device):
    """The reference tests' known answer for a pencil WITHOUT halos or padding (tests/ctest/transpose_tests.cc:333-354):
    every element carries its global linear index, here as an integer bit pattern so that no float rounding is
    involved: int32 for 4-byte elements (index mod 2^32), int64 for 8-byte ones, the pair (index, ~index) for 16-byte
    ones. `pinfo` is a cudecompPencilInfo_t (shape / lo / order by memory position). Flat tensor in memory order."""
    gstride = [1, gdims[0], gdims[0] * gdims[1]]
    idt = torch.int32 if es == 4 else torch.int64
    terms = []
    for k in range(3):
        v = (torch.arange(pinfo.shape[k], device=device, dtype=torch.int64) + pinfo.lo[k]) * gstride[pinfo.order[k]]
        terms.append(v.to(idt))
    g = (terms[2][:, None, None] + terms[1][None, :, None] + terms[0][None, None, :]).reshape(-1)
    if es == 16:
        return torch.stack([g, ~g], dim=1).reshape(-1)
    return g


def mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def calibrate_cpu_gbs(args, pd):
    """Effective GB/s of the oracle on THIS host, from one timed round trip on a thin z-slab (well under a second): what
    the time budget of the CPU legs is planned with. Falls back to 15 GB/s (a 16-core host) if the probe fails."""
    try:
        import numpy as np
        from oracle import oracle as orc
        n = args.n
        nz = min(n, max(pd[1] * 4, 16))
        o = orc.Oracle([n, n, nz], pd, (args.axis_contiguous,) * 3)
        o.set_threads(len(os.sched_getaffinity(0)))
        dt = orc.NP_DTYPES[args.dtype]
        a = [np.ones(max(o.pencil_info(r, ax).size for ax in range(3)), dt) for r in range(o.nranks)]
        b = [np.zeros_like(x) for x in a]
        t = None
        for _ in range(2):  # the first pass also allocates the oracle's staging regions
            t0 = time.perf_counter()
            cur, other = a, b
            for op in OPS:
                o.transpose(op, cur, cur if args.inplace else other)
                if not args.inplace:
                    cur, other = other, cur
            t = time.perf_counter() - t0
        o.release()
        return max(1.0, 4.0 * n * n * nz * np.dtype(dt).itemsize / t / 1e9)
    except Exception:  # noqa: BLE001
        return 15.0


def cpu_sample_nz(args, pd, budget_s, gbs_guess=None, steps=1):
    """z extent of the CPU sample: the WHOLE grid when host memory (input + output + the oracle's send / receive staging =
    4 grids) and the time budget allow it, else the largest power-of-two z-slab that does. The x-y extent is always
    full (what the X<->Y exchange moves); z only has to hold 4 planes per rank of the row communicator. The time is
    planned with the oracle's measured rate on this host (calibrate_cpu_gbs), less 20 %."""
    import numpy as np
    from oracle import oracle as orc
    es = np.dtype(orc.NP_DTYPES[args.dtype]).itemsize
    n = args.n
    nz = n
    if gbs_guess is None:
        gbs_guess = 0.8 * calibrate_cpu_gbs(args, pd)
    avail = mem_available_bytes()
    while nz > max(pd[1] * 4, 8):
        grid_bytes = float(n) * n * nz * es
        fits = avail == 0 or 4.5 * grid_bytes < 0.6 * avail
        quick = 4.0 * grid_bytes / (gbs_guess * 1e9) * steps <= budget_s
        if fits and quick:
            break
        nz //= 2
    return max(nz, pd[1] * 4)


def workload_config(args, pd, sample=None):
    c = {"workload": "%d^3 %s X->Y->Z->Y->X transpose round trip, %s, %s layout, pdims %dx%d" % (
        args.n, args.dtype, "in-place" if args.inplace else "out-of-place",
        "axis-contiguous" if args.axis_contiguous else "default", pd[0], pd[1]),
        "l2": "inputs larger than L2 (no flush needed)", "grid": [args.n] * 3, "pdims": list(pd)}
    if sample:
        c["sample"] = sample
    return c


def bind_near_gpu(torch, local_rank):
    """Pin this process to the cores next to its GPU before any pinned host memory is allocated, so that the staging
    buffers of the end-to-end leg land on the GPU's NUMA node (what the reference's launch scripts do with numactl).
    Returns a short description for the JSON line; does nothing when the topology is not exposed."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node not exposed"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d has no usable cores" % node
        os.sched_setaffinity(0, cpus)
        return "bound to %d cores of numa node %d" % (len(cpus), node)
    except Exception as e:  # noqa: BLE001 -- best effort, never fatal
        return "not bound (%s)" % type(e).__name__


# -------------------------------------------------------------------------------------------------- native arm
def run_native(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cudecomp_b200 import capi as cd

    # The CPU leg runs first, before this process narrows its affinity to the GPU's NUMA node: it gets every core the
    # process was given (threads created later would inherit the narrowed mask).
    cpu_first = None
    if rank == 0 and not args.no_cpu_baseline:
        pd0 = tuple(int(v) for v in args.pdims.split("x")) if args.pdims else GRID_BY_N.get(world, (1, world))
        try:
            cpu_first = cpu_baseline(args, pd0)
        except Exception as e:  # noqa: BLE001 -- a reported baseline, never a reason to lose the bench line
            cpu_first = {"value": None, "unit": "GB/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                         "sample": "failed: %r" % (e,)}

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_near_gpu(torch, local_rank) if not args.no_numa_bind else "disabled"
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pd = GRID_BY_N.get(world, (1, world))
    if args.pdims:
        pd = tuple(int(v) for v in args.pdims.split("x"))
    assert pd[0] * pd[1] == world, "pdims do not match the number of ranks"
    dt_enum = {"float": cd.CUDECOMP_FLOAT, "double": cd.CUDECOMP_DOUBLE, "float_complex": cd.CUDECOMP_FLOAT_COMPLEX,
               "double_complex": cd.CUDECOMP_DOUBLE_COMPLEX}[args.dtype]
    es = cd.DTYPE_SIZES[dt_enum]

    if args.no_wire_wide:
        os.environ["CUDECOMP_B200_WIRE_WIDE"] = "0"
    if args.no_column_chunks:
        os.environ["CUDECOMP_B200_COLUMN_CHUNKS"] = "0"
    if args.phase_head >= 0:
        os.environ["CUDECOMP_B200_PHASE_HEAD"] = str(args.phase_head)
    assert cd.MPI_Init() == 0
    res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res, "cudecompInit")
    cfg = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(cfg))
    cfg.gdims[:] = [args.n] * 3
    cfg.pdims[:] = pd
    cfg.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_NCCL
    for i in range(3):
        cfg.transpose_axis_contiguous[i] = args.axis_contiguous
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    if args.staged or args.ctas:
        cd.check(cd.set_tuning(handle, gd, args.ctas, args.staged))
    if args.bulk:
        cd.check(cd.set_kernel_variant(handle, gd, 1))
    if args.wide:
        cd.check(cd.set_kernel_variant(handle, gd, 2))
    if args.kernel_variant:
        cd.check(cd.set_kernel_variant(handle, gd, args.kernel_variant))
    if args.pull:
        cd.check(cd.set_transfer_mode(handle, gd, 1))
    if args.chunks:
        cd.check(cd.set_pipeline_chunks(handle, gd, args.chunks))
    if args.staged_mode or args.lag:
        cd.check(cd.set_staged_mode(handle, gd, args.staged_mode, args.lag))
    if args.tile_bytes or args.peer_order or args.balance_grid:
        cd.check(cd.set_schedule(handle, gd, args.tile_bytes, args.peer_order, bool(args.balance_grid)))

    sizes = [cd.cudecompGetPencilInfo(handle, gd, ax)[1].size for ax in range(3)]
    S = sizes[0] * es  # bytes of this rank's pencil (equal for the three orientations on even grids)
    nbytes = max(sizes) * es
    res, wsize = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
    res, work = cd.cudecompMalloc(handle, gd, wsize * es)
    cd.check(res, "cudecompMalloc")

    # synthetic pencil: uniform [0,1) reals like the reference benchmark (benchmark/benchmark.cu:472-485)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    a = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    a.uniform_(0.0, 1.0, generator=gen)
    b = a if args.inplace else torch.zeros_like(a)
    stream = torch.cuda.current_stream()

    def round_trip(src, dst, events=None):
        cur, other = src, dst
        for k, op in enumerate(OPS):
            cd.check(cd.TRANSPOSES[op](handle, gd, cur, cur if args.inplace else other, work, dt_enum, None, None,
                                       None, None, stream), op)
            if events is not None:
                events[k + 1].record(stream)
            if not args.inplace:
                cur, other = other, cur
        return cur

    # ---- device-resident timing
    for _ in range(args.warmup):
        round_trip(a, b)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches0 = cd.launch_count()
    barrier()
    t_host0 = time.perf_counter()
    for s in range(args.steps):
        evs[s][0].record(stream)
        round_trip(a, b, evs[s])
    host_us_per_op = (time.perf_counter() - t_host0) / (4 * args.steps) * 1e6  # enqueue cost incl. Python
    torch.cuda.synchronize()
    barrier()
    launches = cd.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    cd.check(cd.check_errors(handle, gd), "device-side handshake")
    total_ms = sum(evs[s][0].elapsed_time(evs[s][4]) for s in range(args.steps))
    op_ms = [sum(evs[s][k].elapsed_time(evs[s][k + 1]) for s in range(args.steps)) / args.steps for k in range(4)]
    ms_step = max_over_ranks(total_ms / args.steps)
    op_ms = [max_over_ranks(v) for v in op_ms]
    value = world * 4.0 * S / (ms_step * 1e-3) / 1e9

    # ---- parity at size (untimed): the reference tests' known answer (tests/ctest/transpose_tests.cc:333-378) -- every
    # element carries its GLOBAL linear index as an integer bit pattern (16-byte elements: the index and its bitwise
    # complement), built and compared on the device. After each of the four operations this rank's whole output pencil
    # must equal the analytic pattern of that orientation, on every rank.
    parity = None
    if not args.no_parity:
        infos = [cd.cudecompGetPencilInfo(handle, gd, ax)[1] for ax in range(3)]
        idt = torch.int32 if es == 4 else torch.int64

        def expected(ax):
            return index_pattern(torch, infos[ax], [args.n] * 3, es, dev)

        def as_ints(t, nelem):
            return t.view(idt)[: nelem * (2 if es == 16 else 1)]

        pa, pb = (a, b) if not args.inplace else (a, a)
        as_ints(pa, sizes[0]).copy_(expected(0))
        if not args.inplace:
            pb.zero_()
        bad_ops, cur, other = [], pa, pb
        AXIS_AFTER = {"XY": 1, "YZ": 2, "ZY": 1, "YX": 0}
        for op in OPS:
            cd.check(cd.TRANSPOSES[op](handle, gd, cur, cur if args.inplace else other, work, dt_enum, None, None, None,
                                       None, stream), op)
            res_t = cur if args.inplace else other
            ax = AXIS_AFTER[op]
            want = expected(ax)
            if not torch.equal(as_ints(res_t, sizes[ax]), want):
                bad_ops.append(op)
            del want
            if not args.inplace:
                cur, other = other, cur
        torch.cuda.synchronize()
        cd.check(cd.check_errors(handle, gd), "device-side handshake (parity run)")
        ok_all = max_over_ranks(float(len(bad_ops))) == 0.0
        parity = {"checked_ops": 4, "ok": bool(ok_all), "ranks": world,
                  "pattern": "global linear index per element as integer bits, compared on the device on every rank "
                             "after each operation (whole output pencil)", "failed_ops_rank0": bad_ops}
        # restore a benign payload for the end-to-end leg
        a.uniform_(0.0, 1.0, generator=gen)

    # ---- end to end: pinned host pencil -> device -> round trip -> pinned host, every step.
    # Steps are software-pipelined over two device buffer sets and three streams (H2D | transposes | D2H), the way a
    # caller streaming pencils through the library would run it: step i's D2H overlaps step i+1's H2D (PCIe is full
    # duplex). Each step still copies its own input in and its own result out inside the timed region.
    e2e = None
    if not args.no_e2e:
        ne = S // 8
        h_in = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        h_in.copy_(a[:ne])
        h_out = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        slots = [(a, b)]
        try:
            slots.append((torch.empty_like(a), a if args.inplace else torch.empty_like(a)))
            if args.inplace:
                slots[1] = (slots[1][0], slots[1][0])
        except torch.cuda.OutOfMemoryError:
            pass  # single slot: no overlap between consecutive steps
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        nslots = len(slots)

        def e2e_run(nsteps):
            ev_h2d = [torch.cuda.Event() for _ in range(nsteps)]
            ev_cmp = [torch.cuda.Event() for _ in range(nsteps)]
            ev_d2h = [torch.cuda.Event() for _ in range(nsteps)]
            for i in range(nsteps):
                x, y = slots[i % nslots]
                with torch.cuda.stream(s_h2d):
                    if i >= nslots:
                        s_h2d.wait_event(ev_d2h[i - nslots])  # the slot's previous result has been read back
                    x[:ne].copy_(h_in, non_blocking=True)
                    ev_h2d[i].record(s_h2d)
                stream.wait_event(ev_h2d[i])
                out = round_trip(x, y)
                ev_cmp[i].record(stream)
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_cmp[i])
                    h_out.copy_(out[:ne], non_blocking=True)
                    ev_d2h[i].record(s_d2h)
            return ev_d2h[-1]

        e2e_run(min(args.warmup, 2) or 1).synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_h2d)
        last = e2e_run(args.steps)
        s_d2h.wait_event(last)
        e1.record(s_d2h)
        torch.cuda.synchronize()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        e2e = {"value": world * 4.0 * S / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(S),
               "d2h_bytes_per_step": int(S), "ms_per_step": e2e_ms,
               "pipelining": "%d device buffer sets; H2D, transposes and D2H on separate streams" % nslots,
               "host_binding": numa}

        # What the host link gives this leg at best: the same pinned buffers copied with nothing else going on, one
        # direction at a time and both at once (all ranks at the same time, like the leg itself). The end-to-end step
        # cannot be shorter than S / duplex rate; the transposes are the remainder.
        def copy_rate(do_h2d, do_d2h, reps=2):
            x = slots[0][0]
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            s_h2d.wait_event(t0)
            s_d2h.wait_event(t0)
            for _ in range(reps):
                if do_h2d:
                    with torch.cuda.stream(s_h2d):
                        x[:ne].copy_(h_in, non_blocking=True)
                if do_d2h:
                    with torch.cuda.stream(s_d2h):
                        h_out.copy_(slots[-1][1][:ne], non_blocking=True)
            ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event()
            ev_a.record(s_h2d)
            ev_b.record(s_d2h)
            stream.wait_event(ev_a)
            stream.wait_event(ev_b)
            t1.record(stream)
            torch.cuda.synchronize()
            ms = max_over_ranks(t0.elapsed_time(t1) / reps)
            return S / (ms * 1e-3) / 1e9

        try:
            e2e["host_link"] = {"h2d_gbs_per_gpu": copy_rate(True, False), "d2h_gbs_per_gpu": copy_rate(False, True),
                                "duplex_gbs_per_gpu_per_direction": copy_rate(True, True), "unit": "GB/s",
                                "note": "pinned-memory copies of the same pencil, all ranks at once, nothing else running"}
            e2e["host_link"]["floor_ms_per_step"] = S / (e2e["host_link"]["duplex_gbs_per_gpu_per_direction"] * 1e9) * 1e3
        except Exception as e:  # noqa: BLE001 -- the probe is context, never a reason to lose the bench line
            e2e["host_link"] = {"error": repr(e)}
        del h_in, h_out, slots

    hbm_peak, peak_src = peaks()
    # dominant kernel: one copy launch per transpose; algorithmic traffic 2S (read once, write once)
    slowest = max(range(4), key=lambda k: op_ms[k])
    kern_ms = sum(op_ms) / 4.0
    achieved = 2.0 * S / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": measured_traffic(world, args),
                "kernel": dominant_kernel(args, world, es),
                "peak_source": peak_src,
                "per_op_ms": dict(zip(OPS, op_ms)), "algorithmic_bytes_per_launch": 2.0 * S,
                "slowest_op": OPS[slowest]}
    nvlink = None
    if world > 1:
        comm = {"XY": pd[0], "YX": pd[0], "YZ": pd[1], "ZY": pd[1]}
        wire = sum(S * (1.0 - 1.0 / comm[op]) for op in OPS)
        t_wire = sum(op_ms[k] for k, op in enumerate(OPS) if comm[op] > 1) * 1e-3
        n_wire = sum(1 for op in OPS if comm[op] > 1)
        nvlink = {"wire_bytes_per_step": wire, "achieved": wire / t_wire / 1e9 if t_wire > 0 else None,
                  "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                  "frac": (wire / t_wire / 1e9) / NVLINK_PEAK_GBS if t_wire > 0 else None,
                  "roofline_ms_per_step": wire / (NVLINK_PEAK_GBS * 1e9) * 1e3 +
                  sum(2.0 * S / (hbm_peak * 1e9) * 1e3 for op in OPS if comm[op] == 1)}
        if t_wire > 0:
            # ncu with NVLink counters (profiles/r2_rowcopy_peer_full.txt): SM stores leave the GPU as 128-byte packets
            # carrying 24 bytes of protocol each, so the link itself moves 1.1875 x the payload
            nvlink["raw_link_gbs_estimate"] = nvlink["achieved"] * NVLINK_PACKET_FACTOR
            nvlink["raw_link_frac_of_900"] = nvlink["achieved"] * NVLINK_PACKET_FACTOR / 900.0
            nvlink["packet_overhead_source"] = "profiles/r2_rowcopy_peer_full.txt (nvltx__bytes / nvltx__bytes_data_user)"
        if t_wire > 0:
            # With more than one rank in the communicator the dominant launches are bound by NVLink egress, not HBM:
            # algorithmic bytes per launch = the bytes that must leave this GPU, peak = measured peer-copy rate.
            hbm_view = dict(roofline)
            roofline = {"bound": "nvlink", "achieved": nvlink["achieved"], "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                        "frac": nvlink["frac"], "traffic": None, "kernel": roofline["kernel"],
                        "peak_source": "measured peer copy, 770 GB/s per direction (B200_PROFILING.md); 900 nominal",
                        "algorithmic_bytes_per_launch": wire / n_wire, "per_op_ms": dict(zip(OPS, op_ms)),
                        "hbm_view": {k: hbm_view[k] for k in ("achieved", "peak", "frac")}}

    cpu = None
    if cpu_first is not None:
        cpu = cpu_first

    if rank == 0:
        line = {"metric": "effective transpose GB/s (4*S/t round trip, whole job)", "value": value, "unit": "GB/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "c128" if es == 16 else args.dtype, "data": "synthetic",
                "config": workload_config(args, pd), "per_gpu_value": value / world,
                "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
                "host_enqueue_us_per_op": host_us_per_op,
                "path": {0: "none", 1: "local", 2: "direct", 3: "staged"}[cd.last_path(handle, gd)]}
        if parity:
            line["parity"] = parity
        if nvlink:
            line["nvlink"] = nvlink
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    barrier()
    cd.cudecompFree(handle, gd, work)
    cd.cudecompGridDescDestroy(handle, gd)
    cd.cudecompFinalize(handle)
    cd.MPI_Finalize()
    if world > 1:
        dist.destroy_process_group()


def dominant_kernel(args, world, es):
    """Name of the kernel most of the step's device time goes to (csrc/kernels.cu), as the engine selects it."""
    t = {4: "unsigned int", 8: "uint2", 16: "uint4"}[es]
    if args.axis_contiguous:
        return "cdb::transposeKernel<%s>" % t if args.kernel_variant == 3 else "cdb::transposeVecKernel<%s>" % t
    if args.bulk:
        return "cdb::rowCopyBulkKernel"
    wide = args.wide or (world > 1 and not args.no_wire_wide and args.kernel_variant == 0)
    v = "cdb::Vec32" if wide else {4: "unsigned int", 8: "uint2", 16: "uint4"}[min(es, 16)]
    if world > 1 and (args.inplace or args.staged) and args.staged_mode == 0:
        return "cdb::rowCopyPhasedKernel<%s>" % v
    return "cdb::rowCopyKernel<%s>" % (v if es == 16 or wide else "uint4")


def measured_traffic(world, args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu captures:
    profiles/traffic.json lists one entry per captured configuration (kernel, commit, source file). None when this
    configuration has no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entries = json.load(f)["entries"]
    except (OSError, ValueError, KeyError):
        return None
    for e in entries:
        if (e.get("n_gpus") == world and e.get("grid") == args.n and e.get("dtype") == args.dtype and
                bool(e.get("inplace")) == bool(args.inplace) and
                bool(e.get("axis_contiguous")) == bool(args.axis_contiguous)):
            return e.get("dram_bytes_per_launch")
    return None


def cpu_baseline(args, pd):
    """The oracle on this box's host cores, bounded sample of the same workload (see run_reference)."""
    import numpy as np
    from oracle import oracle as orc
    n = args.n
    reps = 2
    nz = cpu_sample_nz(args, pd, budget_s=25.0, steps=reps + 1)  # about 10-30 s of CPU work
    o = orc.Oracle([n, n, nz], pd, (args.axis_contiguous,) * 3)
    o.set_threads(len(os.sched_getaffinity(0)))
    dt = orc.NP_DTYPES[args.dtype]
    es = np.dtype(dt).itemsize
    A = [np.ones(max(o.pencil_info(r, ax).size for ax in range(3)), dt) for r in range(o.nranks)]
    B = [np.zeros_like(x) for x in A]

    def step():
        cur, other = A, B
        for op in OPS:
            o.transpose(op, cur, cur if args.inplace else other)
            if not args.inplace:
                cur, other = other, cur
    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    t = (time.perf_counter() - t0) / reps
    threads = o.max_threads()
    o.release()  # the oracle's staging regions are as large as the grid: give them back before the GPU legs
    return {"value": 4.0 * n * n * nz * es / t / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": "%s %dx%dx%d of the %d^3 %s grid, pdims %dx%d as in-process ranks, %d timed round trips" % (
                "the whole grid" if nz == n else "z-slab", n, n, nz, n, args.dtype, pd[0], pd[1], reps)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d --master-addr "
                     "127.0.0.1 --master-port P bench.py --gpus %d ..." % (args.gpus, args.gpus))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

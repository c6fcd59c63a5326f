#!/usr/bin/env python
"""The reference's case sweep on the CPU, through the product's planner, launch preparation and tile decoding.

The reference's tests/test_runner.py expands tests/test_config.yaml into command lines for its legacy executables
(tests/cc/transpose_test.cc, halo_test.cc); oracle/make_ref_testfiles.py --max-per-config N writes those lists (up to
540 per configuration as the reference's runner caps them, 8 268 in all) to a directory. The GPU suite runs a thinned
sample of them through the executables themselves (tests/test_ref_executables_gpu.py). This script runs EVERY line of
such a directory without a GPU: the command line is parsed the way the executables parse it (including halo_test's
`--pdz`, which its option table routes to padding[1]), the transfers are planned by libcudecomp.so, every launch is
prepared by the product's launch code and walked CTA by CTA, lane by lane by the host-side launch emulator
(tests/host_emu: the kernels' own tile decoding, csrc/tiling.h) on numpy buffers, and the result of all ranks must equal
the CPU oracle byte for byte, cells that must stay untouched included:
  transpose lines: X->Y, Y->Z, Z->Y, Y->X, each direct and staged (every dtype size of the configuration); lines without
                   -o (in place) additionally through the fused staged schedule (phased launch, 2 and 4 chunks);
  halo lines:      the three dims in sequence on the line's axis, direct and staged.

    python scripts/ref_sweep_cpu.py oracle/_ref/cases_full --jobs 8 --summary profiles/r2_cpu_reference_sweep.txt
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DTYPE_BYTES = {"R32": 4, "R64": 8, "C32": 8, "C64": 16}


def take(tokens, i, n):
    return [int(v) for v in tokens[i + 1:i + 1 + n]], i + 1 + n


def parse_transpose(line):
    """tests/cc/transpose_test.cc:244-356."""
    t = line.split()
    a = dict(gx=0, gy=0, gz=0, pr=0, pc=0, rank_order=0, ac=[0, 0, 0], gd=[0, 0, 0], he=[[0] * 3 for _ in range(3)],
             pd=[[0] * 3 for _ in range(3)], mem_order=None, out_of_place=False)
    i = 0
    while i < len(t):
        k = t[i]
        if k in ("--gx", "--gy", "--gz", "--pr", "--pc"):
            a[k[2:]] = int(t[i + 1]); i += 2
        elif k == "--backend":
            i += 2
        elif k == "--rank-order":
            a["rank_order"] = int(t[i + 1]); i += 2
        elif k in ("--acx", "--acy", "--acz"):
            a["ac"]["xyz".index(k[4])] = int(t[i + 1]); i += 2
        elif k == "--gd":
            a["gd"], i = take(t, i, 3)
        elif k in ("--hex", "--hey", "--hez"):
            a["he"]["xyz".index(k[4])], i = take(t, i, 3)
        elif k in ("--pdx", "--pdy", "--pdz"):
            a["pd"]["xyz".index(k[4])], i = take(t, i, 3)
        elif k == "--mem_order":
            a["mem_order"], i = take(t, i, 9)
        elif k in ("-o", "--out-of-place"):
            a["out_of_place"] = True; i += 1
        elif k in ("-m", "--use-managed-memory"):
            i += 1
        else:
            raise ValueError("unknown option %r in %r" % (k, line))
    g = [a["gx"], a["gy"], a["gz"]]
    mo = a["mem_order"]
    if mo is not None and all(v < 0 for v in mo):
        mo = None
    return dict(gdims=g, pdims=[a["pr"], a["pc"]], axis_contiguous=[bool(v) for v in a["ac"]],
                mem_order=[mo[0:3], mo[3:6], mo[6:9]] if mo else None,
                gdims_dist=[g[k] - a["gd"][k] for k in range(3)] if any(a["gd"]) else None,
                col_major=a["rank_order"] == 2, halos={str(k): a["he"][k] for k in range(3)},
                pads={str(k): a["pd"][k] for k in range(3)}, out_of_place=a["out_of_place"])


def parse_halo(line):
    """tests/cc/halo_test.cc:259-333: one axis; scalars per dim; --pdz lands in padding[1] like --pdy (its option table)."""
    t = line.split()
    a = dict(gx=0, gy=0, gz=0, pr=0, pc=0, rank_order=0, ac=0, gd=[0, 0, 0], he=[0] * 3, hp=[0] * 3, pd=[0] * 3, ax=0,
             mem_order=None)
    i = 0
    while i < len(t):
        k = t[i]
        if k in ("--gx", "--gy", "--gz", "--pr", "--pc", "--ax", "--ac"):
            a[k[2:]] = int(t[i + 1]); i += 2
        elif k == "--backend":
            i += 2
        elif k == "--rank-order":
            a["rank_order"] = int(t[i + 1]); i += 2
        elif k == "--gd":
            a["gd"], i = take(t, i, 3)
        elif k in ("--hex", "--hey", "--hez"):
            a["he"]["xyz".index(k[4])] = int(t[i + 1]); i += 2
        elif k in ("--hpx", "--hpy", "--hpz"):
            a["hp"]["xyz".index(k[4])] = int(t[i + 1]); i += 2
        elif k == "--pdx":
            a["pd"][0] = int(t[i + 1]); i += 2
        elif k in ("--pdy", "--pdz"):
            a["pd"][1] = int(t[i + 1]); i += 2
        elif k == "--mem_order":
            a["mem_order"], i = take(t, i, 3)
        elif k in ("-m", "--use-managed-memory"):
            i += 1
        else:
            raise ValueError("unknown option %r in %r" % (k, line))
    g = [a["gx"], a["gy"], a["gz"]]
    mo = a["mem_order"]
    if mo is not None and all(v < 0 for v in mo):
        mo = None
    zero = {str(k): [0, 0, 0] for k in range(3)}
    d = dict(gdims=g, pdims=[a["pr"], a["pc"]], axis_contiguous=[bool(a["ac"])] * 3, mem_order=[mo, mo, mo] if mo else None,
             gdims_dist=[g[k] - a["gd"][k] for k in range(3)] if any(a["gd"]) else None, col_major=a["rank_order"] == 2,
             halos=zero, pads=zero)
    return d, a["ax"], a["he"], [bool(v) for v in a["hp"]], a["pd"]


def run_line(job):
    kind, line, sizes, idx = job
    import numpy as np
    from tests import test_launch_emulation as T
    from tests.test_planner_properties import make_config, make_oracle
    from cudecomp_b200 import capi as cd
    from oracle import oracle as orc
    from tests import host_emu as emu
    checks = 0
    try:
        for es in sizes:
            # launches that store into peers run with 256-bit accesses in the product: alternate the variants over the sweep
            s = dict(es=es, tile_bytes=0, peer_order=0, kernel_variant=2 if idx % 2 else 0, grid=0, threads=256, misalign=0)
            if kind == "transpose":
                d = parse_transpose(line)
                T.check_transposes(d, s)
                checks += 8
                if not d["out_of_place"]:
                    for K in (2, 4):
                        T.test_emulated_fused_staged_schedule_equals_oracle.hypothesis.inner_test(d, s, K, True, 1, False)
                        checks += 4
            else:
                d, ax, halo, periods, padding = parse_halo(line)
                cfg, o = make_config(d), make_oracle(d)
                if o.has_empty_pencils(ax):
                    continue
                n = o.nranks
                dt = T.DT[es]
                rng = np.random.default_rng(13 + idx)
                data = [emu.aligned_array(o.pencil_info(r, ax, halo, padding).size, dt, 0) for r in range(n)]
                for x in data:
                    T.rand_fill(x, rng)
                for staged in (False, True):
                    mine = [emu.aligned_array(x.size, dt, 0, 0) for x in data]
                    for m, x in zip(mine, data):
                        m[:] = x
                    ref = [x.copy() for x in data]
                    for dim in range(3):
                        o.halo(ax, dim, ref, halo, periods, padding)
                        plans = [cd.plan_halo_boxes(cfg, r, ax, dim, halo, periods, padding, staged) for r in range(n)]
                        works = [emu.aligned_array(max(o.halo_workspace_size(r, ax, halo), 1), dt, 0, -9) for r in range(n)]
                        snap = [emu.aligned_array(x.size, dt, 0, 0) for x in mine]
                        for sn, m in zip(snap, mine):
                            sn[:] = m
                        legal = mine + works + snap
                        for r in range(n):
                            push = [bx for bx in plans[r] if not bx["is_unpack"]]
                            T.emulate(push, lambda bx: snap[r], lambda bx: (works if staged else mine)[bx["peer_rank"]], legal, s)
                        if staged:
                            for r in range(n):
                                unpack = [bx for bx in plans[r] if bx["is_unpack"]]
                                T.emulate(unpack, lambda bx: works[r], lambda bx: mine[r], legal, s)
                        for r in range(n):
                            if not np.array_equal(mine[r], ref[r]):
                                raise AssertionError("halo differs: dim %d staged %d rank %d" % (dim, staged, r))
                        checks += 1
        return (kind, idx, True, checks, "")
    except Exception as e:  # noqa: BLE001
        return (kind, idx, False, checks, "%s: %s | %s" % (type(e).__name__, str(e)[:300], line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", help="directory written by oracle/make_ref_testfiles.py (index.n4.json + one list per configuration)")
    ap.add_argument("--jobs", type=int, default=max(1, len(os.sched_getaffinity(0))))
    ap.add_argument("--limit", type=int, default=0, help="at most this many lines per configuration (0 = all)")
    ap.add_argument("--summary", default=None)
    args = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with open(os.path.join(args.cases, "index.n4.json")) as f:
        index = json.load(f)
    t0 = time.time()
    rows, failures = [], []
    with ProcessPoolExecutor(max_workers=args.jobs) as pool:
        for name, info in sorted(index.items()):
            with open(os.path.join(args.cases, info["file"])) as f:
                lines = [ln.strip() for ln in f if ln.strip()]
            if args.limit:
                lines = lines[:args.limit]
            kind = "halo" if info["executable"] == "halo_test" else "transpose"
            sizes = sorted({DTYPE_BYTES[d] for d in info["dtypes"]})
            t1 = time.time()
            res = list(pool.map(run_line, [(kind, ln, sizes, i) for i, ln in enumerate(lines)], chunksize=4))
            ok = sum(1 for r in res if r[2])
            failures += [r for r in res if not r[2]]
            rows.append((name, len(lines), info.get("generated", len(lines)), sizes, ok, sum(r[3] for r in res), time.time() - t1))
            print("%-34s %5d lines  element sizes %-10s passed %5d  (%d launch-level comparisons, %.0f s)" % (
                name, len(lines), sizes, ok, rows[-1][5], rows[-1][6]), flush=True)
    total = sum(r[1] for r in rows)
    passed = sum(r[4] for r in rows)
    text = ["# The reference's sweep (tests/test_config.yaml expanded by its tests/test_runner.py) on the CPU: scripts/ref_sweep_cpu.py",
            "# every command line planned by libcudecomp.so, every launch walked by the host-side launch emulator, compared with the oracle",
            "# lines: run here / all combinations the reference's configuration spans for 4 ranks (its runner draws at most 540 of",
            "# them per run; the lists used here are 540 evenly spaced ones, oracle/make_ref_testfiles.py --max-per-config 540)", ""]
    for name, n, gen, sizes, ok, checks, secs in rows:
        text.append("%-34s %5d / %5d lines, element sizes %-10s: %5d passed, %d failed (%d comparisons, %.0f s)" % (
            name, n, gen, sizes, ok, n - ok, checks, secs))
    text.append("")
    text.append("TOTAL %d lines, %d passed, %d failed, %.0f s on %d processes" % (total, passed, total - passed,
                                                                                time.time() - t0, args.jobs))
    for f in failures[:40]:
        text.append("FAILED %s #%d: %s" % (f[0], f[1], f[4]))
    out = "\n".join(text) + "\n"
    print(out)
    if args.summary:
        with open(args.summary, "w") as f:
            f.write(out)
    sys.exit(0 if passed == total else 1)


if __name__ == "__main__":
    main()

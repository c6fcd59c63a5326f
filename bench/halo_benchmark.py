#!/usr/bin/env python
"""Halo-exchange benchmark: BASELINE.json config 4 (2048 x 2048 x 1024 float, slab 1 x N decomposition, halo width 2,
cudecompUpdateHalosX/Y/Z over all three dimensions). Reports microseconds per call and the bytes each GPU sends per
second; these calls are latency-dominated (<= 67 MB per GPU per call), so both are shown.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/halo_benchmark.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def halo_pattern(torch, p, grid, halo, periods, es, device, halo_cells):
    """Known answer of the reference's halo tests (tests/ctest/halo_tests.cc:197-236) as integer bit patterns, for a
    pencil with halos and no padding. halo_cells False: initializePencil -- interior cells carry their global linear
    index, halo cells are unset (-1). True: initializeReference -- the state after the halos of all three dimensions were
    updated: every cell carries the index of the (periodically wrapped) global cell it mirrors, -1 where there is none.
    int32 for 4-byte elements (index mod 2^32), int64 for 8-byte ones, (index, ~index) pairs for 16-byte ones."""
    idt = torch.int32 if es == 4 else torch.int64
    gstride = [1, grid[0], grid[0] * grid[1]]
    terms, masks = [], []
    for k in range(3):
        ax_g = p.order[k]
        g = torch.arange(p.shape[k], device=device, dtype=torch.int64) + (p.lo[k] - halo[ax_g])
        inside = (g >= p.lo[k]) & (g <= p.hi[k])
        if halo_cells:
            valid = torch.ones_like(inside) if periods[ax_g] else ((g >= 0) & (g < grid[ax_g]))
            g = torch.remainder(g, grid[ax_g])
        else:
            valid = inside
        terms.append(g * gstride[ax_g])
        masks.append(valid)
    idx = terms[2][:, None, None] + terms[1][None, :, None] + terms[0][None, None, :]
    ok = masks[2][:, None, None] & masks[1][None, :, None] & masks[0][None, None, :]
    out = torch.where(ok, idx, torch.full_like(idx, -1)).to(idt).reshape(-1)
    if es == 16:
        out = torch.stack([out, ~out], dim=1).reshape(-1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[2048, 2048, 1024])
    ap.add_argument("--pdims", default=None)
    ap.add_argument("--halo", type=int, nargs=3, default=[2, 2, 2])
    ap.add_argument("--dtype", default="float", choices=["float", "double", "float_complex", "double_complex"])
    ap.add_argument("--nonperiodic", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--staged", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed known-answer check")
    args = ap.parse_args()

    import torch
    from cudecomp_b200 import capi as cd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank))) % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    pd = [int(v) for v in args.pdims.split("x")] if args.pdims else [1, world]
    dt_enum = {"float": cd.CUDECOMP_FLOAT, "double": cd.CUDECOMP_DOUBLE, "float_complex": cd.CUDECOMP_FLOAT_COMPLEX,
               "double_complex": cd.CUDECOMP_DOUBLE_COMPLEX}[args.dtype]
    es = cd.DTYPE_SIZES[dt_enum]
    assert cd.MPI_Init() == 0
    res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res)
    cfg = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(cfg))
    cfg.gdims[:] = args.grid
    cfg.pdims[:] = pd
    cfg.halo_comm_backend = cd.CUDECOMP_HALO_COMM_NVSHMEM if args.staged else cd.CUDECOMP_HALO_COMM_NCCL
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    periods = [not args.nonperiodic] * 3
    stream = torch.cuda.current_stream()
    rows = []
    parity = {"checked_pencils": 0, "ok": True, "pattern": "global linear index per cell as integer bits (mod 2^32 for "
              "4-byte elements); after updating dims 0, 1, 2 every cell of the halo-inclusive pencil must hold the index "
              "of the (periodically wrapped) global cell it mirrors, -1 where there is none: the comparator of the "
              "reference's tests/ctest/halo_tests.cc:229-272, on the device, on every rank"}
    idt = torch.int32 if es == 4 else torch.int64

    def pattern(p, halo_cells):
        return halo_pattern(torch, p, args.grid, args.halo, periods, es, dev, halo_cells)

    for ax in range(3):
        res, p = cd.cudecompGetPencilInfo(handle, gd, ax, args.halo)
        cd.check(res)
        res, wsize = cd.cudecompGetHaloWorkspaceSize(handle, gd, ax, args.halo)
        res, work = cd.cudecompMalloc(handle, gd, max(wsize, 64) * es)
        cd.check(res)
        data = torch.zeros(p.size * es // 4, dtype=torch.float32, device=dev)
        shape_g = {p.order[i]: p.shape[i] for i in range(3)}
        if not args.no_parity:
            view = data.view(idt)
            view.copy_(pattern(p, False))
            for dim in range(3):
                cd.check(cd.UPDATE_HALOS[ax](handle, gd, data, work, dt_enum, args.halo, periods, dim, None, stream))
            want = pattern(p, True)
            good = bool(torch.equal(view, want))
            del want
            torch.cuda.synchronize()
            parity["checked_pencils"] += 1
            parity["ok"] = parity["ok"] and good
        for dim in range(3):
            face = args.halo[dim] * shape_g[(dim + 1) % 3] * shape_g[(dim + 2) % 3]
            call = lambda: cd.check(cd.UPDATE_HALOS[ax](handle, gd, data, work, dt_enum, args.halo, periods, dim,  # noqa: E731
                                                        None, stream))
            for _ in range(args.warmup):
                call()
            torch.cuda.synchronize()
            cd.MPI_Barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                call()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = cd.MPI_Allreduce_max(e0.elapsed_time(e1) / args.steps)
            path = cd.last_path(handle, gd)
            sent = 2 * face * es if path in (2, 3) else 0
            rows.append(dict(pencil="XYZ"[ax], dim=dim, us=ms * 1e3, path=["none", "local", "direct", "staged"][path],
                             bytes_sent_per_gpu=sent, moved_bytes=2 * face * es,
                             gbs=(2 * face * es / (ms * 1e-3) / 1e9) if ms > 0 else None))
        cd.check(cd.cudecompFree(handle, gd, work))
    if not args.no_parity:
        parity["ok"] = cd.MPI_Allreduce_max(0.0 if parity["ok"] else 1.0) == 0.0
        parity["ranks"] = world
    if rank == 0:
        line = {"benchmark": "halo update", "grid": args.grid, "pdims": pd, "halo": args.halo, "dtype": args.dtype,
                "periodic": not args.nonperiodic, "n_gpus": world, "calls": rows}
        if not args.no_parity:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    cd.cudecompGridDescDestroy(handle, gd)
    cd.cudecompFinalize(handle)
    cd.MPI_Finalize()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Distributed 3-D complex-to-complex FFT on top of the transpose engine: the caller of BASELINE.json config 3.

Same operation sequence as the reference's benchmark (benchmark/benchmark.cu:501-591): 1-D FFTs along the pencil axis
(cuFFT through torch.fft, batched over the pencil) interleaved with the four global transposes,
    forward : FFT_x, X->Y, FFT_y, Y->Z, FFT_z        backward : IFFT_z, Z->Y, IFFT_y, Y->X, IFFT_x
timed as (forward + backward) / 2 after warm-up, GFLOP/s = 5 N log2(N) / t (benchmark.cu:656-662), and the same
correctness check: random data, forward + backward + 1/N scaling must reproduce the input (max error <= 1e-10 for
double, 5e-4 for float, benchmark.cu:21-27,613-643). `--check-global` additionally compares the forward transform with
numpy.fft.fftn of the whole field (small grids only).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/fft_benchmark.py --grid 1024
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs="+", default=[1024])
    ap.add_argument("--pdims", default=None)
    ap.add_argument("--dtype", default="double_complex", choices=["double_complex", "float_complex"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--axis-contiguous", action="store_true")
    ap.add_argument("--check-global", action="store_true")
    ap.add_argument("--inplace", action="store_true",
                    help="one data buffer: in-place FFTs and in-place transposes (the reference benchmark's default)")
    ap.add_argument("--out", default=None, help="write the JSON result here as well")
    args = ap.parse_args()

    import numpy as np
    import torch
    from cudecomp_b200 import capi as cd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank))) % ndev)
    dev = torch.device("cuda", torch.cuda.current_device())
    g = args.grid if len(args.grid) == 3 else [args.grid[0]] * 3
    if args.pdims:
        pd = [int(v) for v in args.pdims.split("x")]
    else:
        pd = {1: [1, 1], 2: [1, 2], 4: [2, 2], 8: [2, 4]}.get(world, [1, world])
    cdt = torch.complex128 if args.dtype == "double_complex" else torch.complex64
    dt_enum = cd.CUDECOMP_DOUBLE_COMPLEX if args.dtype == "double_complex" else cd.CUDECOMP_FLOAT_COMPLEX
    es = 16 if args.dtype == "double_complex" else 8

    assert cd.MPI_Init() == 0
    res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res, "cudecompInit")
    cfg = cd.cudecompGridDescConfig_t()
    cd.check(cd.cudecompGridDescConfigSetDefaults(cfg))
    cfg.gdims[:] = g
    cfg.pdims[:] = pd
    cfg.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_NCCL
    for i in range(3):
        cfg.transpose_axis_contiguous[i] = args.axis_contiguous
    res, gd = cd.cudecompGridDescCreate(handle, cfg)
    cd.check(res, "cudecompGridDescCreate")
    pinfo = [cd.cudecompGetPencilInfo(handle, gd, ax)[1] for ax in range(3)]
    nelem = max(p.size for p in pinfo)
    res, wsize = cd.cudecompGetTransposeWorkspaceSize(handle, gd)
    res, work = cd.cudecompMalloc(handle, gd, wsize * es)
    cd.check(res, "cudecompMalloc")

    def view(buf, ax):
        """Pencil `ax` of the flat buffer as a (shape[2], shape[1], shape[0]) tensor, and the torch dim of the pencil
        axis (memory position 0 is the last torch dim)."""
        p = pinfo[ax]
        t = buf[:p.size].view(p.shape[2], p.shape[1], p.shape[0])
        pos = list(p.order).index(ax)
        return t, 2 - pos

    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    a = torch.empty(nelem, dtype=cdt, device=dev)
    b = torch.empty(nelem, dtype=cdt, device=dev)
    ref = torch.view_as_complex(torch.rand(pinfo[0].size, 2, generator=gen, device=dev,
                                           dtype=torch.float64 if es == 16 else torch.float32))
    stream = torch.cuda.current_stream()

    def fft(src, dst, ax, inverse):
        s, d = view(src, ax)
        o, _ = view(dst, ax)
        if inverse:
            torch.fft.ifft(s, dim=d, norm="forward", out=o)  # unnormalised, like cuFFT
        else:
            torch.fft.fft(s, dim=d, out=o)

    def transpose(op, src, dst):
        cd.check(cd.TRANSPOSES[op](handle, gd, src, dst, work, dt_enum, None, None, None, None, stream), op)

    def forward(x, y):      # data in x (X pencil) -> result in y (Z pencil, spectral)
        if args.inplace:
            fft(x, x, 0, False)
            transpose("XY", x, x)
            fft(x, x, 1, False)
            transpose("YZ", x, x)
            fft(x, x, 2, False)
            y[:pinfo[2].size].copy_(x[:pinfo[2].size]) if y is not x else None
            return
        fft(x, y, 0, False)
        transpose("XY", y, x)
        fft(x, y, 1, False)
        transpose("YZ", y, x)
        fft(x, y, 2, False)

    def backward(x, y):     # data in y (Z pencil) -> result in y (X pencil), unnormalised
        if args.inplace:
            fft(y, y, 2, True)
            transpose("ZY", y, y)
            fft(y, y, 1, True)
            transpose("YX", y, y)
            fft(y, y, 0, True)
            return
        fft(y, x, 2, True)
        transpose("ZY", x, y)
        fft(y, x, 1, True)
        transpose("YX", x, y)
        fft(y, x, 0, True)
        y[:pinfo[0].size].copy_(x[:pinfo[0].size])

    # ---- correctness
    if args.inplace:
        b = a  # a single data buffer
    a[:pinfo[0].size].copy_(ref)
    forward(a, b)
    spectral = b[:pinfo[2].size].clone()
    backward(a, b)
    n_total = float(g[0]) * g[1] * g[2]
    err = (b[:pinfo[0].size] / n_total - ref).abs().max().item()
    tol = 1e-10 if es == 16 else 5e-4
    max_err = cd.MPI_Allreduce_max(err)  # control-plane reduction over the library's bootstrap
    ok = max_err <= tol
    global_err = None
    if args.check_global:
        # every rank rebuilds the whole field from the per-rank seeds and transforms it with numpy
        full = np.zeros((g[2], g[1], g[0]), dtype=np.complex128)
        for r in range(world):
            gr = torch.Generator(device=dev)
            gr.manual_seed(4321 + r)
            pr = pencil_of_rank(cd, cfg, g, pd, r, 0, args.axis_contiguous)
            blk = torch.view_as_complex(torch.rand(pr["size"], 2, generator=gr, device=dev,
                                                   dtype=torch.float64 if es == 16 else torch.float32)).cpu().numpy()
            place(full, blk, pr)
        want = np.fft.fftn(full)
        mine = pencil_of_rank(cd, cfg, g, pd, rank, 2, args.axis_contiguous)
        got = np.zeros_like(full)
        place(got, spectral.cpu().numpy(), mine)
        sl = region(mine)
        global_err = float(np.abs(got[sl] - want[sl]).max() / np.abs(want).max())
        ok = ok and global_err < (1e-12 if es == 16 else 1e-5)

    # ---- timing
    for _ in range(args.warmup):
        forward(a, b)
        backward(a, b)
    torch.cuda.synchronize()
    cd.MPI_Barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        forward(a, b)
        backward(a, b)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = cd.MPI_Allreduce_max(e0.elapsed_time(e1) / args.steps / 2.0)  # ms per transform, slowest rank
    gflops = 5.0 * n_total * np.log2(n_total) / (ms * 1e-3) / 1e9
    if rank == 0:
        line = {"benchmark": "3-D C2C FFT (cuFFT per pencil + 4 transposes), time per forward-or-backward transform",
                "grid": g, "pdims": pd, "dtype": args.dtype, "n_gpus": world, "ms": ms, "gflops": gflops,
                "max_roundtrip_error": max_err, "tolerance": tol, "global_fftn_rel_error": global_err,
                "passed": bool(ok), "axis_contiguous": args.axis_contiguous, "inplace": args.inplace}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(line, f)
    cd.cudecompFree(handle, gd, work)
    cd.cudecompGridDescDestroy(handle, gd)
    cd.cudecompFinalize(handle)
    cd.MPI_Finalize()
    sys.exit(0 if ok else 1)


def pencil_of_rank(cd, cfg, g, pd, r, ax, axis_contiguous):
    """Pencil geometry of any rank from the oracle-free closed form (row-major ranks, even or uneven splits)."""
    order = [(ax + i) % 3 for i in range(3)] if axis_contiguous else [0, 1, 2]
    pidx = [r // pd[1], r % pd[1]]
    lo, ext, j = [0] * 3, [0] * 3, 0
    for i in range(3):
        if i == ax:
            ext[i] = g[i]
        else:
            q, m = divmod(g[i], pd[j])
            ext[i] = q + (1 if pidx[j] < m else 0)
            lo[i] = pidx[j] * q + min(pidx[j], m)
            j += 1
    return dict(order=order, lo=lo, ext=ext, size=ext[0] * ext[1] * ext[2])


def region(p):
    return (slice(p["lo"][2], p["lo"][2] + p["ext"][2]), slice(p["lo"][1], p["lo"][1] + p["ext"][1]),
            slice(p["lo"][0], p["lo"][0] + p["ext"][0]))


def place(full, flat, p):
    """Scatter a flat pencil (memory order p['order']) into the (z, y, x) global array."""
    shp = [p["ext"][p["order"][2]], p["ext"][p["order"][1]], p["ext"][p["order"][0]]]
    blk = flat.reshape(shp)
    # torch/numpy dims (2,1,0) hold global axes order[0], order[1], order[2]; bring them to (z, y, x)
    axes_now = [p["order"][2], p["order"][1], p["order"][0]]  # global axis of each numpy dim
    perm = [axes_now.index(2), axes_now.index(1), axes_now.index(0)]
    full[region(p)] = blk.transpose(perm)


if __name__ == "__main__":
    main()

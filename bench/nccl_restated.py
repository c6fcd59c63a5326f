#!/usr/bin/env python
"""The reference's NCCL transpose arm, restated with torch ops and torch.distributed -- a GPU-side BASELINE and an
independent cross-check, not part of the product (the product never imports this file).

The reference cannot be built in this image (no MPI, cuTENSOR, NVHPC: DESIGN.md section 6), so the comparison SURVEY.md
section 8(c) asks for -- "compare byte-for-byte against a rebuilt NCCL-send/recv path (same pack semantics +
ncclSend/ncclRecv, public NCCL only)" -- is realised here with the same three phases and the same wire format as
reference include/internal/transpose.h:196-905 (default branch, no halos):

  pack    per destination rank i, the sub-block a in [off_a[i], +splits_a[i]) of the source pencil is copied into the
          contiguous send region at element offset off_a[i] * |b| * |c|, in the SOURCE memory order
          (transpose.h:549-598; one strided torch copy per peer instead of the batched kernel)
  a2a     torch.distributed.all_to_all_single on the row / column group = grouped ncclSend/ncclRecv, which is exactly
          what the reference's NCCL arm issues (include/internal/comm_routines.h:297-324)
  unpack  per source rank j, the received block (rank j's memory order, a-extent = my split) is copied / permuted into
          the destination pencil at b in [off_b[j], +splits_b[j]) (transpose.h:833-893; cuTENSOR permute when the
          orders differ, :80-157 -- here torch's strided copy)

Geometry is restated from the reference's formulas (src/cudecomp.cc:1335-1373, include/internal/common.h:318-346,
579-589) in plain Python so that this file depends on neither the product library nor the oracle.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/nccl_restated.py --grid 1024
prints one JSON line: effective GB/s of the X->Y->Z->Y->X round trip, same accounting as bench.py.
"""
import argparse
import json
import os
import sys

OPS = {"XY": (0, 1), "YZ": (1, 2), "ZY": (2, 1), "YX": (1, 0)}  # (source pencil axis a, destination pencil axis b)


def get_splits(n, p, pad):
    """reference include/internal/common.h:579-589"""
    s = [n // p + (1 if i < n % p else 0) for i in range(p)]
    s[min(n, p) - 1] += pad
    return s


def offsets(splits):
    out, acc = [], 0
    for s in splits:
        out.append(acc)
        acc += s
    return out


class Geometry:
    """Pencil shapes and communicators of a pdims[0] x pdims[1] row-major process grid (reference
    src/cudecomp.cc:1335-1373, common.h:318-346); `order[axis][k]` = global axis at memory position k (0 fastest)."""

    def __init__(self, gdims, pdims, axis_contiguous=(False, False, False), mem_order=None, gdims_dist=None):
        self.gdims = list(gdims)
        self.dist = list(gdims_dist) if gdims_dist else list(gdims)
        self.pdims = list(pdims)
        self.nranks = pdims[0] * pdims[1]
        if mem_order:
            self.order = [list(o) for o in mem_order]
        else:
            self.order = [[(ax + i) % 3 for i in range(3)] if axis_contiguous[ax] else [0, 1, 2] for ax in range(3)]

    def pidx(self, rank):
        return [rank // self.pdims[1], rank % self.pdims[1]]

    def rank_of(self, pidx):
        return pidx[0] * self.pdims[1] + pidx[1]

    def extents(self, rank, axis):
        """extent of every GLOBAL axis in `rank`'s `axis`-pencil, and its lower corner"""
        pidx = self.pidx(rank)
        ext, lo, j = [0] * 3, [0] * 3, 0
        for i in range(3):
            if i == axis:
                ext[i], lo[i] = self.gdims[i], 0
                continue
            s = get_splits(self.dist[i], self.pdims[j], self.gdims[i] - self.dist[i])
            ext[i], lo[i] = s[pidx[j]], offsets(s)[pidx[j]]
            j += 1
        return ext, lo

    def torch_shape(self, rank, axis):
        """shape of the pencil as a torch tensor (slowest dimension first)"""
        ext, _ = self.extents(rank, axis)
        return [ext[self.order[axis][2 - d]] for d in range(3)]

    def torch_dim(self, axis, g):
        """torch dimension that holds global axis g in an `axis`-pencil"""
        return 2 - self.order[axis].index(g)

    def comm(self, rank, a, b):
        """ranks of the communicator of transpose a -> b in communicator order, and my index in it: the row
        communicator when the Z pencil is involved, else the column communicator (transpose.h:227)"""
        pidx = self.pidx(rank)
        ci = 1 if 2 in (a, b) else 0
        members = []
        for i in range(self.pdims[ci]):
            q = list(pidx)
            q[ci] = i
            members.append(self.rank_of(q))
        return members, pidx[ci], ci


class RestatedTranspose:
    """pack / exchange / unpack of one rank. `exchange(send, in_counts, out_counts, members)` returns the receive
    buffer; the default uses torch.distributed (NCCL on GPUs); tests inject an in-process exchange."""

    def __init__(self, geom, rank, exchange=None):
        self.g = geom
        self.rank = rank
        self.exchange = exchange or self._dist_exchange
        self.groups = {}

    # ---- geometry of one operation
    def plan(self, op):
        a, b = OPS[op]
        c = 3 - a - b
        g = self.g
        members, me, ci = g.comm(self.rank, a, b)
        P = len(members)
        splits_a = get_splits(g.dist[a], P, g.gdims[a] - g.dist[a])
        splits_b = get_splits(g.dist[b], P, g.gdims[b] - g.dist[b])
        ext_src, _ = g.extents(self.rank, a)
        nc = ext_src[c]
        send_counts = [splits_a[i] * splits_b[me] * nc for i in range(P)]
        recv_counts = [splits_b[j] * splits_a[me] * nc for j in range(P)]
        return dict(a=a, b=b, c=c, members=members, me=me, P=P, splits_a=splits_a, splits_b=splits_b,
                    off_a=offsets(splits_a), off_b=offsets(splits_b), nc=nc, send_counts=send_counts,
                    recv_counts=recv_counts)

    def pack(self, op, src, send):
        """src: flat tensor holding my a-pencil; send: flat staging tensor (>= pencil size)"""
        p = self.plan(op)
        g, a = self.g, p["a"]
        view = src[:_prod(g.torch_shape(self.rank, a))].view(g.torch_shape(self.rank, a))
        da = g.torch_dim(a, a)
        pos = 0
        for i in range(p["P"]):
            blk = view.narrow(da, p["off_a"][i], p["splits_a"][i])
            n = p["send_counts"][i]
            send[pos:pos + n].view(blk.shape).copy_(blk)
            pos += n
        return p

    def unpack(self, op, recv, dst, p=None):
        p = p or self.plan(op)
        g, a, b, c = self.g, p["a"], p["b"], p["c"]
        dview = dst[:_prod(g.torch_shape(self.rank, b))].view(g.torch_shape(self.rank, b))
        db = g.torch_dim(b, b)
        pos = 0
        for j in range(p["P"]):
            n = p["recv_counts"][j]
            ext = {a: p["splits_a"][p["me"]], b: p["splits_b"][j], c: p["nc"]}
            # the block as rank j packed it: its source memory order, slowest first
            shape_src = [ext[g.order[a][2 - d]] for d in range(3)]
            blk = recv[pos:pos + n].view(shape_src)
            # reorder its dimensions to the destination pencil's: destination dim d holds global axis order[b][2 - d]
            perm = [g.torch_dim(a, g.order[b][2 - d]) for d in range(3)]
            dview.narrow(db, p["off_b"][j], p["splits_b"][j]).copy_(blk.permute(perm))
            pos += n
        return p

    def transpose(self, op, src, dst, send, recv):
        p = self.pack(op, src, send)
        if p["P"] == 1:
            self.unpack(op, send, dst, p)  # single rank: the send region is the receive region (transpose.h:342-362)
            return
        self.exchange(send, recv, p["send_counts"], p["recv_counts"], p["members"])
        self.unpack(op, recv, dst, p)

    # ---- torch.distributed exchange = grouped ncclSend / ncclRecv
    def create_groups(self):
        """every rank creates every row and column group (torch.distributed requires it), keyed by member tuple"""
        import torch.distributed as dist
        g = self.g
        seen = set()
        for r in range(g.nranks):
            for (a, b) in ((0, 1), (1, 2)):
                members = tuple(g.comm(r, a, b)[0])
                if members in seen:
                    continue
                seen.add(members)
                grp = dist.new_group(list(members)) if len(members) > 1 else None
                self.groups[members] = grp

    def _dist_exchange(self, send, recv, send_counts, recv_counts, members):
        import torch
        import torch.distributed as dist
        n_in, n_out = sum(send_counts), sum(recv_counts)
        es = send.element_size()
        # exchanged as raw bytes: the payload is never computed on, and not every backend accepts complex tensors
        dist.all_to_all_single(recv[:n_out].view(torch.uint8), send[:n_in].view(torch.uint8), [c * es for c in recv_counts],
                               [c * es for c in send_counts], group=self.groups[tuple(members)])


def _prod(shape):
    out = 1
    for s in shape:
        out *= s
    return out


def selftest_gloo(args):
    """Every rank fills its X pencil with the global linear index of each cell, runs the four transposes over gloo and
    checks every intermediate pencil against the same analytic pattern (the reference tests' known answer,
    tests/cc/transpose_test.cc:103-155)."""
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pd = [int(v) for v in args.pdims.split("x")] if args.pdims else {1: [1, 1], 2: [1, 2], 4: [2, 2], 8: [2, 4]}.get(world, [1, world])
    gd = [args.grid, args.grid + 1, args.grid + 2]
    geom = Geometry(gd, pd, [args.axis_contiguous] * 3)
    rt = RestatedTranspose(geom, rank)
    rt.create_groups()

    def pattern(axis):
        ext, lo = geom.extents(rank, axis)
        idx = [torch.arange(lo[g], lo[g] + ext[g], dtype=torch.float64) for g in range(3)]
        shape = geom.torch_shape(rank, axis)
        out = torch.zeros(shape, dtype=torch.float64)
        for g in range(3):
            view = [1, 1, 1]
            view[geom.torch_dim(axis, g)] = ext[g]
            out = out + idx[g].view(view) * [1, gd[0], gd[0] * gd[1]][g]
        return out.reshape(-1)

    nelem = max(_prod(geom.torch_shape(rank, ax)) for ax in range(3))
    bufs = [torch.zeros(nelem, dtype=torch.float64), torch.zeros(nelem, dtype=torch.float64)]
    send, recv = torch.zeros(nelem, dtype=torch.float64), torch.zeros(nelem, dtype=torch.float64)
    x = pattern(0)
    bufs[0][:x.numel()] = x
    cur = 0
    for op in ("XY", "YZ", "ZY", "YX"):
        rt.transpose(op, bufs[cur], bufs[1 - cur], send, recv)
        cur = 1 - cur
        want = pattern(OPS[op][1])
        if not torch.equal(bufs[cur][:want.numel()], want):
            print("rank %d: %s differs from the analytic pattern" % (rank, op), flush=True)
            return 1
    ok = torch.tensor([1.0])
    dist.all_reduce(ok)
    if rank == 0 and ok.item() == world:
        print("SELFTEST OK", flush=True)
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--dtype", default="double_complex", choices=["double_complex", "float_complex", "double", "float"])
    ap.add_argument("--pdims", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--axis-contiguous", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--selftest-gloo", action="store_true", help="CPU tensors over gloo: checks the distributed code path")
    args = ap.parse_args()
    if args.selftest_gloo:
        return selftest_gloo(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank))) % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pd = [int(v) for v in args.pdims.split("x")] if args.pdims else {1: [1, 1], 2: [1, 2], 4: [2, 2], 8: [2, 4]}.get(world, [1, world])
    tdt = {"double_complex": torch.complex128, "float_complex": torch.complex64, "double": torch.float64,
           "float": torch.float32}[args.dtype]
    es = torch.empty(0, dtype=tdt).element_size()
    geom = Geometry([args.grid] * 3, pd, [args.axis_contiguous] * 3)
    rt = RestatedTranspose(geom, rank)
    rt.create_groups()
    nelem = max(_prod(geom.torch_shape(rank, ax)) for ax in range(3))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    real = torch.float64 if tdt in (torch.complex128, torch.float64) else torch.float32
    x = torch.rand(nelem * (2 if tdt.is_complex else 1), generator=gen, device=dev, dtype=real)
    a = torch.view_as_complex(x.view(-1, 2)) if tdt.is_complex else x
    ref = a.clone()
    b = torch.zeros_like(a)
    send, recv = torch.empty_like(a), torch.empty_like(a)

    def round_trip():
        rt.transpose("XY", a, b, send, recv)
        rt.transpose("YZ", b, a, send, recv)
        rt.transpose("ZY", a, b, send, recv)
        rt.transpose("YX", b, a, send, recv)

    for _ in range(args.warmup):
        round_trip()
    nx = _prod(geom.torch_shape(rank, 0))
    ok = bool(torch.equal(a[:nx], ref[:nx]))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        round_trip()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    S = nx * es
    if rank == 0:
        line = {"impl": "nccl-restated", "metric": "effective transpose GB/s (4*S/t round trip, whole job)",
                "value": world * 4 * S / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "round_trip_is_identity": ok,
                "config": {"workload": "%d^3 %s X->Y->Z->Y->X round trip, out-of-place, pdims %dx%d, pack (torch strided "
                                       "copies) + all_to_all_single (NCCL) + unpack (torch strided copies)" %
                                       (args.grid, args.dtype, pd[0], pd[1])}}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(line) + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

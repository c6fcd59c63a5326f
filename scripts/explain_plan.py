#!/usr/bin/env python
"""What the engine will do for a configuration, computed on the host (no GPU): per operation the boxes each rank
pushes, their tiling and grid-stride rounds, and -- for staged / in-place calls -- the chunked schedule of the fused launch:
which peers every step talks to and how much of the local unpack overlaps later pushes. A design aid; every number here
is arithmetic on the plans (cudecompB200Plan* entry points), not a measurement. The time model at the end uses the rates
measured in round 2 (profiles/r2_*): HBM copy 6.45 TB/s, 5.0 TB/s for the five-stream read/write mix of a fused staged
launch, SM stores over NVLink 0.69 TB/s per direction (256-bit stores), 4 us per in-kernel handshake, 10 us per chunk.
It gives 7.81 / 10.48 / 8.9 ms for the 1024^3 complex128 2x4 round trip out of place / in place with separate launches /
in place fused; measured 7.83 / 10.5 / 9.0-9.1 ms (profiles/r2_n8_results.md).

    python scripts/explain_plan.py --grid 1024 --pdims 2x4 --dtype double_complex --chunks 8
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OPS = {"XY": (0, 1), "YZ": (1, 1), "ZY": (2, -1), "YX": (1, -1)}
ES = {"float": 4, "double": 8, "float_complex": 8, "double_complex": 16}
HBM, HBM_MIX, LINK, HANDSHAKE_US, CHUNK_US = 6.45e12, 5.0e12, 0.69e12, 4.0, 10.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs="+", default=[1024])
    ap.add_argument("--pdims", default="2x4")
    ap.add_argument("--dtype", default="double_complex", choices=sorted(ES))
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--tile-bytes", type=int, default=32768)
    ap.add_argument("--ctas", type=int, default=370)
    ap.add_argument("--axis-contiguous", action="store_true")
    args = ap.parse_args()
    from cudecomp_b200 import capi as cd
    g = args.grid if len(args.grid) == 3 else [args.grid[0]] * 3
    pd = [int(v) for v in args.pdims.split("x")]
    es = ES[args.dtype]
    c = cd.cudecompGridDescConfig_t()
    cd.cudecompGridDescConfigSetDefaults(c)
    c.gdims[:] = g
    c.pdims[:] = pd
    for i in range(3):
        c.transpose_axis_contiguous[i] = args.axis_contiguous
    print("grid %s, pdims %dx%d, %s, rank %d" % (g, pd[0], pd[1], args.dtype, args.rank))
    total_direct = total_staged = total_chunked = 0.0
    for op, (ax, d) in OPS.items():
        boxes = cd.plan_transpose_boxes(c, args.rank, ax, d)
        nbytes = [int(np.prod(b["extent"])) * es for b in boxes]
        S = sum(nbytes)
        wire = sum(n for n, b in zip(nbytes, boxes) if b["peer_rank"] != args.rank)
        tiles = [-(-n // args.tile_bytes) for n in nbytes]
        slots = len(boxes) * max(tiles)
        rounds = slots / args.ctas
        util = slots / (args.ctas * -(-slots // args.ctas))
        t_direct = max(wire / LINK, 2 * S / HBM) * 1e3 + (2 * HANDSHAKE_US * 1e-3 if wire else 0)
        t_unpack = 2 * S / HBM * 1e3
        print("\n%s: %d boxes, pencil %.1f MB, %.1f MB leave the GPU; %d tiles of %d KiB -> %.2f rounds of %d CTAs "
              "(last-round utilisation %.1f %%)" % (op, len(boxes), S / 1e6, wire / 1e6, sum(tiles), args.tile_bytes // 1024,
                                                     rounds, args.ctas, 100 * util))
        for b, n in zip(boxes, nbytes):
            print("   -> rank %d: extent %s, %.1f MB" % (b["peer_rank"], b["extent"], n / 1e6))
        print("   model: direct %.3f ms; staged (in place) %.3f ms = push + local unpack %.3f ms" %
              (t_direct, t_direct + t_unpack, t_unpack))
        total_direct += t_direct
        total_staged += t_direct + t_unpack
        if len(boxes) > 1 and args.chunks > 1:
            K = args.chunks
            # as the engine plans it: in place, column chunks where they apply (element size in bits 8-15 of the flags)
            pb = cd.plan_pipelined_transpose_boxes(c, args.rank, ax, d, None, None, None, None, 1 + (es << 8), K,
                                                   max_boxes=8192)
            K = 1 + max(b["step"] for b in pb)
            unp, peers = [0] * K, [set() for _ in range(K)]
            for b in pb:
                n = int(np.prod(b["extent"])) * es
                if b["is_unpack"]:
                    unp[b["step"]] += n
                else:
                    peers[b["step"]].add(b["peer_rank"])
            exposed = unp[-1]
            t_chunked = max(wire / LINK, 4 * S / HBM_MIX) * 1e3 + (K * CHUNK_US + HANDSHAKE_US) * 1e-3 + 2 * exposed / HBM * 1e3
            print("   fused, %d chunks, in place: peers per step %s; unpacked beside later pushes %.0f %%, after the last "
                  "push %.0f %%  -> model %.3f ms" % (K, sorted({len(p) for p in peers}),
                                                      100 * (1 - exposed / max(sum(unp), 1)), 100 * exposed / max(sum(unp), 1),
                                                      t_chunked))
            print("     unpack MB per step: %s" % [round(u / 1e6) for u in unp])
            total_chunked += t_chunked
        else:
            total_chunked += t_direct + t_unpack
    print("\nround trip, model: out of place (direct) %.2f ms; in place, separate launches %.2f ms; in place fused (K = %d) %.2f ms" %
          (total_direct, total_staged, args.chunks, total_chunked))


if __name__ == "__main__":
    main()

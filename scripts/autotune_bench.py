#!/usr/bin/env python
"""BASELINE.json config 5: autotune sweep over all P_row x P_col factorisations (768^3 complex128 by default).
Launch with torchrun (one rank per GPU). Prints the library's autotune log and one JSON line with the selection."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=768)
    ap.add_argument("--dtype", default="double_complex")
    ap.add_argument("--inplace", action="store_true")
    ap.add_argument("--backend", action="store_true", help="also autotune the schedule (transpose backend)")
    args = ap.parse_args()
    import torch
    from cudecomp_b200 import capi as cd
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    assert cd.MPI_Init() == 0
    res, handle = cd.cudecompInit(cd.MPI_COMM_WORLD)
    cd.check(res)
    cfg = cd.cudecompGridDescConfig_t()
    cd.cudecompGridDescConfigSetDefaults(cfg)
    cfg.gdims[:] = [args.grid] * 3
    cfg.pdims[:] = [0, 0]
    cfg.transpose_comm_backend = cd.CUDECOMP_TRANSPOSE_COMM_NCCL
    opt = cd.cudecompGridDescAutotuneOptions_t()
    cd.cudecompGridDescAutotuneOptionsSetDefaults(opt)
    opt.dtype = {"double_complex": cd.CUDECOMP_DOUBLE_COMPLEX, "float_complex": cd.CUDECOMP_FLOAT_COMPLEX,
                 "double": cd.CUDECOMP_DOUBLE, "float": cd.CUDECOMP_FLOAT}[args.dtype]
    opt.autotune_transpose_backend = args.backend
    for i in range(4):
        opt.transpose_use_inplace_buffers[i] = args.inplace
    t0 = time.time()
    res, gd = cd.cudecompGridDescCreate(handle, cfg, opt)
    cd.check(res, "cudecompGridDescCreate (autotune)")
    if rank == 0:
        print(json.dumps({"autotune": {"grid": args.grid, "dtype": args.dtype, "inplace": args.inplace,
                                       "selected_pdims": list(cfg.pdims),
                                       "selected_backend": cd.cudecompTransposeCommBackendToString(cfg.transpose_comm_backend),
                                       "seconds": time.time() - t0}}), flush=True)
    cd.cudecompGridDescDestroy(handle, gd)
    cd.cudecompFinalize(handle)
    cd.MPI_Finalize()


if __name__ == "__main__":
    main()

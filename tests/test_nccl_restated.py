"""bench/nccl_restated.py -- the reference's NCCL arm restated with torch ops (pack -> all_to_all -> unpack in the
reference's wire format) -- checked on the CPU: (1) all ranks in one process with an in-process exchange, against the
oracle, over random decompositions and memory orders; (2) its torch.distributed code path on 4 real processes over gloo.
The restated arm is the GPU-side baseline and the byte-for-byte cross-check SURVEY.md section 8(c) asks for
(tests/test_zz_nccl_crosscheck_gpu.py); this file makes sure the baseline itself is right."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import oracle as orc
from tests.test_planner_properties import decompositions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_restated():
    """bench/ is a directory of scripts next to bench.py, not a package: load the file by path"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("nccl_restated", os.path.join(ROOT, "bench", "nccl_restated.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_nr = load_restated()
OPS, Geometry, RestatedTranspose, _prod = _nr.OPS, _nr.Geometry, _nr.RestatedTranspose, _nr._prod
EXAMPLES = int(os.environ.get("CDB_HYPOTHESIS_EXAMPLES", "150"))


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(decompositions())
def test_restated_nccl_arm_equals_oracle(d):
    if d["col_major"]:
        d = dict(d, col_major=False)  # the restated arm is row-major only (the reference's default)
    o = orc.Oracle(d["gdims"], d["pdims"], d["axis_contiguous"], d["mem_order"], d["gdims_dist"], False)
    geom = Geometry(d["gdims"], d["pdims"], d["axis_contiguous"], d["mem_order"], d["gdims_dist"])
    n = o.nranks
    rng = np.random.default_rng(1)
    for op, (a, b) in OPS.items():
        if o.has_empty_pencils(a) or o.has_empty_pencils(b):
            continue
        for r in range(n):
            assert _prod(geom.torch_shape(r, a)) == o.pencil_info(r, a).size
        ins = [rng.integers(1, 1 << 40, o.pencil_info(r, a).size).astype(np.int64) for r in range(n)]
        want = [np.zeros(o.pencil_info(r, b).size, np.int64) for r in range(n)]
        o.transpose(op, ins, want)
        mailbox = {}

        def make_exchange(r):
            def exchange(send, recv, send_counts, recv_counts, members):
                pos = 0
                for i, m in enumerate(members):
                    mailbox[(r, m)] = send[pos:pos + send_counts[i]].clone()
                    pos += send_counts[i]
            return exchange

        ranks = [RestatedTranspose(geom, r, make_exchange(r)) for r in range(n)]
        sends = [torch.zeros(max(x.size, w.size) + 1, dtype=torch.int64) for x, w in zip(ins, want)]
        plans = []
        for r in range(n):
            p = ranks[r].pack(op, torch.from_numpy(ins[r]), sends[r])
            ranks[r].exchange(sends[r], None, p["send_counts"], p["recv_counts"], p["members"])
            plans.append(p)
        for r in range(n):
            p = plans[r]
            recv = torch.cat([mailbox[(m, r)] for m in p["members"]]) if p["members"] else torch.zeros(0, dtype=torch.int64)
            assert [mailbox[(m, r)].numel() for m in p["members"]] == p["recv_counts"]
            out = torch.zeros(want[r].size, dtype=torch.int64)
            ranks[r].unpack(op, recv, out, p)
            assert np.array_equal(out.numpy(), want[r]), (d, op, r)


@pytest.mark.parametrize("nranks,grid,extra", [(2, 12, []), (4, 10, ["--axis-contiguous"]), (4, 9, ["--pdims", "4x1"])])
def test_restated_nccl_arm_over_gloo(nranks, grid, extra):
    from tests._launcher import free_port
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "bench", "nccl_restated.py"), "--selftest-gloo",
           "--grid", str(grid), "--dtype", "double"] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SELFTEST OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]

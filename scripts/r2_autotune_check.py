#!/usr/bin/env python
"""Quick hardware check of the autotuner with its schedule phase on by default: the autotune cases of
tests/test_gpu_parity.py (4 ranks sharing the GPUs that are there), plus one that tunes the backend for in-place buffers
on a grid large enough for the chunk alternatives to differ. Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._launcher import run_ranks  # noqa: E402

CASES = [
    dict(kind="autotune", name="AutotuneTransposeGrid", gdims=[24, 20, 28], dtype="double", n_trials=2),
    dict(kind="autotune", name="AutotuneTransposeBackend", gdims=[24, 20, 28], dtype="float_complex", autotune_backend=True,
         n_trials=1),
    dict(kind="autotune", name="AutotuneTransposeBackendLarger", gdims=[256, 128, 96], dtype="double_complex",
         autotune_backend=True, n_trials=1),
    dict(kind="autotune", name="AutotuneHaloGrid", gdims=[24, 20, 28], dtype="float", grid_mode=1, halo=[1, 1, 1], n_trials=1),
    dict(kind="transpose", name="ChainAfterAutotune", gdims=[32, 40, 48], pdims=[2, 2], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"]),
]

if __name__ == "__main__":
    results, logs = run_ranks(4, "gpu", CASES, timeout=100)
    ok = True
    for i, c in enumerate(CASES):
        bad = [r for r in range(4) if not results[r][i]["ok"]]
        print(c["name"], "OK" if not bad else "FAILED on ranks %s: %s" % (bad, results[bad[0]][i].get("msg")),
              results[0][i].get("msg", ""))
        ok = ok and not bad
    sel = [l for l in (logs[0] if logs else "").splitlines() if "SELECTED" in l]
    print("\n".join(sel[:8]))
    sys.exit(0 if ok else 1)

"""Runs the REFERENCE's own legacy test executables (tests/cc/transpose_test.cc, tests/cc/halo_test.cc) against this
library. The binaries are built UNMODIFIED from the reference sources by oracle/ref_tests.mk into oracle/_ref/
(they travel to the GPU box with the snapshot); the case lists are what the reference's tests/test_runner.py
generates from its tests/test_config.yaml (oracle/make_ref_testfiles.py -> tests/golden/ref_cases/). Each executable
carries the reference's own known-answer generator and comparator, so "Passed all tests." is the reference's verdict.
"""
import json
import os
import signal
import subprocess

import pytest

from tests._launcher import free_port

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref")
# default: the small committed sample; CUDECOMP_REF_CASES points at a larger set from oracle/make_ref_testfiles.py
CASES = os.environ.get("CUDECOMP_REF_CASES") or os.path.join(ROOT, "tests", "golden", "ref_cases")
with open(os.path.join(CASES, "index.n4.json")) as f:
    INDEX = json.load(f)

# (config, dtype) pairs: every configuration with its first dtype, plus all four dtypes of the two base sweeps
RUNS = [(name, dt) for name, info in sorted(INDEX.items()) for dt in info["dtypes"]]


def device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def run_mpi(nranks, argv, timeout, extra_env=None):
    port = free_port()
    procs = []
    ndev = device_count()
    for r in range(nranks):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(nranks), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        if 0 < ndev < nranks:
            # executables that call cudaSetDevice(local rank): report "nodes" of ndev ranks so the ranks share the GPUs
            env.setdefault("CUDECOMP_B200_SHIM_RANKS_PER_NODE", str(ndev))
        env.update(extra_env or {})
        # rank 0 reports; the other ranks' output is kept for debugging when CUDECOMP_REF_LOG_DIR is set
        log_dir = os.environ.get("CUDECOMP_REF_LOG_DIR")
        sink = subprocess.DEVNULL
        if r > 0 and log_dir:
            os.makedirs(log_dir, exist_ok=True)
            sink = open(os.path.join(log_dir, "%s.rank%d.log" % (os.path.basename(argv[0]), r)), "a")
        procs.append(subprocess.Popen(argv, env=env, stdout=subprocess.PIPE if r == 0 else sink,
                                      stderr=subprocess.STDOUT, text=True, start_new_session=True))
    try:
        out, _ = procs[0].communicate(timeout=timeout)
        codes = [procs[0].returncode] + [p.wait(timeout=60) for p in procs[1:]]
    except subprocess.TimeoutExpired:
        out, codes = "TIMEOUT", [-1]
    finally:
        for p in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, signal.SIGKILL)
                except ProcessLookupError:
                    pass
                p.wait()
    return out, codes


@pytest.mark.parametrize("config,dtype", RUNS, ids=["%s-%s" % r for r in RUNS])
def test_reference_executable(config, dtype):
    info = INDEX[config]
    exe = os.path.join(REF_BIN, "%s_%s" % (info["executable"], dtype))
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref binaries not built (make -f oracle/ref_tests.mk needs /root/reference)")
    out, codes = run_mpi(info["nranks"], [exe, "--testfile", os.path.join(CASES, info["file"])], timeout=1500)
    log_dir = os.environ.get("CUDECOMP_REF_LOG_DIR")
    if log_dir:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, "%s-%s.rank0.log" % (config, dtype)), "w") as f:
            f.write(out)
    errors = [l for l in out.splitlines() if "CUDECOMP:ERROR" in l or "FAILED" in l][:8]
    tail = "\n".join(errors + out.splitlines()[-6:])
    assert all(c == 0 for c in codes), "%s\n%s" % (codes, tail)
    assert "Passed all tests." in out, tail
    assert "Running %d tests" % info["kept"] in out

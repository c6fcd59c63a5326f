"""The reference's CURRENT GoogleTest suites -- tests/ctest/api_tests.cc and tests/ctest/halo_tests.cc with their own
support files and MPI-aware main() -- compiled UNMODIFIED against this library (oracle/ref_tests.mk ->
oracle/_ref/ctest_*; GoogleTest itself comes from oracle/gtest_shim, pinned by tests/test_gtest_shim.py) and run the way
the reference's CTest does (tests/ctest/CMakeLists.txt:102-196): 4 ranks, the halo suite once per backend label.

The reference's harness refuses to let ranks share a GPU unless CUDA MPS is running (gpu_test_utils.cc:75-84: its
MPI/NCCL backends deadlock under time slicing). This library's ranks may share a device, so on boxes with fewer than
4 GPUs the test points CUDA_MPS_PIPE_DIRECTORY at a directory holding the pid file that harness looks for; the
NCCL-labelled cases then skip themselves (they additionally want NCCL >= 2.30 for multi-rank-per-GPU)."""
import os
import re
import tempfile

import pytest

from tests.test_ref_executables_gpu import REF_BIN, run_mpi

pytestmark = [pytest.mark.gpu]

RUNS = [
    ("api", "ctest_api_tests", [], {}),
    ("halo_mpi", "ctest_halo_tests", ["--gtest_filter=MpiBackends/*"], {}),
    ("halo_nccl", "ctest_halo_tests", ["--gtest_filter=NcclBackends/*"], {"CUDECOMP_TEST_KEEPALIVE_BACKEND": "nccl"}),
]


@pytest.mark.parametrize("label,exe,args,env", RUNS, ids=[r[0] for r in RUNS])
def test_reference_gtest_suite(label, exe, args, env, monkeypatch):
    import torch
    path = os.path.join(REF_BIN, exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref binaries not built (make -f oracle/ref_tests.mk needs /root/reference)")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if torch.cuda.device_count() < 4:
        mps = tempfile.mkdtemp(prefix="cdb200_mps_")
        open(os.path.join(mps, "nvidia-cuda-mps-control.pid"), "w").write("0\n")
        monkeypatch.setenv("CUDA_MPS_PIPE_DIRECTORY", mps)
    out, codes = run_mpi(4, [path] + args, timeout=420)
    failed = re.findall(r"^\[  FAILED  \] (\S+)$", out, re.M)
    assert all(c == 0 for c in codes) and not failed, "%s\n%s\n%s" % (codes, failed[:10], out[-3000:])
    m = re.search(r"\[  PASSED  \] (\d+) tests", out)
    ran = re.search(r"\[==========\] (\d+) tests ran", out)
    assert m and ran and int(ran.group(1)) > 0, out[-2000:]
    skipped = re.search(r"\[  SKIPPED \] (\d+) tests", out)
    print("%s: %s ran, %s passed, %s skipped" % (label, ran.group(1), m.group(1), skipped.group(1) if skipped else 0))
    if label != "halo_nccl":
        assert int(m.group(1)) > 0, out[-2000:]

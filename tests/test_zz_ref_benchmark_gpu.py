"""The reference's own FFT benchmark (benchmark/benchmark.cu: cuFFT per pencil + the four transposes, with its own
max-error check against the input) built UNMODIFIED against this library by oracle/ref_tests.mk and run on 4 ranks.
SURVEY.md section 8(f) row 1. The binary picks its device as cudaSetDevice(local rank of
MPI_COMM_TYPE_SHARED) (benchmark.cu:135-139); on a box with fewer GPUs than ranks run_mpi tells the MPI shim to report
"nodes" of device_count ranks (as a real launcher with one rank per GPU would), so every rank gets a valid device."""
import os
import re

import pytest

from tests.test_ref_executables_gpu import REF_BIN, run_mpi

pytestmark = [pytest.mark.gpu]

RUNS = [
    ("benchmark_c2c", ["--gx", "64", "--gy", "64", "--gz", "64", "-r", "2", "-c", "2", "-b", "4", "-t", "2", "-w", "1"]),
    ("benchmark_c2c", ["--gx", "64", "--gy", "48", "--gz", "40", "-r", "2", "-c", "2", "-b", "4", "-o", "-t", "2", "-w", "1",
                       "--acx", "1", "--acy", "1", "--acz", "1"]),
    ("benchmark_r2c", ["--gx", "64", "--gy", "64", "--gz", "64", "-r", "2", "-c", "2", "-b", "1", "-t", "2", "-w", "1"]),
    ("benchmark_c2c_f", ["--gx", "64", "--gy", "64", "--gz", "64", "-r", "4", "-c", "1", "-b", "6", "-t", "2", "-w", "1"]),
    ("benchmark_r2c_f", ["--gx", "64", "--gy", "64", "--gz", "64", "-r", "1", "-c", "4", "-b", "4", "-o", "-t", "2", "-w", "1"]),
    ("benchmark_c2c", ["--gx", "64", "--gy", "64", "--gz", "64", "-r", "0", "-c", "0", "-b", "0", "-t", "2", "-w", "1"]),
]


@pytest.mark.parametrize("exe,args", RUNS, ids=["%s-%d" % (e, i) for i, (e, _) in enumerate(RUNS)])
def test_reference_fft_benchmark(exe, args):
    path = os.path.join(REF_BIN, exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref binaries not built (make -f oracle/ref_tests.mk needs /root/reference)")
    out, codes = run_mpi(4, [path] + args, timeout=300)
    assert all(c == 0 for c in codes), out[-3000:]
    assert "Result Summary:" in out and "FAILURE" not in out, out[-3000:]
    m = re.search(r"Max error: ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) < (5e-4 if exe.endswith("_f") else 1e-10), out[-1500:]

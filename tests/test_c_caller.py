"""The boundary from actual C: tests/c_caller/basic_usage.c is compiled with gcc -std=c11 against include/cudecomp.h
and linked with libcudecomp.so, then run on one rank (geometry queries and argument checking only, no GPU needed)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller_compiles_links_and_runs(tmp_path):
    exe = str(tmp_path / "basic_usage_c")
    lib_dir = os.path.join(ROOT, "cudecomp_b200", "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "include", "mpi_shim"), "-I" + cuda_inc,
           os.path.join(ROOT, "tests", "c_caller", "basic_usage.c"), "-L" + lib_dir, "-lcudecomp",
           "-Wl,-rpath," + lib_dir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE")}
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert run.returncode == 0 and "C caller OK" in run.stdout, run.stdout + run.stderr

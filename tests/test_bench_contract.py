"""bench.py contract checks that need no GPU: the reference arm (CPU oracle on the host cores) prints one JSON line
with the keys the driver reads, and the native arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "128",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic"
    # the whole 128^3 grid fits any host: the config object is then exactly what the native arm prints for this
    # workload (no "sample" key), and the sample description says "the whole grid"
    assert "workload" in d["config"] and "sample" not in d["config"]
    assert d["cpu_baseline"]["sample"].startswith("the whole grid")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--grid", "64", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--grid", "64", "--steps", "1", "--warmup", "0",
                          "--no-e2e", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0  # no silent CPU fallback
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_the_bench_parity_pattern_is_the_oracles_known_answer():
    """bench.py checks every operation's output against index_pattern() on the device; here the same function on CPU
    tensors is held against the oracle's restatement of the reference tests' generator (oracle.pattern_pencil) for
    default, axis-contiguous and explicit memory orders, even and uneven splits, all element sizes."""
    import importlib.util
    import numpy as np
    import torch
    from oracle import oracle as orc
    spec = importlib.util.spec_from_file_location("bench_main2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cases = [([12, 10, 8], [2, 2], (False,) * 3, None), ([9, 10, 11], [2, 2], (True,) * 3, None),
             ([7, 6, 5], [1, 4], (False, True, False), None), ([8, 8, 8], [2, 2], (False,) * 3, [[1, 0, 2], [2, 1, 0], [0, 2, 1]])]
    for gdims, pdims, ac, mo in cases:
        o = orc.Oracle(gdims, pdims, ac, mo)
        for rank in range(o.nranks):
            for ax in range(3):
                p = o.pencil_info(rank, ax)
                want = orc.pattern_pencil(p, gdims, np.int64)
                for es in (4, 8, 16):
                    got = bench.index_pattern(torch, p, gdims, es, "cpu").numpy()
                    if es == 16:
                        assert np.array_equal(got[0::2], want) and np.array_equal(got[1::2], ~want)
                    else:
                        assert np.array_equal(got.astype(np.int64), want), (gdims, pdims, ac, rank, ax, es)


def test_cpu_sample_is_the_whole_grid_when_it_fits(monkeypatch):
    """The CPU arm times the benchmark's own configuration when the host can hold it, else a z-slab (bench.cpu_sample_nz)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_main3", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class A:
        n, dtype = 1024, "double_complex"
    monkeypatch.setattr(bench, "mem_available_bytes", lambda: 251 << 30)
    assert bench.cpu_sample_nz(A, (1, 1), budget_s=180.0, steps=13) == 1024
    monkeypatch.setattr(bench, "mem_available_bytes", lambda: 60 << 30)  # 4.5 grids of 17 GB do not fit
    nz = bench.cpu_sample_nz(A, (2, 4), budget_s=180.0, steps=13)
    assert 16 <= nz < 1024 and nz & (nz - 1) == 0
    monkeypatch.setattr(bench, "mem_available_bytes", lambda: 251 << 30)
    assert bench.cpu_sample_nz(A, (2, 4), budget_s=5.0, steps=3) < 1024  # time budget too small for the whole grid


def test_the_halo_benchmark_patterns_are_the_oracles_known_answers():
    """bench/halo_benchmark.py checks its runs against halo_pattern() on the device; on CPU tensors it must equal the
    oracle's restatement of the reference's initializePencil / initializeReference (tests/ctest/halo_tests.cc:197-236),
    periodic and not, for several layouts and element sizes."""
    import importlib.util
    import numpy as np
    import torch
    from oracle import oracle as orc
    spec = importlib.util.spec_from_file_location("halo_bench", os.path.join(ROOT, "bench", "halo_benchmark.py"))
    hb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hb)
    halo = [2, 1, 2]
    for gdims, pdims, ac in (([12, 10, 8], [2, 2], (False,) * 3), ([9, 10, 11], [1, 4], (True,) * 3), ([8, 9, 7], [4, 1], (False,) * 3)):
        o = orc.Oracle(gdims, pdims, ac)
        for periods in ([True] * 3, [False, True, False]):
            for rank in range(o.nranks):
                for ax in range(3):
                    p = o.pencil_info(rank, ax, halo)
                    start = orc.pattern_pencil(p, gdims, np.int64)
                    after = orc.halo_reference(p, gdims, np.int64, periods)
                    for es in (4, 8, 16):
                        got0 = hb.halo_pattern(torch, p, gdims, halo, periods, es, "cpu", False).numpy()
                        got1 = hb.halo_pattern(torch, p, gdims, halo, periods, es, "cpu", True).numpy()
                        if es == 16:
                            assert np.array_equal(got0[0::2], start) and np.array_equal(got1[0::2], after)
                            assert np.array_equal(got1[1::2], ~after)
                        else:
                            assert np.array_equal(got0.astype(np.int64), start), (gdims, pdims, rank, ax, es)
                            assert np.array_equal(got1.astype(np.int64), after), (gdims, pdims, periods, rank, ax, es)

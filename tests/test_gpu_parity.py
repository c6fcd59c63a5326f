"""Parity of the CUDA path with the oracle, through the C ABI, on real GPUs (pytest -m gpu).

The case matrix is the reference's own (tests/ctest/transpose_tests.cc:163-273, halo_tests.cc:103-146,
tests/test_runner.py:80-90). Every case runs on N ranks = N processes; with fewer GPUs than ranks the ranks share
a GPU. Each rank checks its result (a) against the analytic global-index pattern -- the reference's own pass/fail
criterion -- and (b) byte for byte against the CPU oracle run on the same seeded inputs, including the cells the
operation must leave untouched (output halos, padding).
"""
import pytest

from tests import cases as C
from tests._launcher import run_ranks

pytestmark = pytest.mark.gpu


def _run(nranks, cases, timeout=1500):
    results, logs = run_ranks(nranks, "gpu", cases, timeout=timeout)
    return results, logs


def _assert_case(results, i, case):
    msgs = ["rank %d: %s" % (r, results[r][i].get("msg")) for r in range(len(results)) if not results[r][i]["ok"]]
    assert not msgs, "%s\n%s" % (case["name"], "\n".join(msgs))


# ------------------------------------------------------------------------------------------- single rank (1x1)
SINGLE = C.transpose_single_rank() + C.halo_baseline((1, 1)) + C.legacy_mem_order_chain((1, 1), stride=4) + \
    C.legacy_mem_order_chain((1, 1), dtype="double_complex", out_of_place=True, stride=7)


@pytest.fixture(scope="module")
def single_results():
    return _run(1, SINGLE)[0]


@pytest.mark.parametrize("i", range(len(SINGLE)), ids=[c["name"] for c in SINGLE])
def test_single_rank(single_results, i):
    _assert_case(single_results, i, SINGLE[i])


# ------------------------------------------------------------------------------------------------- 4 ranks, 2x2
FOUR = C.transpose_baseline((2, 2)) + C.transpose_coverage_2x2() + C.halo_baseline((2, 2)) + C.halo_coverage() + \
    C.legacy_mem_order_chain((2, 2), stride=3) + \
    C.legacy_mem_order_chain((2, 2), dtype="double_complex", out_of_place=True, stride=5) + [
        dict(kind="transpose", name="Slab4x1_chain", gdims=[16, 12, 20], pdims=[4, 1], dtype="double",
             ops=["XY", "YZ", "ZY", "YX"], out_of_place=True),
        dict(kind="transpose", name="Slab1x4_chain_inplace", gdims=[16, 12, 20], pdims=[1, 4], dtype="float_complex",
             ops=["XY", "YZ", "ZY", "YX"]),
        dict(kind="transpose", name="NativeAlltoAllPath_YZ_1x4", gdims=[8, 8, 8], pdims=[1, 4], dtype="float",
             op="YZ", out_of_place=True),
        dict(kind="transpose", name="EmptyPencils_NotSupported", gdims=[2, 2, 2], pdims=[4, 1], dtype="float", op="XY",
             expect=2, fills=["pattern"]),
        dict(kind="halo", name="HaloTooWide_InvalidUsage", gdims=C.GDIMS, pdims=[2, 2], dtype="float", axis=0,
             halo=[0, 6, 0], periods=[True] * 3, expect=1, fills=["pattern"], dims=[1]),
        dict(kind="transpose", name="FewCtas_XY", gdims=[40, 36, 44], pdims=[2, 2], dtype="double", op="XY",
             out_of_place=True, grid_ctas=3),
        # extents that keep every row 16-byte aligned: the vectorised transpose kernel (transposeVecKernel) runs
        dict(kind="transpose", name="VecTranspose_float_axis_contiguous_oop", gdims=[136, 72, 64], pdims=[2, 2],
             dtype="float", ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, axis_contiguous=[True] * 3),
        dict(kind="transpose", name="VecTranspose_double_axis_contiguous_inplace", gdims=[64, 72, 136], pdims=[2, 2],
             dtype="double", ops=["XY", "YZ", "ZY", "YX"], axis_contiguous=[True] * 3),
        dict(kind="transpose", name="VecTranspose_c128_mem_order_halo", gdims=[48, 64, 40], pdims=[2, 2],
             dtype="double_complex", ops=["XY", "YZ", "ZY", "YX"], out_of_place=True,
             mem_order=[[1, 2, 0], [2, 0, 1], [0, 1, 2]],
             halos={"0": [1, 1, 1], "1": [1, 1, 1], "2": [1, 1, 1]}),
        dict(kind="transpose", name="VecTranspose_float_halo4_pairwise", gdims=[72, 64, 48], pdims=[4, 1], dtype="float",
             ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, axis_contiguous=[True] * 3, peer_order=1,
             halos={"0": [4, 4, 4], "1": [4, 4, 4], "2": [4, 4, 4]}),
        dict(kind="transpose", name="ElementwiseTranspose_variant3_float", gdims=[136, 72, 64], pdims=[2, 2],
             dtype="float", ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, axis_contiguous=[True] * 3, kernel_variant=3),
        dict(kind="autotune", name="AutotuneTransposeGrid", gdims=[24, 20, 28], dtype="double", n_trials=2),
        dict(kind="autotune", name="AutotuneTransposeBackend", gdims=[24, 20, 28], dtype="float_complex",
             autotune_backend=True, n_trials=1),
        dict(kind="autotune", name="AutotuneHaloGrid", gdims=[24, 20, 28], dtype="float", grid_mode=1,
             halo=[1, 1, 1], n_trials=1),
    ]


@pytest.fixture(scope="module")
def four_results():
    return _run(4, FOUR)[0]


@pytest.mark.parametrize("i", range(len(FOUR)), ids=[c["name"] for c in FOUR])
def test_four_ranks(four_results, i):
    _assert_case(four_results, i, FOUR[i])


def test_direct_and_staged_paths_are_both_exercised(four_results):
    """Out-of-place exchanges with exportable buffers take the one-kernel direct path (2); in-place ones and the
    NVSHMEM backend values stage through the workspace (3)."""
    by_name = {c["name"]: four_results[0][i] for i, c in enumerate(FOUR)}
    assert set(by_name["BaselineDefaultLayout_XY_float_P2x2_OutOfPlace"]["paths"]) == {2}
    assert set(by_name["BaselineDefaultLayout_XY_float_P2x2_InPlace"]["paths"]) == {3}
    assert set(by_name["StagedBackend_XY"]["paths"]) == {3}
    assert set(by_name["ForceStaged_YZ_AxisContiguous"]["paths"]) == {3}
    assert 2 in set(by_name["BaselineDefaultLayoutPeriodic_Axis0_float_P2x2"]["paths"])
    assert set(by_name["StagedBackend"]["paths"]) <= {1, 3}


# ------------------------------------------------------------------------------------- 3 ranks (non power of two)
THREE = C.transpose_coverage_3x1() + C.halo_3x1()


@pytest.fixture(scope="module")
def three_results():
    return _run(3, THREE)[0]


@pytest.mark.parametrize("i", range(len(THREE)), ids=[c["name"] for c in THREE])
def test_three_ranks(three_results, i):
    _assert_case(three_results, i, THREE[i])


# ---------------------------------------------------------------- 2 ranks: BASELINE.json config 1 (128^3 double)
TWO = [
    dict(kind="transpose", name="Config1_128cubed_double_1x2_%s" % ("oop" if oop else "inplace"), gdims=[128] * 3,
         pdims=[1, 2], dtype="double", ops=["XY", "YZ", "ZY", "YX"], out_of_place=oop)
    for oop in (False, True)
] + [
    dict(kind="transpose", name="Config1_128cubed_double_2x1_%s" % ("oop" if oop else "inplace"), gdims=[128] * 3,
         pdims=[2, 1], dtype="double", ops=["XY", "YZ", "ZY", "YX"], out_of_place=oop)
    for oop in (False, True)
] + [
    dict(kind="transpose", name="Uneven_130x126x134_c128_2x1_axis_contiguous", gdims=[130, 126, 134], pdims=[2, 1],
         dtype="double_complex", ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, axis_contiguous=[True] * 3,
         fills=["random"]),
    dict(kind="halo", name="Halo_128x132x124_1x2_axis0", gdims=[128, 132, 124], pdims=[1, 2], dtype="float", axis=0,
         halo=[2, 2, 2], periods=[True] * 3, fills=["random"]),
    dict(kind="halo", name="Halo_128x132x124_2x1_axis2_nonperiodic", gdims=[128, 132, 124], pdims=[2, 1],
         dtype="double_complex", axis=2, halo=[1, 2, 1], periods=[False] * 3, fills=["random"]),
]


@pytest.fixture(scope="module")
def two_results():
    return _run(2, TWO)[0]


@pytest.mark.parametrize("i", range(len(TWO)), ids=[c["name"] for c in TWO])
def test_two_ranks(two_results, i):
    _assert_case(two_results, i, TWO[i])


# ------------------------------------------------------------------------- the transposes inside their caller
def _run_script(nranks, argv, timeout=600):
    import os
    import signal
    import subprocess
    import sys
    from tests._launcher import ROOT, free_port
    port = free_port()
    procs = []
    for r in range(nranks):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(nranks), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable] + argv, env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE if r == 0 else subprocess.DEVNULL,
                                      stderr=subprocess.STDOUT, text=True, start_new_session=True))
    try:
        out, _ = procs[0].communicate(timeout=timeout)
        codes = [procs[0].returncode] + [p.wait(timeout=60) for p in procs[1:]]
    finally:
        for p in procs:
            if p.poll() is None:
                try:
                    os.killpg(p.pid, signal.SIGKILL)
                except ProcessLookupError:
                    pass
                p.wait()
    return out, codes


@pytest.mark.parametrize("layout", [[], ["--axis-contiguous"], ["--inplace"], ["--inplace", "--axis-contiguous"]],
                         ids=["default", "axis_contiguous", "inplace", "inplace_axis_contiguous"])
def test_distributed_fft_matches_numpy_fftn(layout):
    """benchmark/benchmark.cu's sequence (FFT per pencil + 4 transposes) on 4 ranks: forward transform equals
    numpy.fft.fftn of the whole field, forward + backward reproduces the input (tolerance 1e-10, benchmark.cu:21-27)."""
    import json
    out, codes = _run_script(4, ["bench/fft_benchmark.py", "--grid", "48", "40", "36", "--check-global", "--steps", "1",
                                 "--warmup", "1"] + layout)
    assert all(c == 0 for c in codes), out[-3000:]
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line["passed"] and line["max_roundtrip_error"] <= 1e-10 and line["global_fftn_rel_error"] < 1e-12


# ------------------------------------------------------------------ 8 ranks: the 2x4 grid of the headline benchmark
EIGHT = [
    dict(kind="transpose", name="Grid2x4_chain_c128_oop", gdims=[32, 40, 48], pdims=[2, 4], dtype="double_complex",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True),
    dict(kind="transpose", name="Grid2x4_chain_c64_inplace", gdims=[32, 40, 48], pdims=[2, 4], dtype="float_complex",
         ops=["XY", "YZ", "ZY", "YX"]),
    dict(kind="transpose", name="Grid2x4_uneven_axis_contiguous", gdims=[30, 29, 35], pdims=[2, 4], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, axis_contiguous=[True] * 3),
    dict(kind="transpose", name="Grid4x2_halo_padding_float", gdims=[30, 29, 35], pdims=[4, 2], dtype="float",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True,
         halos={"0": [1, 1, 1], "1": [1, 2, 1], "2": [2, 1, 1]}, pads={"0": [1, 0, 0], "1": [0, 1, 0], "2": [0, 0, 2]}),
    dict(kind="transpose", name="Slab1x8_chain", gdims=[24, 32, 40], pdims=[1, 8], dtype="double", out_of_place=True,
         ops=["XY", "YZ", "ZY", "YX"]),
    dict(kind="transpose", name="Slab8x1_chain_inplace", gdims=[24, 32, 40], pdims=[8, 1], dtype="float_complex",
         ops=["XY", "YZ", "ZY", "YX"]),
    dict(kind="halo", name="Halo2x4_axis0_periodic", gdims=[32, 40, 48], pdims=[2, 4], dtype="float", axis=0,
         halo=[2, 2, 2], periods=[True] * 3),
    dict(kind="halo", name="Halo1x8_axis2_nonperiodic_c128", gdims=[32, 40, 48], pdims=[1, 8], dtype="double_complex",
         axis=2, halo=[1, 2, 1], periods=[False] * 3),
    dict(kind="autotune", name="Autotune8", gdims=[32, 40, 48], dtype="double", n_trials=1, autotune_backend=True),
]


@pytest.fixture(scope="module")
def eight_results():
    return _run(8, EIGHT)[0]


@pytest.mark.parametrize("i", range(len(EIGHT)), ids=[c["name"] for c in EIGHT])
def test_eight_ranks(eight_results, i):
    _assert_case(eight_results, i, EIGHT[i])

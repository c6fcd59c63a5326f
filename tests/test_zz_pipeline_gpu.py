"""Chunked schedules of staged transposes on the GPU (cudecompB200SetPipelineChunks / cudecompB200SetStagedMode).

The schedule itself -- which piece may be unpacked after which push, including the in-place hazard analysis -- is
property-tested on the host against the oracle (tests/test_planner_properties.py, tests/test_launch_emulation.py).
Two device executions exist: the default, ONE phased launch with per-chunk flags (kernels.cu rowCopyPhasedKernel,
engine.cc runFusedStaged), and the round-1 schedule of K handshaking push launches with unpacks on a side stream
(staged_mode=1, also what layouts with differing memory orders use). Both are run here.
"""
import pytest

from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu]

CASES = [
    dict(kind="transpose", name="Pipe4_inplace_2x2_default", gdims=[32, 40, 48], pdims=[2, 2], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"], pipeline_chunks=4),
    dict(kind="transpose", name="Pipe3_inplace_2x2_uneven_c128", gdims=[30, 29, 35], pdims=[2, 2], dtype="double_complex",
         ops=["XY", "YZ", "ZY", "YX"], pipeline_chunks=3),
    dict(kind="transpose", name="Pipe4_staged_oop_axis_contiguous", gdims=[32, 40, 48], pdims=[2, 2], dtype="float",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, force_staged=True, axis_contiguous=[True] * 3, pipeline_chunks=4),
    dict(kind="transpose", name="Pipe8_inplace_4x1_halo_padding", gdims=[24, 32, 40], pdims=[4, 1], dtype="float_complex",
         ops=["XY", "YZ", "ZY", "YX"], pipeline_chunks=8,
         halos={"0": [1, 1, 1], "1": [1, 1, 1], "2": [1, 1, 1]}, pads={"0": [1, 0, 0], "1": [0, 1, 0], "2": [0, 0, 2]}),
    dict(kind="transpose", name="Pipe2_inplace_1x4", gdims=[24, 32, 40], pdims=[1, 4], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"], pipeline_chunks=2),
]
# every case above in the round-1 execution (separate launches), and the fused one with the shortest and a long lag
CASES = CASES + [dict(c, name=c["name"] + "_launches", staged_mode=1) for c in CASES] + [
    dict(CASES[0], name="Fused_lag1", fused_lag=1), dict(CASES[1], name="Fused_lag3_uneven", fused_lag=3),
    dict(CASES[3], name="Fused_lag1_halo_padding", fused_lag=1),
    dict(kind="transpose", name="Fused_auto_chunks_inplace_2x2_c128", gdims=[256, 128, 96], pdims=[2, 2],
         dtype="double_complex", ops=["XY", "YZ", "ZY", "YX"] * 2),
    dict(kind="transpose", name="Fused_16chunks_inplace_1x4_f32", gdims=[128, 96, 256], pdims=[1, 4], dtype="float",
         ops=["XY", "YZ", "ZY", "YX"] * 2, pipeline_chunks=16),
    dict(kind="transpose", name="Fused_oop_forced_staging_4x1", gdims=[96, 128, 64], pdims=[4, 1], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, force_staged=True, pipeline_chunks=5),
    # rows of 8 KiB: Y<->Z is chunked along x (column chunks, plan.cc), X<->Y along z
    dict(kind="transpose", name="Fused_column_chunks_inplace_2x2_c128", gdims=[1024, 24, 20], pdims=[2, 2],
         dtype="double_complex", ops=["XY", "YZ", "ZY", "YX"] * 2, pipeline_chunks=4),
    dict(kind="transpose", name="Fused_column_chunks_inplace_1x4_uneven_f64", gdims=[1030, 18, 22], pdims=[1, 4],
         dtype="double", ops=["XY", "YZ", "ZY", "YX"], pipeline_chunks=3),
]


@pytest.fixture(scope="module")
def pipe_results():
    return run_ranks(4, "gpu", CASES, timeout=420)[0]


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_pipelined_staged_transposes(pipe_results, i):
    bad = ["rank %d: %s" % (r, pipe_results[r][i].get("msg")) for r in range(4) if not pipe_results[r][i]["ok"]]
    assert not bad, "\n".join(bad)
    assert 3 in set(pipe_results[0][i]["paths"])  # the staged path really ran


# --------------------------------------------------------- back-to-back operations without host synchronisation
STRESS = [
    dict(kind="stress", name="Stress_fused_inplace_2x2", gdims=[128, 96, 64], pdims=[2, 2], dtype="double_complex", reps=12,
         pipeline_chunks=8),
    dict(kind="stress", name="Stress_fused_inplace_lag1_1x4", gdims=[96, 64, 128], pdims=[1, 4], dtype="double", reps=12,
         pipeline_chunks=6, fused_lag=1),
    dict(kind="stress", name="Stress_fused_inplace_auto_4x1_uneven", gdims=[101, 99, 67], pdims=[4, 1], dtype="float", reps=12),
    dict(kind="stress", name="Stress_fused_column_chunks_2x2", gdims=[1024, 32, 24], pdims=[2, 2], dtype="double_complex",
         reps=12, pipeline_chunks=8),
    dict(kind="stress", name="Stress_launches_inplace_2x2", gdims=[128, 96, 64], pdims=[2, 2], dtype="double", reps=6,
         pipeline_chunks=4, staged_mode=1),
    dict(kind="stress", name="Stress_direct_oop_2x2", gdims=[128, 96, 64], pdims=[2, 2], dtype="double_complex", reps=12,
         out_of_place=True),
    dict(kind="stress", name="Stress_mixed_direct_then_staged_2x2", gdims=[64, 96, 128], pdims=[2, 2], dtype="double", reps=6,
         out_of_place=True, force_staged=True, pipeline_chunks=3),
]


@pytest.fixture(scope="module")
def stress_results():
    return run_ranks(4, "gpu", STRESS, timeout=420)[0]


@pytest.mark.parametrize("i", range(len(STRESS)), ids=[c["name"] for c in STRESS])
def test_back_to_back_round_trips(stress_results, i):
    bad = ["rank %d: %s" % (r, stress_results[r][i].get("msg")) for r in range(4) if not stress_results[r][i]["ok"]]
    assert not bad, "\n".join(bad)


# ------------------------------------------------------------------------------------ TMA bulk row-copy variant
BULK_CASES = [
    dict(kind="transpose", name="Bulk_oop_2x2_c128", gdims=[256, 64, 48], pdims=[2, 2], dtype="double_complex",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, kernel_variant=1),
    dict(kind="transpose", name="Bulk_inplace_2x2_double", gdims=[512, 48, 40], pdims=[2, 2], dtype="double",
         ops=["XY", "YZ", "ZY", "YX"], kernel_variant=1),
    dict(kind="transpose", name="Bulk_falls_back_on_short_rows", gdims=[9, 10, 11], pdims=[2, 2], dtype="float",
         ops=["XY", "YZ", "ZY", "YX"], out_of_place=True, kernel_variant=1),
    dict(kind="transpose", name="Bulk_pipelined_inplace", gdims=[256, 64, 48], pdims=[2, 2], dtype="double_complex",
         ops=["XY", "YZ", "ZY", "YX"], kernel_variant=1, pipeline_chunks=4),
]


@pytest.fixture(scope="module")
def bulk_results():
    return run_ranks(4, "gpu", BULK_CASES, timeout=420)[0]


@pytest.mark.parametrize("i", range(len(BULK_CASES)), ids=[c["name"] for c in BULK_CASES])
def test_tma_bulk_variant(bulk_results, i):
    bad = ["rank %d: %s" % (r, bulk_results[r][i].get("msg")) for r in range(4) if not bulk_results[r][i]["ok"]]
    assert not bad, "\n".join(bad)

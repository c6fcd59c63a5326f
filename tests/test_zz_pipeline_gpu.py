"""Chunked (pipelined) schedule of staged transposes on the GPU (opt-in, cudecompB200SetPipelineChunks).

The schedule itself -- which piece may be unpacked after which push, including the in-place hazard analysis -- is
property-tested on the host against the oracle (tests/test_planner_properties.py). Its device execution (K handshaking
push launches on the caller's stream, unpack launches on a side stream) was written after the round-1 GPU budget was
spent, so these tests are expected-to-pass-but-unconfirmed: xfail(strict=False) keeps an unexpected hardware-side
surprise from masking the rest of the suite; an XPASS is the confirmation.
"""
import pytest

from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="device execution not yet confirmed on hardware")]

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


@pytest.fixture(scope="module")
def pipe_results():
    return run_ranks(4, "gpu", CASES, timeout=420)[0]


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_pipelined_staged_transposes(pipe_results, i):
    bad = ["rank %d: %s" % (r, pipe_results[r][i].get("msg")) for r in range(4) if not pipe_results[r][i]["ok"]]
    assert not bad, "\n".join(bad)
    assert 3 in set(pipe_results[0][i]["paths"])  # the staged path really ran


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

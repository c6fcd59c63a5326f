"""Opt-in launch schedules on the GPU (cudecompB200SetSchedule): tile sizes, the pairwise slot order and the balanced
grid must give byte-identical results to the default schedule. Their index arithmetic is property-tested on the host
(tests/test_launch_emulation.py walks the same launches with the kernels' own decode functions); confirmed on B200 by the
round-1 driver run (GPUTEST_r01.json)."""
import pytest

from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu]

BASE = dict(kind="transpose", ops=["XY", "YZ", "ZY", "YX"])
CASES = [
    dict(BASE, name="Pairwise_oop_2x2_c128", gdims=[64, 40, 48], pdims=[2, 2], dtype="double_complex", out_of_place=True,
         peer_order=1),
    dict(BASE, name="Pairwise_oop_1x4_uneven_float", gdims=[31, 30, 29], pdims=[1, 4], dtype="float", out_of_place=True,
         peer_order=1),
    dict(BASE, name="Pairwise_inplace_4x1_axis_contiguous", gdims=[32, 40, 48], pdims=[4, 1], dtype="double",
         axis_contiguous=[True] * 3, peer_order=1),
    dict(BASE, name="Tile4k_oop_2x2_long_rows", gdims=[1030, 12, 10], pdims=[2, 2], dtype="double_complex",
         out_of_place=True, tile_bytes=4096),
    dict(BASE, name="Tile64k_balanced_inplace_2x2", gdims=[96, 40, 48], pdims=[2, 2], dtype="float_complex",
         tile_bytes=65536, balance_grid=1),
    dict(BASE, name="Pairwise_tile8k_balanced_halo_padding", gdims=[48, 32, 40], pdims=[2, 2], dtype="double",
         out_of_place=True, peer_order=1, tile_bytes=8192, balance_grid=1,
         halos={"0": [1, 1, 1], "1": [1, 1, 1], "2": [1, 1, 1]}, pads={"0": [1, 0, 0], "1": [0, 1, 0], "2": [0, 0, 2]}),
    dict(BASE, name="Pull_oop_2x2_c128", gdims=[64, 40, 48], pdims=[2, 2], dtype="double_complex", out_of_place=True, pull=1),
    dict(BASE, name="Pull_oop_4x1_uneven_axis_contiguous_halos", gdims=[31, 30, 29], pdims=[4, 1], dtype="float",
         out_of_place=True, pull=1, axis_contiguous=[True] * 3,
         halos={"0": [1, 1, 1], "1": [1, 0, 1], "2": [0, 1, 1]}, pads={"0": [1, 0, 0], "1": [0, 1, 0], "2": [0, 0, 2]}),
    dict(BASE, name="PullStaged_inplace_2x2", gdims=[32, 40, 48], pdims=[2, 2], dtype="double", pull=1),
    dict(BASE, name="PullStaged_inplace_1x4_c128_private_workspace", gdims=[24, 32, 40], pdims=[1, 4], dtype="double_complex",
         pull=1, work_alloc="torch"),  # nobody writes into a peer's workspace, so it need not come from cudecompMalloc
    dict(BASE, name="PullChunked4_inplace_2x2", gdims=[32, 40, 48], pdims=[2, 2], dtype="double", pull=1, pipeline_chunks=4),
    dict(BASE, name="PullChunked3_inplace_1x4_uneven_axis_contiguous", gdims=[30, 29, 35], pdims=[1, 4], dtype="float_complex",
         pull=1, pipeline_chunks=3, axis_contiguous=[True] * 3),
    dict(BASE, name="PullStaged_forced_oop_axis_contiguous", gdims=[30, 29, 35], pdims=[2, 2], dtype="float",
         out_of_place=True, force_staged=True, pull=1, axis_contiguous=[True] * 3),
    dict(BASE, name="Wide256_oop_2x2_c128", gdims=[64, 40, 48], pdims=[2, 2], dtype="double_complex", out_of_place=True,
         kernel_variant=2),
    dict(BASE, name="Wide256_inplace_2x2_uneven_float_falls_back", gdims=[31, 30, 29], pdims=[2, 2], dtype="float",
         kernel_variant=2),
    dict(BASE, name="Pairwise_bulk_oop_2x2", gdims=[256, 24, 20], pdims=[2, 2], dtype="double_complex", out_of_place=True,
         peer_order=1, kernel_variant=1),
]


@pytest.fixture(scope="module")
def schedule_results():
    return run_ranks(4, "gpu", CASES, timeout=420)[0]


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_opt_in_schedules(schedule_results, i):
    bad = ["rank %d: %s" % (r, schedule_results[r][i].get("msg")) for r in range(4) if not schedule_results[r][i]["ok"]]
    assert not bad, "\n".join(bad)

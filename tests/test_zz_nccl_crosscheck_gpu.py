"""Byte-for-byte cross-check of the library against the reference's NCCL arm restated with public NCCL only
(bench/nccl_restated.py: pack -> all_to_all_single -> unpack in the reference's wire format), SURVEY.md section 8(c).
The restated arm itself is pinned to the oracle on the CPU (tests/test_nccl_restated.py). NCCL needs one GPU per rank, so
the cases skip themselves on smaller boxes. Hardware runs: profiles/r2_n2_schedules.md (2 GPUs), profiles/r2_n8_*.txt (8 GPUs)."""
import pytest

from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu]

CASES = {
    2: [dict(kind="nccl_crosscheck", name="c128_1x2", gdims=[64, 48, 40], pdims=[1, 2], dtype="double_complex"),
        dict(kind="nccl_crosscheck", name="f32_2x1_uneven", gdims=[31, 30, 29], pdims=[2, 1], dtype="float")],
    4: [dict(kind="nccl_crosscheck", name="c128_2x2", gdims=[64, 48, 40], pdims=[2, 2], dtype="double_complex"),
        dict(kind="nccl_crosscheck", name="f64_2x2_axis_contiguous", gdims=[33, 30, 29], pdims=[2, 2], dtype="double",
             axis_contiguous=[True] * 3),
        dict(kind="nccl_crosscheck", name="c64_4x1_gdims_dist", gdims=[40, 36, 32], pdims=[4, 1], dtype="float_complex",
             gdims_dist=[38, 33, 32])],
    8: [dict(kind="nccl_crosscheck", name="c128_2x4", gdims=[128, 96, 64], pdims=[2, 4], dtype="double_complex"),
        dict(kind="nccl_crosscheck", name="c128_1x8", gdims=[64, 48, 80], pdims=[1, 8], dtype="double_complex")],
}


@pytest.mark.parametrize("nranks", sorted(CASES))
def test_library_equals_restated_nccl_arm(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    results, _ = run_ranks(nranks, "gpu", CASES[nranks], timeout=420)
    for r in range(nranks):
        for c, case in zip(results[r], CASES[nranks]):
            assert c["ok"], (case["name"], r, c.get("msg"))
            assert 2 in set(c["paths"]) or 1 in set(c["paths"])  # direct peer stores (or local when the communicator is 1 rank)

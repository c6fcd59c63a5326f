"""The reference's API contract tests (tests/ctest/api_tests.cc:571-1547) on 4 ranks, host only: argument validation,
struct versioning, multiple live handles, config round trips, shifted ranks, empty-pencil rejection. The per-rank
restatement lives in tests/_api_battery.py."""
import pytest

from tests._api_battery import TEST_NAMES
from tests._launcher import run_ranks


@pytest.fixture(scope="module")
def api_results():
    results, _ = run_ranks(4, "api", [dict(name=n) for n in TEST_NAMES], timeout=300)
    return results


@pytest.mark.parametrize("i", range(len(TEST_NAMES)), ids=TEST_NAMES)
def test_api_contract_on_4_ranks(api_results, i):
    for rank in range(4):
        r = api_results[rank][i]
        assert r.get("ok"), "rank %d: %s" % (rank, r.get("failures") or r.get("msg"))

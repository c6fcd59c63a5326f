"""The reference's API contract tests (tests/ctest/api_tests.cc) on 4 ranks WITH a device behind every handle: adds the
cudecompMalloc / cudecompFree parts of SupportsMultipleLiveHandlesWithIndependentResources (api_tests.cc:575-608) to
what tests/test_api_contract.py checks on the host. Confirmed on B200 by the round-1 driver run (GPUTEST_r01.json)."""
import pytest

from tests._api_battery import TEST_NAMES
from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def api_results():
    results, _ = run_ranks(4, "api_gpu", [dict(name=n) for n in TEST_NAMES], timeout=300)
    return results


@pytest.mark.parametrize("i", range(len(TEST_NAMES)), ids=TEST_NAMES)
def test_api_contract_on_4_ranks_with_devices(api_results, i):
    for rank in range(4):
        r = api_results[rank][i]
        assert r.get("ok"), "rank %d: %s" % (rank, r.get("failures") or r.get("msg"))

"""The GoogleTest stand-in that lets the reference's gtest suites compile unmodified (oracle/gtest_shim/gtest/gtest.h;
this image has no GoogleTest): its own behaviour is pinned here on the CPU with a self-test program -- registration of
TEST / TEST_F / TEST_P x INSTANTIATE_TEST_SUITE_P, fatal vs non-fatal assertions, skips, exceptions, listeners, the
filter syntax -- so that a green or red run of the reference's suites on the GPU box means what it says."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "gtest_shim")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("gtest_shim") / "selftest")
    res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + SHIM, os.path.join(SHIM, "selftest.cc"),
                          "-o", out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def run(exe, *args):
    res = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_full_run(exe):
    out = run(exe, "positional")
    ok = re.findall(r"^\[       OK \] (\S+)", out, re.M)
    failed = re.findall(r"^\[  FAILED  \] (\S+)$", out, re.M)
    skipped = re.findall(r"^\[  SKIPPED \] (\S+)$", out, re.M)
    assert "Plain.PassesWithEveryComparison" in ok and "Fixture.SeesSetUp" in ok and "Fixture.SeesSetUpAgain" in ok
    # 3 params x 2 bodies x 2 instantiations; names from the generator and the default (index) naming
    assert "First/Param.ValueMatchesLabel/one_0" in ok and "Second/Param.ValueMatchesLabel/2" in ok
    assert "First/Param.OddValuesFail/two_1" in ok and "First/Param.OddValuesFail/one_0" in failed
    assert set(skipped) == {"Plain.Skips"}
    expected_failed = {"Plain.NonFatalFailuresContinue", "Plain.FatalFailureReturns", "Plain.FatalInHelperOnlyLeavesTheHelper",
                       "Plain.ExceptionIsAFailure", "First/Param.OddValuesFail/one_0", "First/Param.OddValuesFail/three_2",
                       "Second/Param.OddValuesFail/0", "Second/Param.OddValuesFail/2"}
    assert set(failed) == expected_failed
    assert "Running 20 tests" in out and "[  PASSED  ] 11 tests." in out
    assert "REACHED_AFTER_NONFATAL" in out and "REACHED_AFTER_HELPER" in out and "NOT_REACHED" not in out
    assert "first" in out and "5 is odd" in out and "second" in out and "stops here" in out and "helper got 8" in out
    assert "never shown" not in out and "boom" in out and "not today" in out
    m = re.search(r"SETUPS=(\d+) TEARDOWNS=(\d+) LISTENER_FAILURES=(\d+) RC=(\d+) ARGC=(\d+)", out)
    # 4 non-fatal + 1 fatal + 1 fatal in helper + 1 exception + 4 odd parameters
    assert tuple(map(int, m.groups())) == (2, 2, 11, 1, 2)


def test_filter_and_listing(exe):
    out = run(exe, "--gtest_filter=First/*:Fixture.*-*OddValuesFail*")
    assert "Running 5 tests" in out and "[  PASSED  ] 5 tests." in out and "RC=0" in out
    listed = run(exe, "--gtest_list_tests", "--gtest_filter=Plain.*").split("\n")
    assert [l for l in listed if l.startswith("Plain.")] == [
        "Plain.PassesWithEveryComparison", "Plain.NonFatalFailuresContinue", "Plain.FatalFailureReturns",
        "Plain.FatalInHelperOnlyLeavesTheHelper", "Plain.Skips", "Plain.ExceptionIsAFailure"]

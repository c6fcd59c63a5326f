"""Autotune candidate filtering, restated from the reference's AutotuneCandidateFilterTest (tests/ctest/api_tests.cc:319-443):
environment inclusion / exclusion lists, family disable flags, process-grid ranges, malformed values. CPU only."""
import os
from contextlib import contextmanager

import pytest

from cudecomp_b200 import capi as cd

T = {n: getattr(cd, "CUDECOMP_TRANSPOSE_COMM_" + n) for n in
     ("MPI_P2P", "MPI_P2P_PL", "MPI_A2A", "NCCL", "NCCL_PL", "NVSHMEM", "NVSHMEM_PL", "NVSHMEM_SM")}
H = {n: getattr(cd, "CUDECOMP_HALO_COMM_" + n) for n in ("MPI", "MPI_BLOCKING", "NCCL", "NVSHMEM", "NVSHMEM_BLOCKING")}
INV = cd.CUDECOMP_RESULT_INVALID_USAGE


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update(kw)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def options():
    o = cd.cudecompGridDescAutotuneOptions_t()
    assert cd.cudecompGridDescAutotuneOptionsSetDefaults(o) == 0
    return o


def test_all_backend_values_are_candidates_by_default():
    res, t, h, p = cd.autotune_candidates(options(), 4)
    assert res == 0 and sorted(t) == sorted(T.values()) and sorted(h) == sorted(H.values())
    assert p == [(4, 1), (2, 2), (1, 4)]  # locality-first order for row-major ranks (src/autotune.cc:94-106)
    assert cd.autotune_candidates(options(), 4, cd.CUDECOMP_RANK_ORDER_COL_MAJOR)[3] == [(1, 4), (2, 2), (4, 1)]
    assert cd.autotune_candidates(options(), 8)[3] == [(8, 1), (4, 2), (2, 4), (1, 8)]


def test_transpose_inclusion_and_exclusion_lists():
    with env(CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS="NCCL,MPI_A2A,MPI_P2P"):
        assert sorted(cd.autotune_candidates(options())[1]) == sorted([T["MPI_P2P"], T["MPI_A2A"], T["NCCL"]])
    with env(CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS="^MPI_P2P"):
        assert sorted(cd.autotune_candidates(options())[1]) == sorted(v for k, v in T.items() if k != "MPI_P2P")


def test_halo_inclusion_list():
    with env(CUDECOMP_AUTOTUNE_HALO_BACKENDS="NCCL,MPI_BLOCKING"):
        assert sorted(cd.autotune_candidates(options())[2]) == sorted([H["MPI_BLOCKING"], H["NCCL"]])


def test_family_disable_flags():
    o = options()
    o.disable_mpi_backends = True
    _, t, h, _ = cd.autotune_candidates(o)
    assert sorted(t) == sorted(v for k, v in T.items() if not k.startswith("MPI"))
    assert sorted(h) == sorted(v for k, v in H.items() if not k.startswith("MPI"))
    o = options()
    o.disable_nccl_backends = True
    _, t, h, _ = cd.autotune_candidates(o)
    assert sorted(t) == sorted(v for k, v in T.items() if not k.startswith("NCCL"))
    assert sorted(h) == sorted(v for k, v in H.items() if k != "NCCL")
    o = options()
    o.disable_nvshmem_backends = True
    _, t, h, _ = cd.autotune_candidates(o)
    assert sorted(t) == sorted(v for k, v in T.items() if not k.startswith("NVSHMEM"))
    assert sorted(h) == sorted(v for k, v in H.items() if not k.startswith("NVSHMEM"))


def test_malformed_and_empty_backend_sets_are_rejected():
    with env(CUDECOMP_AUTOTUNE_HALO_BACKENDS="MPI,,NCCL"):
        assert cd.autotune_candidates(options())[0] == INV
    with env(CUDECOMP_AUTOTUNE_HALO_BACKENDS="^MPI,MPI_BLOCKING,NCCL,NVSHMEM,NVSHMEM_BLOCKING"):
        assert cd.autotune_candidates(options())[0] == INV
    with env(CUDECOMP_AUTOTUNE_TRANSPOSE_BACKENDS="NCCL,BOGUS"):
        assert cd.autotune_candidates(options())[0] == INV
    o = options()
    o.disable_mpi_backends = o.disable_nccl_backends = o.disable_nvshmem_backends = True
    assert cd.autotune_candidates(o)[0] == INV


def test_process_grid_ranges():
    with env(CUDECOMP_AUTOTUNE_P_ROW_RANGE="2,4"):
        assert sorted(cd.autotune_candidates(options(), 4)[3]) == sorted([(4, 1), (2, 2)])
    with env(CUDECOMP_AUTOTUNE_P_COL_RANGE="2,4"):
        assert sorted(cd.autotune_candidates(options(), 4)[3]) == sorted([(2, 2), (1, 4)])
    with env(CUDECOMP_AUTOTUNE_P_ROW_RANGE="2,2", CUDECOMP_AUTOTUNE_P_COL_RANGE="2,2"):
        assert cd.autotune_candidates(options(), 4)[3] == [(2, 2)]


@pytest.mark.parametrize("name,value", [("CUDECOMP_AUTOTUNE_P_ROW_RANGE", "2,1"), ("CUDECOMP_AUTOTUNE_P_COL_RANGE", "3,3"),
                                        ("CUDECOMP_AUTOTUNE_P_ROW_RANGE", "2"), ("CUDECOMP_AUTOTUNE_P_ROW_RANGE", "a,b"),
                                        ("CUDECOMP_AUTOTUNE_P_COL_RANGE", "1,2,3")])
def test_malformed_and_empty_process_grid_ranges_are_rejected(name, value):
    with env(**{name: value}):
        assert cd.autotune_candidates(options(), 4)[0] == INV


def test_uninitialised_options_are_rejected():
    assert cd.autotune_candidates(cd.cudecompGridDescAutotuneOptions_t(), 4)[0] == INV

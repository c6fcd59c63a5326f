"""include/cudecomp.h of this repo against the reference's public ABI (tests/golden/abi_golden.json, extracted from the
reference's own header by tests/golden/make_golden.py): every enumerator and its value, every struct member in order
with type and array bound, the magics and versions, the 24 exported prototypes and the 5 header-inline wrappers with
their exact return and parameter types, and every C name the reference's Fortran module binds must be there and be
exported by libcudecomp.so. Parameter names are not part of the ABI and are ignored."""
import json
import os
import sys

import pytest

from cudecomp_b200 import capi as cd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from abi_parse import parse_header  # noqa: E402


@pytest.fixture(scope="module")
def abi():
    with open(os.path.join(ROOT, "tests", "golden", "abi_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def ours():
    out = parse_header(open(os.path.join(ROOT, "include", "cudecomp.h")).read())
    out["defines"].update(parse_header(open(os.path.join(ROOT, "include", "cudecomp_version.h")).read())["defines"])
    return json.loads(json.dumps(out))


def test_enumerators_and_values(abi, ours):
    assert len(abi["enums"]) == 6
    assert ours["enums"] == abi["enums"]
    for name, vals in abi["enums"].items():
        for k, v in vals:
            assert getattr(cd, k) == v, k  # the ctypes view agrees too


def test_struct_members(abi, ours):
    assert set(abi["structs"]) == {"cudecompGridDescConfig_t", "cudecompGridDescAutotuneOptions_t", "cudecompPencilInfo_t"}
    assert ours["structs"] == abi["structs"]
    for sname, members in abi["structs"].items():
        fields = [f[0] for f in getattr(cd, sname)._fields_]
        assert fields == [m[1] for m in members], sname


def test_magics_and_versions(abi, ours):
    assert len(abi["defines"]) >= 9
    assert ours["defines"] == abi["defines"]


def test_prototypes(abi, ours):
    assert len(abi["functions"]) == 24 and len(abi["inline_wrappers"]) == 5
    assert ours["functions"] == abi["functions"]
    assert ours["inline_wrappers"] == abi["inline_wrappers"]
    for name in abi["functions"]:
        assert hasattr(cd.lib, name), name


def test_fortran_module_symbols_are_exported(abi):
    assert len(abi["fortran_bindings"]) >= 20 and "cudecompInit_F" in abi["fortran_bindings"]
    missing = [n for n in abi["fortran_bindings"] if not hasattr(cd.lib, n)]
    assert not missing, missing

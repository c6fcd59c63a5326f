#!/usr/bin/env python
"""Extracts the known-answer tables of the reference's API tests into tests/golden/*.json.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources (all in /root/reference/tests/ctest/api_tests.cc):
  * kExpected{Default,ColumnMajor,GdimsDist}PencilInfo  (:92-153)  -- gdims 9x10x11, pdims 2x2, halo (1,2,1),
    padding (1,0,2), [axis][rank] -> shape / lo / hi / order / size
  * expectShiftedRanks(...) calls                         (:1386-1408) -- per-rank neighbour tables
  * dtype sizes (:449-459) and backend names (:467-493)
The parser reads the C++ initialiser lists; nothing is typed by hand.

Also writes abi_golden.json: the reference's public ABI as read from ITS include/cudecomp.h and cudecomp_version.h by
tests/golden/abi_parse.py (enumerators and values, struct members, magics / versions, the 24 prototypes and the 5
header-inline wrappers reduced to return and parameter types) plus the C names its Fortran module binds
(src/cudecomp_m.cuf). tests/test_abi_golden.py holds this repo's header and library against it.
"""
import json
import os
import re
import sys

REF = "/root/reference/tests/ctest/api_tests.cc"
OUT = os.path.dirname(os.path.abspath(__file__))


def ints(s):
    return [int(v) for v in re.findall(r"-?\d+", s)]


def parse_constants(src):
    consts = {}
    for name in ("kGdims", "kGdimsDist", "kPdims", "kHaloExtents", "kPadding"):
        m = re.search(r"constexpr std::array<int32_t, \d>\s+%s\{([^}]*)\}" % name, src)
        consts[name] = ints(m.group(1))
    m = re.search(r"kHaloPeriods\{([^}]*)\}", src)
    consts["kHaloPeriods"] = [v.strip() == "true" for v in m.group(1).split(",")]
    return consts


def parse_pencil_table(src, name, consts):
    m = re.search(r"%s\[3\]\[kApiTestRanks\] = \{(.*?)\n\};" % name, src, re.S)
    body = m.group(1)
    rows = re.findall(r"\{\{([^}]*)\}, \{([^}]*)\}, \{([^}]*)\}, \{([^}]*)\}, kHaloExtents, kPadding, (\d+)\}", body)
    assert len(rows) == 12, (name, len(rows))
    table = []
    for axis in range(3):
        per_rank = []
        for rank in range(4):
            shape, lo, hi, order, size = rows[axis * 4 + rank]
            per_rank.append(dict(shape=ints(shape), lo=ints(lo), hi=ints(hi), order=ints(order),
                                 halo_extents=consts["kHaloExtents"], padding=consts["kPadding"], size=int(size)))
        table.append(per_rank)
    return table


def parse_shifted(src):
    out = {}
    for test, key in (("ReturnsExpectedRanksForRowMajorLayout", "row_major"),
                      ("ReturnsExpectedRanksForColumnMajorLayout", "col_major")):
        m = re.search(r"TEST_F\(ApiGetShiftedRankTest, %s\)(.*?)\n\}" % test, src, re.S)
        calls = re.findall(r"expectShiftedRanks\(active_comm_, handle_, grid_desc, (\d), (\d), (-?\d+), (true|false), "
                           r"\{([^}]*)\}\);", m.group(1))
        out[key] = [dict(axis=int(a), dim=int(d), displacement=int(s), periodic=(p == "true"), expected=ints(e))
                    for a, d, s, p, e in calls]
        assert len(out[key]) == 6
    return out


def parse_strings(src):
    pairs_t = re.findall(r'EXPECT_STREQ\("([^"]*)", cudecompTransposeCommBackendToString\((\w+)\)\)', src)
    pairs_h = re.findall(r'EXPECT_STREQ\("([^"]*)", cudecompHaloCommBackendToString\((\w+)\)\)', src)
    sizes = re.findall(r"cudecompGetDataTypeSize\((CUDECOMP_\w+), &dtype_size\)\);\s*EXPECT_EQ\((\d+), dtype_size\)",
                       src)
    pairs_t = [p for p in pairs_t if p[1].startswith("CUDECOMP_")]
    pairs_h = [p for p in pairs_h if p[1].startswith("CUDECOMP_")]
    return dict(transpose_backend_names=[[n, s] for s, n in pairs_t], halo_backend_names=[[n, s] for s, n in pairs_h],
                dtype_sizes=[[n, int(s)] for n, s in sizes])


def main():
    if not os.path.exists(REF):
        sys.exit("reference not mounted at /root/reference; fixtures are generated in the build container only")
    src = open(REF).read()
    consts = parse_constants(src)
    golden = dict(
        source="reference tests/ctest/api_tests.cc (v0.7.0)",
        gdims=consts["kGdims"], gdims_dist=consts["kGdimsDist"], pdims=consts["kPdims"],
        halo_extents=consts["kHaloExtents"], padding=consts["kPadding"], halo_periods=consts["kHaloPeriods"],
        pencil_info=dict(
            default=parse_pencil_table(src, "kExpectedDefaultPencilInfo", consts),
            column_major=parse_pencil_table(src, "kExpectedColumnMajorPencilInfo", consts),
            gdims_dist=parse_pencil_table(src, "kExpectedGdimsDistPencilInfo", consts)),
        shifted_ranks=parse_shifted(src),
    )
    golden.update(parse_strings(src))
    path = os.path.join(OUT, "api_golden.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote", path)

    sys.path.insert(0, OUT)
    from abi_parse import fortran_bindings, parse_header
    abi = parse_header(open("/root/reference/include/cudecomp.h").read())
    abi["defines"].update(parse_header(open("/root/reference/include/cudecomp_version.h").read())["defines"])
    abi["fortran_bindings"] = fortran_bindings(open("/root/reference/src/cudecomp_m.cuf").read())
    abi["source"] = "reference include/cudecomp.h, include/cudecomp_version.h, src/cudecomp_m.cuf (v0.7.0)"
    path = os.path.join(OUT, "abi_golden.json")
    with open(path, "w") as f:
        json.dump(abi, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()

"""CUDECOMP_ENABLE_PERFORMANCE_REPORT compatibility (reference docs/env_vars.rst:39-95, src/performance.cc): the summary
table printed at cudecompGridDescDestroy and the CSV files keep the reference's columns."""
import glob
import os
import tempfile

import pytest

from tests._launcher import run_ranks

pytestmark = [pytest.mark.gpu]


def test_performance_report_table_and_csv():
    out_dir = tempfile.mkdtemp(prefix="cdb200_perf_")
    base = dict(kind="transpose", gdims=[40, 36, 44], pdims=[2, 1], dtype="double", fills=["random"])
    cases = [dict(base, name="oop", ops=["XY", "YZ", "ZY", "YX"] * 3, out_of_place=True),
             dict(base, name="inplace", ops=["XY", "YX"] * 3),
             dict(kind="halo", name="halo", gdims=[40, 36, 44], pdims=[2, 1], dtype="float", axis=0, halo=[1, 1, 1],
                  periods=[True] * 3, fills=["random", "pattern"])]
    results, logs = run_ranks(2, "gpu", cases, timeout=300, extra_env={
        "CUDECOMP_ENABLE_PERFORMANCE_REPORT": "1", "CUDECOMP_PERFORMANCE_REPORT_WARMUP_SAMPLES": "1",
        "CUDECOMP_PERFORMANCE_REPORT_DETAIL": "2", "CUDECOMP_PERFORMANCE_REPORT_WRITE_DIR": out_dir})
    for r in range(2):
        for c in results[r]:
            assert c["ok"], c
    log = logs[0]
    assert "CUDECOMP: ===== Performance Summary =====" in log
    assert "Transpose Performance Data:" in log and "Halo Performance Data:" in log
    rows = [l for l in log.splitlines() if l.startswith("CUDECOMP: Transpose")]
    assert any(l.split()[1] == "TransposeXY" and l.split()[2] == "D" for l in rows), rows
    # an out-of-place exchange is one kernel: everything is A2A time, nothing local; in-place has a local unpack
    xy = [l.split() for l in rows if l.split()[1] == "TransposeXY" and l.split()[2] == "D"]  # table rows, not sample headers
    assert len(xy) == 2
    for f in xy:
        total, a2a, local = float(f[-4]), float(f[-3]), float(f[-2])
        assert total > 0 and abs(total - (a2a + local)) < 1e-3 + 0.02 * total
    assert "CUDECOMP: HaloX" in log
    tfiles = glob.glob(os.path.join(out_dir, "cudecomp-perf-report-transpose-aggregated-*pdims_2x1-gdims_40x36x44-*.csv"))
    hfiles = glob.glob(os.path.join(out_dir, "cudecomp-perf-report-halo-aggregated-*.csv"))
    assert tfiles and hfiles
    text = open(tfiles[0]).read()
    assert ("operation,dtype,input_halo_extents,output_halo_extents,input_padding,output_padding,inplace,managed,"
            "samples,total_ms,A2A_ms,local_ms,A2A_BW_GBps") in text
    assert "# Process grid: [2, 1]" in text
    assert "operation,dtype,dim,halo_extent,periods,padding,managed,samples,total_ms,SR_ms,local_ms,SR_BW_GBps" in \
        open(hfiles[0]).read()
    # detail level 2: per-sample rows of both ranks, and the per-sample CSV files (reference src/performance.cc:560-770)
    assert "CUDECOMP: Per-Sample Details:" in log
    assert "CUDECOMP: TransposeXY (dtype=D, halo extents=[0,0,0]/[0,0,0], padding=[0,0,0]/[0,0,0], inplace=N, managed=N) samples:" in log
    sample_rows = [l.split() for l in log.splitlines() if l.startswith("CUDECOMP: ") and len(l.split()) == 7 and
                   l.split()[1] in ("0", "1") and l.split()[2].isdigit()]
    assert {r[1] for r in sample_rows} == {"0", "1"}  # both ranks' samples gathered on rank 0
    sfiles = glob.glob(os.path.join(out_dir, "cudecomp-perf-report-transpose-samples-*.csv"))
    assert sfiles and ("operation,dtype,input_halo_extents,output_halo_extents,input_padding,output_padding,inplace,managed,"
                       "rank,sample,total_ms,A2A_ms,local_ms,A2A_BW_GBps") in open(sfiles[0]).read()
    assert glob.glob(os.path.join(out_dir, "cudecomp-perf-report-halo-samples-*.csv"))

"""The host-only helper scripts keep working: scripts/explain_plan.py (plans, tiling and the time model for a
configuration) and scripts/r2_summarize.py (markdown table from bench JSON lines)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_explain_plan_reproduces_the_measured_round_trips():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "explain_plan.py"), "--grid", "1024", "--pdims", "2x4",
                          "--chunks", "8"], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    last = out.stdout.strip().splitlines()[-1]
    # "round trip, model: out of place (direct) 7.81 ms; in place, separate launches 10.48 ms; in place fused (K = 8) 8.9 ms"
    nums = [float(tok) for tok in last.replace(";", " ").split() if tok.replace(".", "", 1).isdigit() and "." in tok]
    direct, staged, chunked = nums
    # measured on 8 GPUs (profiles/r2_n8_results.md): 7.83 ms out of place, 10.5 ms separate launches, 9.0-9.1 ms fused
    assert abs(direct - 7.83) < 0.3 and abs(staged - 10.5) < 0.3 and abs(chunked - 9.05) < 0.4
    assert direct < chunked < staged
    # every step of every chunked operation is a complete all-to-all (no step talks to a single peer of a 4-rank group),
    # and most of the local unpack overlaps later pushes (a hazard test that is merely too cautious would pass every
    # correctness test and silently lose this)
    import re
    for line in out.stdout.splitlines():
        if "peers per step" in line:
            assert "peers per step [2]" in line or "peers per step [4]" in line, line
            overlapped = float(re.search(r"unpacked beside later pushes (\d+) %", line).group(1))
            assert overlapped >= 75, line


def test_r2_summarize_reads_bench_lines(tmp_path):
    for name, src in (("r2_n8_default.json", "r1_n8_bench.json"), ("r2_n8_inplace.json", "r1_n8_bench_inplace.json")):
        with open(os.path.join(ROOT, "profiles", src)) as f:
            line = json.loads(f.readline())
        (tmp_path / name).write_text(json.dumps(line) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "r2_summarize.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "### 8 GPUs" in out.stdout and "| default |" in out.stdout and "| inplace |" in out.stdout


def test_runbooks_only_use_flags_bench_py_knows():
    """The scripts/r2_*.sh runbooks spend GPU minutes: every `bench <label> <args>` line must parse with bench.py's own
    argument parser (a typo would only show up on the GPU box)."""
    import glob
    import importlib.util
    import re
    import shlex
    spec = importlib.util.spec_from_file_location("bench_main", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    lines = []
    for path in sorted(glob.glob(os.path.join(ROOT, "scripts", "r2_*.sh"))):
        for line in open(path).read().splitlines():
            m = re.match(r"\s*(?:[A-Z_0-9]+=\S+ )?(?:for \w+ in [^;]+; do )?bench (\S+)(.*?)(?:; done)?$", line)
            if m and "$label" not in line and "()" not in line:
                lines.append(re.sub(r"\$[A-Z]\w*", "", m.group(2)))  # $Q / $P: flag bundles defined in the script
    assert len(lines) > 60
    old_argv = sys.argv
    try:
        for args in lines:
            args = re.sub(r"\$\w+", "8", args)  # loop variables ($k, $t)
            sys.argv = ["bench.py"] + shlex.split(args)
            parsed = bench.parse_args()
            assert parsed.n in (512, 640, 1024)
    finally:
        sys.argv = old_argv


def test_reference_sweep_on_the_cpu_runs_the_committed_case_lists():
    """scripts/ref_sweep_cpu.py: the reference's own command lines (tests/golden/ref_cases, what the GPU suite feeds to the
    reference's executables) parsed the way those executables parse them, planned by the library, walked by the launch
    emulator and compared with the oracle -- here the first two lines of every configuration; the whole 8 268-line sweep
    is recorded in profiles/r2_cpu_reference_sweep.txt."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_sweep_cpu", os.path.join(ROOT, "scripts", "ref_sweep_cpu.py"))
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    d = sweep.parse_transpose("--pr 1 --pc 4 --backend 1 --gx 128 --gy 124 --gz 132 --gd 0 0 2 --hex 0 0 0 --hey 1 1 1 "
                              "--hez 0 0 0 --pdx 0 0 0 --pdy 1 0 1 --pdz 0 0 0 --mem_order 0 1 2 1 2 0 2 0 1 -m -o")
    assert d["gdims"] == [128, 124, 132] and d["pdims"] == [1, 4] and d["gdims_dist"] == [128, 124, 130]
    assert d["halos"]["1"] == [1, 1, 1] and d["pads"]["1"] == [1, 0, 1] and d["mem_order"][1] == [1, 2, 0] and d["out_of_place"]
    d, ax, halo, periods, padding = sweep.parse_halo("--pr 2 --pc 2 --backend 1 --gx 64 --gy 66 --gz 62 --gd 0 0 0 --hex 1 --hey 0 "
                                                     "--hez 2 --hpx 1 --hpy 0 --hpz 1 --pdx 1 --pdy 2 --pdz 3 --ax 2 --ac 1 "
                                                     "--mem_order 2 0 1")
    # halo_test's option table routes --pdz to padding[1] (tests/cc/halo_test.cc:274-276)
    assert (ax, halo, periods, padding) == (2, [1, 0, 2], [True, False, True], [1, 3, 0])
    assert d["axis_contiguous"] == [True] * 3 and d["mem_order"] == [[2, 0, 1]] * 3
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ref_sweep_cpu.py"),
                          os.path.join(ROOT, "tests", "golden", "ref_cases"), "--jobs", "4", "--limit", "2"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "TOTAL 28 lines, 28 passed, 0 failed" in out.stdout

#!/usr/bin/env python
"""Turns the JSON lines the scripts/r2_*.sh runbooks leave in gpurun_out/ into one markdown table per GPU count
(label, ms per round trip, per-operation ms, roofline fraction, path), next to the default schedule of the same run, so
that the winners can be made defaults and the table can be committed under profiles/.
    python scripts/r2_summarize.py [gpurun_out] > profiles/r2_schedules.md"""
import glob
import json
import os
import re
import sys


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    by_n = {}
    for path in sorted(glob.glob(os.path.join(d, "r2_n*_*.json"))):
        m = re.match(r"r2_n(\d+)_(.+)\.json", os.path.basename(path))
        if not m:
            continue
        try:
            with open(path) as f:
                line = json.loads(f.readline())
        except (ValueError, OSError):
            continue
        if "ms_per_step" not in line:
            continue
        by_n.setdefault(int(m.group(1)), []).append((m.group(2), line))
    for n in sorted(by_n):
        rows = by_n[n]
        base = {lab: l for lab, l in rows}
        print("### %d GPU%s\n" % (n, "" if n == 1 else "s"))
        print("| run | workload | ms / round trip | vs default | XY / YZ / ZY / YX ms | bound | achieved GB/s | frac | path |")
        print("|---|---|---|---|---|---|---|---|---|")
        for lab, l in rows:
            r = l.get("roofline", {})
            ops = r.get("per_op_ms", {})
            wl = l.get("config", {}).get("workload", "")
            ref = base.get("c64_512" if lab.startswith("c64_512") else ("inplace" if "inplace" in lab else "default"))
            rel = "%.3f" % (l["ms_per_step"] / ref["ms_per_step"]) if ref and ref is not l else "1"
            print("| %s | %s | %.3f | %s | %s | %s | %.1f | %.3f | %s |" % (
                lab, wl.split(" X->")[0] + (", in place" if "in-place" in wl else ""), l["ms_per_step"], rel,
                " / ".join("%.3f" % ops.get(k, float("nan")) for k in ("XY", "YZ", "ZY", "YX")), r.get("bound", "?"),
                r.get("achieved", float("nan")), r.get("frac", float("nan")), l.get("path", "?")))
            e2e = l.get("e2e")
            if e2e:
                hl = e2e.get("host_link", {})
                print("| %s (e2e) | host buffers in and out | %.1f | | | host link | %.1f | | %s; link %s |" % (
                    lab, e2e.get("ms_per_step", float("nan")), e2e.get("value", float("nan")), e2e.get("host_binding", ""),
                    {k: round(v, 1) for k, v in hl.items() if isinstance(v, float)}))
        print()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py full  gpurun_out/prof.ncu-rep   profiles/r1_rowcopy_full.txt
  python profiles/summarize_ncu.py list  gpurun_out/launches.csv   profiles/r1_launches.txt
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "lts__t_sectors_srcunit_tex_aperture_peer.sum", "lts__t_sectors_aperture_peer.sum",
    "nvltx__bytes.sum", "nvlrx__bytes.sum", "pcie__write_bytes.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, summary of %s\n" % rep.split("/")[-1])
        for k, r in enumerate(rows[2:]):
            f.write("\n[launch %d] %s  grid=%s block=%s\n" % (k, r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")],
                                                               r[hdr.index("Block Size")]))
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("  %-70s %s %s\n" % (m, r[i], units[i]))
    print("wrote", out)


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    hdr = ["ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size", "Grid Size",
           "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"]
    total = {}
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\n")
        f.write("# id kernel grid block duration\n")
        for r in rows:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0][-60:]
            val = float(d["Metric Value"].replace(",", ""))
            unit = d["Metric Unit"]
            ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
            total[name] = total.get(name, 0) + ns
            f.write("%s %s %s %s %.1f us\n" % (d["ID"], name, d["Grid Size"], d["Block Size"], ns / 1e3))
        f.write("\n# share of device time by kernel\n")
        s = sum(total.values()) or 1
        for k, v in sorted(total.items(), key=lambda kv: -kv[1]):
            f.write("%6.2f%%  %10.1f us  %s\n" % (100 * v / s, v / 1e3, k))
    print("wrote", out)


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])

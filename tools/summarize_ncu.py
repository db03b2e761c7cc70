"""Summaries of ncu output for profiles/ (the raw reports stay in gpurun_out/).
  summarize_ncu.py launches <launch-list.csv> <out.csv> [title]      per-kernel totals/shares of a --metrics gpu__time_duration.sum pass
  summarize_ncu.py full <report.ncu-rep> <out.csv>                   selected metrics per captured launch of a --set full report
"""
import collections
import csv
import re
import subprocess
import sys

UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace(", ", ";")


def launches(src, dst, title=""):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ms = float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
        k = short(row["Kernel Name"])
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    T = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none; times are cold-cache and serialised: compare SHARES\n")
        f.write(f"# total {T:.3f} ms over {n} launches\nkernel,launches,total_ms,share,avg_us\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1]:.4f},{v[1] / T:.3f},{1e3 * v[1] / v[0]:.1f}\n")


def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in FULL if c in idx]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, one row per captured launch (units in the second row)\n")
        f.write("kernel," + ",".join(cols) + "\n")
        f.write("," + ",".join(units[idx[c]] for c in cols) + "\n")
        for d in data:
            f.write(short(d[idx["Kernel Name"]]).replace(",", ";") + "," + ",".join(d[idx[c]].replace(",", "") for c in cols) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])

"""Summarise gpurun_out/ ncu artefacts into profiles/<tag>.md (tracked).  Usage:
   python tools/ncu_summary.py <tag> <launches.csv> [<report.ncu-rep>] [--note "..."]"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
           "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def main():
    tag, lpath = sys.argv[1], sys.argv[2]
    rep = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    out = [f"# ncu summary `{tag}`", "", note, "",
           "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    agg = launches(lpath)
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:70]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
    if rep:
        raw = (open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        out += ["", f"## `ncu --set full --clock-control none` ({rep.split('/')[-1]}), per launch", "",
                "| kernel | " + " | ".join(m.split(".")[0].replace("__", " ") for m in METRICS) + " |", "|---|" + "---:|" * len(METRICS)]
        idx = [hdr.index(m) if m in hdr else -1 for m in METRICS]
        ki = hdr.index("Kernel Name")
        seen = collections.Counter()
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "")
            seen[name] += 1
            if seen[name] > 2:
                continue
            out.append(f"| `{name[:50]}` | " + " | ".join((r[i] + " " + units[i]) if i >= 0 else "n/a" for i in idx) + " |")
    open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()

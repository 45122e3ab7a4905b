"""Summarise the CSV exports of tools/ncu_round2.sh (gpurun_out/<tag>_*) into profiles/<out>.md and profiles/traffic.json.
usage: python tools/ncu_report.py <tag> <out-name>"""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
KERNELS = [("k_lin_pipe", "k_lin_pipe"), ("k_pt_pipeint0", "k_pt_pipe<0>"), ("k_pt_pipeint1", "k_pt_pipe<1>"), ("k_linearize_cm", "k_linearize_cm"),
           ("k_schur_cm", "k_schur_cm"), ("k_spmv_cm", "k_spmv_cm"), ("k_schur_pairs", "k_schur_pairs"), ("k_cg_bsr", "k_cg_bsr<1>"), ("k_cam_pipe", "k_cam_pipe<2>")]
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)"), ("launch__occupancy_limit_registers", "CTAs/SM (regs)"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts")]


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
        name = r[ki].split("(glba::")[0].split("(const")[0].split("(int")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def raw(path):
    rows = list(csv.reader(open(path)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def stalls(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) > 40]
    cols = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {n: sum(int(r[cols[n]]) for r in data) for n in names}
    tot = max(1, sum(agg.values()))
    sass = collections.Counter()
    for r in data:
        op = r[1].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        for key in ("UBLKCP", "SYNCS", "LDG.E.ENL2.256", "STG.E.ENL2.256", "BAR.SYNC", "ACQBULK", "PREEXIT", "MEMBAR", "CCTL", "RED", "SHFL", "LDS"):
            if op.startswith(key):
                sass[key] += 1
    return sorted(((n[6:], 100.0 * v / tot) for n, v in agg.items()), key=lambda x: -x[1])[:6], sass


md = [f"# ncu summary `{out}` (round 2, final kernels)", "",
      "`bash tools/ncu_round2.sh` on one B200: C4 (1 800 cameras on a 60x30 street grid, 1 M points, 5.0 M observations), "
      "`python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lm-iters 2`; both passes with `--clock-control none`, exported on the box with "
      "`ncu -i ... --page raw/source --csv`.  Per-launch ncu times are cold-cache and serialised: compare SHARES; the CUDA-event times of the "
      "same kernels are in the bench line (`kernels`).", "",
      "## Launch list (`--metrics gpu__time_duration.sum`)", "", "| kernel | launches | mean us | share |", "|---|---:|---:|---:|"]
agg = launches(os.path.join(G, f"{tag}_launches.csv"))
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    md.append(f"| `{n[:70]}` | {a[0]} | {a[1] / a[0] / 1000:.2f} | {100 * a[1] / tot:.1f} % |")
md += ["", "## `--set full` of one launch of every hot kernel", ""]
traffic = {}
for key, name in KERNELS:
    try:
        R = raw(os.path.join(G, f"{tag}_{key}_raw.csv"))
        st, sass = stalls(os.path.join(G, f"{tag}_{key}_source.csv"))
    except Exception as e:
        md.append(f"### `{name}`: no capture ({e})")
        continue
    md += [f"### `{name}`", "", "| metric | value |", "|---|---:|"]
    for m, label in WANT:
        if m in R:
            md.append(f"| {label} | {R[m][1]} {R[m][0]} |")
    rd, wr = float(R["dram__bytes_read.sum"][1].replace(",", "")), float(R["dram__bytes_write.sum"][1].replace(",", ""))
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    traffic[name] = rd * unit.get(R["dram__bytes_read.sum"][0], 1e6) + wr * unit.get(R["dram__bytes_write.sum"][0], 1e6)
    md.append(f"| warp stall samples | {', '.join('%s %.0f %%' % s for s in st)} |")
    md.append(f"| SASS of this kernel (static count) | {', '.join('%s x%d' % kv for kv in sorted(sass.items()))} |")
    md.append("")
open(os.path.join(ROOT, "profiles", out + ".md"), "w").write("\n".join(md) + "\n")
json.dump({"source": f"profiles/{out}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, C4)", **traffic},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(md[:60]))

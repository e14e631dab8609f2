#!/usr/bin/env python
"""Summaries of ncu output for profiles/.

  python tools/ncu_summary.py launches <ncu --csv log> "<title>" "<command>"   -> markdown launch list
  python tools/ncu_summary.py full <file.ncu-rep> "<title>" "<command>"        -> markdown table of --set full metrics
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path, title, cmd):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = [f"# {title}", "", f"Command: `{cmd}`",
           "(cold-cache, serialised launch times: compare SHARES, not absolutes)", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    out += ["", f"Total {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches."]
    return "\n".join(out)


METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU instructions"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (spill fills)"),
    ("smsp__inst_executed_op_local_st.sum", "local stores (spills)"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_hit.sum", "local load sectors hit in L1"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local load sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed-latency dependency) / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
]


def full(path, title, cmd):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [re.sub(r"\(.*", "", d[hdr.index("Kernel Name")]).replace("void ", "") for d in data]
    out = [f"# {title}", "", f"Command: `{cmd}`", "", "| metric | " + " | ".join(f"`{n}`" for n in names) + " |",
           "|---|" + "---:|" * len(names)]
    for key, label in METRICS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for d in data:
            try:
                vals.append(f"{float(d[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(d[i])
        out.append(f"| {label} ({units[i]}) | " + " | ".join(vals) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    mode, path, title, cmd = sys.argv[1:5]
    print(launches(path, title, cmd) if mode == "launches" else full(path, title, cmd))

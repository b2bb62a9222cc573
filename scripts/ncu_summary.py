#!/usr/bin/env python
"""Turn what a GPU visit left in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/ncu_summary.py <tag> [--rep gpurun_out/prof_x.ncu-rep ...] [--launches gpurun_out/launches.csv]

Writes profiles/<tag>_ncu.md (per-kernel table from `ncu --page raw`), profiles/<tag>_launches.md
(the launch list with each kernel's share of the step) and merges the per-launch DRAM traffic of every
profiled kernel into profiles/traffic.json (bench.py reports it as roofline.traffic).
Runs here, without a GPU: it only reads reports.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]


def short_name(full):
    m = re.search(r"(?:vael::)?(\w+)\s*(<[^(]*>)?\s*\(", full)
    return (m.group(1) + (m.group(2) or "")) if m else full[:60]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1)


def to_us(value, unit):
    v = float(value.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def read_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", nargs="*", default=[])
    ap.add_argument("--launches", default=None)
    ap.add_argument("--note", default="")
    ap.add_argument("--rows", type=int, default=None, help="batch rows per launch of the profiled command")
    a = ap.parse_args()
    os.makedirs(PROFILES, exist_ok=True)
    tpath = os.path.join(PROFILES, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}

    if a.rep:
        lines = [f"# ncu --set full summary — {a.tag}", "", a.note, "",
                 "Per-launch values from `ncu -i <rep> --page raw --csv` (captured with `--clock-control none`; "
                 "durations under ncu are cold-cache and serialised — bench.py's CUDA-event numbers are the measured ones).", ""]
        for rep in a.rep:
            lines += [f"## {os.path.basename(rep)}", ""]
            for d in read_rep(rep):
                name = short_name(d["Kernel Name"][0])
                lines += [f"### {name}", "", "| metric | value |", "|---|---|"]
                for key, label in METRICS:
                    if key in d and d[key][0] != "":
                        lines.append(f"| {label} (`{key}`) | {d[key][0]} {d[key][1]} |")
                rd = to_bytes(*d["dram__bytes_read.sum"])
                wr = to_bytes(*d["dram__bytes_write.sum"])
                us = to_us(*d["gpu__time_duration.sum"])
                lines += [f"| DRAM traffic (read+write) | {(rd + wr) / 1e9:.4f} GB |",
                          f"| DRAM GB/s under ncu | {(rd + wr) / us / 1e3:.0f} |", ""]
                base = name.split("<")[0]
                traffic[base] = {"bytes": rd + wr, "read": rd, "write": wr, "kernel": name, "rows": a.rows,
                                 "source": f"profiles/{a.tag}_ncu.md"}
        open(os.path.join(PROFILES, f"{a.tag}_ncu.md"), "w").write("\n".join(lines) + "\n")
        json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)

    if a.launches:
        txt = open(a.launches).read()
        rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
        ours = [r for r in rows if "vael::" in r["Kernel Name"] or "ad2_" in r["Kernel Name"] or "generic_" in r["Kernel Name"]]
        tot_ours = sum(float(r["Metric Value"]) for r in ours) or 1.0
        per = {}
        for r in ours:
            n = short_name(r["Kernel Name"])
            per.setdefault(n, []).append(float(r["Metric Value"]))
        lines = [f"# launch list — {a.tag}", "", a.note, "",
                 "`ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command; times are "
                 "cold-cache and serialised, so the SHARE per kernel is what is compared with bench.py's live events.", "",
                 "## vael_b200 kernels in the captured steps", "", "| kernel | launches | mean us | share of vael_b200 time |",
                 "|---|---|---|---|"]
        for n, v in per.items():
            lines.append(f"| `{n}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot_ours:.3f} |")
        lines += ["", "## every launch, in order", "", "| id | kernel | block | grid | ns |", "|---|---|---|---|---|"]
        for r in rows:
            lines.append(f"| {r['ID']} | `{short_name(r['Kernel Name'])[:70]}` | {r['Block Size']} | {r['Grid Size']} | {r['Metric Value']} |")
        open(os.path.join(PROFILES, f"{a.tag}_launches.md"), "w").write("\n".join(lines) + "\n")
    print("wrote", [f for f in os.listdir(PROFILES) if f.startswith(a.tag)], "traffic keys:", sorted(traffic))


if __name__ == "__main__":
    sys.exit(main())

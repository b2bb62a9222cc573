"""profiles/<tag>_launches.md from the ncu launch list of ONE train step (scripts/step_launches.py under
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`): per-kernel totals, vael_b200's share.
    python scripts/step_launch_summary.py <tag> <launches.csv> "<command line>" """
import csv, io, os, re, sys
from collections import OrderedDict
tag, path, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
txt = open(path).read()
txt = txt[txt.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(txt)))
agg = OrderedDict()
total = 0.0
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; total += us
ours = sum(v[1] for k, v in agg.items() if "vael::" in k or k.startswith("jit_"))
n = sum(v[0] for v in agg.values())
lines = [f"# launch list of ONE VAEL train step — {tag}", "", f"Command: `{cmd}`", "",
         f"{n} launches, {total / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu).",
         f"**vael_b200 kernels: {ours:.1f} us = {100 * ours / total:.2f} % of the step.**", "", "## by kernel", "",
         "| kernel | launches | total us | share of step |", "|---|---|---|---|"]
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    nm = re.sub(r"\s+", " ", k)[:110]
    nm = f"**`{nm}`**" if ("vael::" in k or k.startswith("jit_")) else f"`{nm}`"
    lines.append(f"| {nm} | {c} | {us:.1f} | {100 * us / total:.2f} % |")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_launches.md")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, "ours", ours, "total", total)

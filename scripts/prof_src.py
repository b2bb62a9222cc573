"""Per-CUDA-line instruction counts of one kernel launch in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, launch, units = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", launch,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
name = next((r[1] for r in rows if r and r[0] == "Function Name"), "?")
ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows:
    if r and r[0].isdigit() and len(r) > ii:
        try:
            data.append((int(r[ii]), int(r[si] or 0), int(r[0]), r[1][:110]))
        except ValueError:
            pass
tot = sum(d[0] for d in data)
print(name, "total warp instructions", tot, "per unit", tot / units)
for d in sorted(data, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print(f"{d[0] / units:9.0f} {100 * d[0] / tot:5.1f}% samp {d[1]:>6} L{d[2]:<4} {d[3]}")

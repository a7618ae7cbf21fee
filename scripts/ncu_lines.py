"""Per-source-line summary of an ncu report (source page, cuda,sass view): samples and instructions per line.

usage: python scripts/ncu_lines.py report.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
lines = []
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "", "Function Name"):
        continue
    try:
        ln = int(r[0]); samp = int(r[4]); inst = int(r[7])
    except ValueError:
        continue
    lines.append((samp, inst, fname, ln, r[1].strip()[:110]))
    tot_s += samp; tot_i += inst
lines.sort(reverse=True)
print(f"total samples {tot_s}, warp instructions {tot_i}")
for samp, inst, f, ln, src in lines[:top]:
    print(f"{100*samp/max(tot_s,1):5.1f}% samp {100*inst/max(tot_i,1):5.1f}% inst  {f}:{ln}  {src}")

# optional: bucket by line ranges given as file:lo-hi:name ...
if len(sys.argv) > 3:
    print()
    for spec in sys.argv[3:]:
        f, rng, name = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(x[0] for x in lines if x[2] == f and lo <= x[3] <= hi)
        i = sum(x[1] for x in lines if x[2] == f and lo <= x[3] <= hi)
        print(f"{name:28s} {100*s/max(tot_s,1):5.1f}% samp {100*i/max(tot_i,1):5.1f}% inst")

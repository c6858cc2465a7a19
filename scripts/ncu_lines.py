"""Per-source-line share of executed warp instructions (and lane efficiency) of the first kernel in an .ncu-rep
captured with --import-source on.  Usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[2]
iex, ith = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
lines = []
for r in rows[3:]:
    if len(r) > ith and r[0] == "File Path":
        if lines:
            break
    if len(r) > ith and r[0].isdigit() and r[iex].isdigit() and r[ith].isdigit():
        lines.append(r)
tot = sum(int(r[iex]) for r in lines)
print(f"total warp instructions {tot}")
for r in sorted(lines, key=lambda r: -int(r[iex]))[:top]:
    ex, th = int(r[iex]), int(r[ith])
    print(f"{r[0]:>5s} {100 * ex / tot:5.1f}%  lanes {th / max(ex, 1):5.1f}  {r[1].strip()[:110]}")

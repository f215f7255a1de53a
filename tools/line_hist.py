#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report (cuda,sass view).
usage: tools/line_hist.py <report.ncu-rep> <kernel regex> [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows, cur_file, total = [], "", 0
rd = csv.reader(io.StringIO(txt))
hdr = None
for r in rd:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or r[0] == "": continue
    try:
        n = int(r[ie]); s = int(r[isamp])
    except ValueError:
        continue
    rows.append((n, s, cur_file, r[0], r[1].strip()[:110]))
    total += n
stot = sum(r[1] for r in rows)
print(f"total warp instructions {total}, samples {stot}")
for n, s, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{n / total:6.1%} {s / max(stot,1):6.1%}  {f}:{ln}  {src}")

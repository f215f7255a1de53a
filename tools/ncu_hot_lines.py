#!/usr/bin/env python
"""Per-source-line totals from `ncu -i rep --page source --csv --print-source sass,cuda --kernel-name regex:K`.

  tools/ncu_hot_lines.py <source_page.csv> [top_n]
Prints, per (file, line): stall samples, share, instructions executed, and the dominant stall reasons."""
import csv
import sys
from collections import defaultdict


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    fname, hdr = None, None
    lines = {}
    tot_s = tot_i = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            s, ins = int(r[i_s]), int(r[i_i])
        except ValueError:
            continue
        st = {h: int(r[k]) for k, h in stall_cols if r[k].isdigit() and int(r[k])}
        lines[(fname, int(r[0]))] = (s, ins, r[1].strip()[:90], st)
        tot_s += s
        tot_i += ins
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    byfile = defaultdict(lambda: [0, 0])
    for (f, _), (s, ins, _, _) in lines.items():
        byfile[f][0] += s
        byfile[f][1] += ins
    for f, (s, ins) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:24s} samples {s / tot_s:6.1%}  inst {ins / tot_i:6.1%}")
    print()
    for (f, ln), (s, ins, src, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        top3 = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{s / tot_s:6.1%} inst {ins / tot_i:6.1%}  {f}:{ln:<5d} {src}\n        [{top3}]")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Opcode histogram of the executed instructions of one kernel from `ncu --page source --csv` output.
usage: tools/sass_hist.py <report.ncu-rep> <kernel regex> [top]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ops, stalls = collections.Counter(), collections.Counter()
tot = 0; live = 0; samples = 0
for r in rows:
    try: n = int(r["Instructions Executed"])
    except (ValueError, KeyError, TypeError): continue
    src = r["Source"].strip()
    if src.startswith("@"): src = src.split(None, 1)[1]
    op = src.split()[0].rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "LDG", "STS", "ATOMS", "SHF", "LOP3", "ISETP", "SHFL")) else op.split(".")[0]
    ops[base] += n; tot += n; live += n > 0
    s = int(r.get("# Samples") or 0); samples += s; stalls[base] += s
print(f"static instructions {len(rows)}, executed at least once {live}, warp instructions executed {tot}")
for op, n in ops.most_common(top):
    print(f"{op:16s} {n:12d} {n / tot:6.1%}   stall samples {stalls[op] / max(samples, 1):6.1%}")

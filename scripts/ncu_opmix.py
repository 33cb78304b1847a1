#!/usr/bin/env python
"""Instruction mix of one kernel from an .ncu-rep source page: executed warp instructions and stall samples per opcode.
usage: python scripts/ncu_opmix.py rep.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several kernels may match: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr = b["rows"][0]
    iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, sm = collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in b["rows"][1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        op = r[iS].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "BAR")) and "." in op else "")
        ex[op] += int(r[iE]); sm[op] += int(r[iSamp] or 0)
        tot += int(r[iE]); tots += int(r[iSamp] or 0)
    print("==", b["name"][:90], "total inst", tot, "samples", tots)
    for op, c in ex.most_common(top):
        print(f"  {op:14s} {c:12d} {100*c/tot:5.1f}%   samples {100*sm[op]/max(tots,1):5.1f}%")

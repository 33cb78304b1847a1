#!/usr/bin/env python
"""Development: per-source-line view of an ncu report (needs -lineinfo): samples, executed instructions and top stalls per CUDA source line.
usage: python scripts/ncu_src.py rep.ncu-rep kernel_regex [file_filter]"""
import csv, io, subprocess, sys, re
from collections import defaultdict
rep, kre = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
print([x for x in h][:12])

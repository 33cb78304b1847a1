#!/usr/bin/env python
"""Development: join an ncu SASS source page with nvdisasm line info -> per-CUDA-source-line samples / executed instructions / top stalls.
usage: python scripts/ncu_lines.py rep.ncu-rep kernel_regex mangled_substring [lib.so]"""
import csv, io, subprocess, sys, re, os, tempfile, glob
from collections import defaultdict, Counter
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else "vision-spectra_b200/lib/libvspectra.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(tmp + "/*.cubin")[0]], capture_output=True, text=True).stdout
# locate the function, walk lines
line_of = {}
cur = None; infn = False
for ln in dis.split("\n"):
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
    if m and cur: line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
sel = sys.argv[5] if len(sys.argv) > 5 else None  # substring of the demangled kernel name (several kernels match the regex)
if sel:
    k0 = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and sel in r[1]][0]
    rows = rows[k0:]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, iex, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
agg = defaultdict(lambda: [0, 0, Counter()])
base = None; tot_s = tot_e = 0
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name": break  # the next kernel of a multi-kernel match
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    key = line_of.get(a - base, ("?", 0))
    e, s = int(r[iex]), int(r[isamp])
    agg[key][0] += s; agg[key][1] += e; tot_s += s; tot_e += e
    for j in stall_cols:
        if r[j] not in ("", "0"): agg[key][2][h[j][6:]] += int(r[j])
print(f"total samples {tot_s}  warp-instructions {tot_e}")
src = {}
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    f, l = key
    if f not in src:
        try: src[f] = open(os.path.join("vision-spectra_b200/csrc", f)).read().split("\n")
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:70] if 0 < l <= len(src[f]) else ""
    print(f"{f}:{l:4d} samp {s/tot_s*100:5.1f}% exec {e/tot_e*100:5.1f}%  {dict(st.most_common(3))}  | {text}")
# phase ranges (optional): VSP_RANGES="name:lo-hi,lo-hi;name2:..."
rng = os.environ.get("VSP_RANGES")
if rng:
    for spec in rng.split(";"):
        name, rs = spec.split(":")
        s = e = 0; st = Counter()
        for r_ in rs.split(","):
            lo, hi = map(int, r_.split("-"))
            for (f, l), (ss, ee, stt) in agg.items():
                if f.startswith("sbr_band") and lo <= l <= hi:
                    s += ss; e += ee; st.update(stt)
        print(f"{name:8s} samples {s/tot_s*100:5.1f}%  exec {e/tot_e*100:5.1f}%  {dict(st.most_common(5))}")

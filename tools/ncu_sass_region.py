#!/usr/bin/env python3
"""SASS rows of an ncu report whose source line falls in [lo, hi] of r2p2_kernels.cuh, with samples and top stall reason."""
import csv, io, os, re, subprocess, sys, tempfile
rep, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "r2p2_b200/csrc/libr2p2_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and key in l and l.rstrip().endswith(":"))
lines = []; cur = ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    s = l.strip()
    if re.match(r"^(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", s) and not s.startswith(".") and s.endswith(";"): lines.append((cur, s))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:episode_kernel"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr = rows[1]; data = rows[2:]; ci = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
acc = 0
for (loc, s), r in zip(lines, data):
    if loc[0] == "r2p2_kernels.cuh" and lo <= loc[1] <= hi:
        n = int(r[ci["# Samples"]] or 0); acc += n
        top = max(st, key=lambda h: int(r[ci[h]] or 0))
        print("%4d %5d %5.2f%% %-16s %s" % (loc[1], n, 100.0 * n / tot, top if n else "", s[:90]))
print("region samples %d of %d = %.1f%%" % (acc, tot, 100.0 * acc / tot))

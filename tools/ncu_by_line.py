#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <kernel mangled-name substring> [top]
Needs nvdisasm/cuobjdump (line info comes from the cubin inside r2p2_b200/csrc/libr2p2_b200.so, built with -lineinfo).
The ncu CLI's `--print-source cuda` view carries no metrics, so SASS rows are matched, in order, with
`nvdisasm -g` output of the same function.
"""
import csv, io, os, re, subprocess, sys, tempfile
rep, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "r2p2_b200/csrc/libr2p2_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and key in l and l.rstrip().endswith(":"))
lines = []   # (file, line) per instruction in order
cur = ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    s = l.strip()
    if re.match(r"^(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", s) and not s.startswith(".") and s.endswith(";"):
        lines.append((cur, s))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:episode_kernel"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
print("sass rows", len(data), "nvdisasm instr", len(lines))
n = min(len(data), len(lines))
agg = {}
tot_i = tot_s = 0
for (loc, s), r in zip(lines[:n], data[:n]):
    ie = int(r[ci["Instructions Executed"]] or 0); sm = int(r[ci["# Samples"]] or 0)
    a = agg.setdefault(loc, [0, 0]); a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
src = {}
for (f, ln) in agg:
    p = os.path.join(ROOT, "r2p2_b200/csrc", f)
    if f not in src and os.path.exists(p):
        src[f] = open(p).read().splitlines()
print("total warp-inst %d samples %d" % (tot_i, tot_s))
for (f, ln), (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:110] if f in src and ln - 1 < len(src[f]) else ""
    print("%-16s %4d  inst %5.2f%%  samples %5.2f%% | %s" % (f, ln, 100.0 * ie / max(tot_i, 1), 100.0 * sm / max(tot_s, 1), text))

# ---- coarse buckets -------------------------------------------------------------------------------------
def bucket(f, ln):
    if f == "r2p2_math.h":
        if ln < 131: return "math: bits/norm2/sqrt"
        if ln < 309: return "math: sincos"
        if ln < 362: return "math: atan2"
        return "math: exp/tanh/logistic"
    if f in ("r2p2_kernels.cuh", "r2p2_kernel_ws.cuh"):
        return "kernel"
    if "intrinsics" in f: return "intrinsics (shfl/funnelshift/...)"
    return "other:" + f
if os.environ.get("BUCKETS"):
    # kernel file: split by line ranges given as "name:lo-hi,name:lo-hi"
    ranges = []
    for item in os.environ["BUCKETS"].split(","):
        nm, fn, r = item.split(":"); lo, hi = r.split("-"); ranges.append((nm, {"k": "r2p2_kernels.cuh", "w": "r2p2_kernel_ws.cuh"}[fn], int(lo), int(hi)))
    b = {}
    for (f, ln), (ie, sm) in agg.items():
        name = bucket(f, ln)
        if name == "kernel":
            name = next((nm for nm, fn, lo, hi in ranges if fn == f and lo <= ln <= hi), "kernel: other " + f)
        a = b.setdefault(name, [0, 0]); a[0] += ie; a[1] += sm
    print("---- buckets")
    for name, (ie, sm) in sorted(b.items(), key=lambda kv: -kv[1][1]):
        print("%-28s inst %5.2f%%  samples %5.2f%%" % (name, 100.0 * ie / tot_i, 100.0 * sm / tot_s))

#!/usr/bin/env python3
"""Aggregate an exported ncu SASS source page (ncu -i rep --page source --csv | gzip) by CUDA source line.

usage: ncu_by_line_csv.py <source.csv.gz> <mangled-name substring of the kernel instance> [top] [--sass]
The line info comes from `nvdisasm -g` of the cubin inside r2p2_b200/csrc/libr2p2_b200.so (same build as the capture).
BUCKETS="name:k|w:lo-hi,..." adds coarse buckets by line range of r2p2_kernels.cuh (k) / r2p2_kernel_ws.cuh (w).
"""
import csv, gzip, io, os, re, subprocess, sys, tempfile, collections
src_csv, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 60
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("SASS_TXT"):        # `nvdisasm -g -c` output of the build that was profiled, kept from before later rebuilds
    sass = open(os.environ["SASS_TXT"]).read().splitlines()
else:
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "r2p2_b200/csrc/libr2p2_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and key in l and l.rstrip().endswith(":"))
lines = []
cur = ("?", 0)
inl = []
for l in sass[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    s = l.strip()
    if re.match(r"^(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", s) and not s.startswith(".") and s.endswith(";"):
        lines.append((cur, s))
rows = list(csv.reader(io.TextIOWrapper(gzip.open(src_csv), encoding="utf-8")))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]; data = rows[h + 1:]
ci = {x: i for i, x in enumerate(hdr)}
print("sass rows", len(data), "nvdisasm instr", len(lines))
n = min(len(data), len(lines))
agg = {}; ops = collections.Counter(); tot_i = tot_s = 0
for (loc, s), r in zip(lines[:n], data[:n]):
    ie = int(r[ci["Instructions Executed"]] or 0); sm = int(r[ci["# Samples"]] or 0)
    a = agg.setdefault(loc, [0, 0]); a[0] += ie; a[1] += sm
    tot_i += ie; tot_s += sm
    m = re.match(r"^(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", s)
    ops[m.group(3)] += ie
src = {}
for (f, ln) in agg:
    p = os.path.join(ROOT, "r2p2_b200/csrc", f)
    if f not in src and os.path.exists(p):
        if os.environ.get("SRC_REV"):     # the source as of the profiled build
            src[f] = subprocess.run(["git", "show", "%s:r2p2_b200/csrc/%s" % (os.environ["SRC_REV"], f)], capture_output=True, text=True, cwd=ROOT).stdout.splitlines()
        else:
            src[f] = open(p).read().splitlines()
print("total warp-inst %d samples %d" % (tot_i, tot_s))
for (f, ln), (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[f][ln - 1].strip()[:100] if f in src and ln - 1 < len(src[f]) else ""
    print("%-18s %4d  inst %5.2f%%  samples %5.2f%% | %s" % (f, ln, 100.0 * ie / max(tot_i, 1), 100.0 * sm / max(tot_s, 1), text))
print("---- opcode mix")
for op, c in ops.most_common(30):
    print("%-10s %5.2f%%" % (op, 100.0 * c / tot_i))
def bucket(f, ln):
    if f == "r2p2_math.h":
        if ln < 131: return "math: bits/norm2/sqrt"
        if ln < 317: return "math: sincos"
        if ln < 370: return "math: atan2"
        return "math: exp/tanh/logistic"
    if f in ("r2p2_kernels.cuh", "r2p2_kernel_ws.cuh"):
        return "kernel"
    return "other:" + f
if os.environ.get("BUCKETS"):
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
    for name, (ie, sm) in sorted(b.items(), key=lambda kv: -kv[1][0]):
        print("%-34s inst %5.2f%%  samples %5.2f%%" % (name, 100.0 * ie / tot_i, 100.0 * sm / tot_s))

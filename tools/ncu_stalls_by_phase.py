import csv, gzip, io, sys, re, collections
src_csv, key, sassf = sys.argv[1], sys.argv[2], sys.argv[3]
sass = open(sassf).read().splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and key in l and l.rstrip().endswith(":"))
lines=[]; cur=("?",0)
for l in sass[start+1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    s=l.strip()
    if re.match(r"^(/\*[0-9a-f]+\*/\s+)?(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+", s) and not s.startswith(".") and s.endswith(";"):
        lines.append((cur,s))
rows = list(csv.reader(io.TextIOWrapper(gzip.open(src_csv), encoding="utf-8")))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr=rows[h]; data=rows[h+1:]; ci={x:i for i,x in enumerate(hdr)}
reasons=[x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
def bucket(f, ln):
    if f == "r2p2_math.h":
        if ln < 131: return "norm2/sqrt"
        if ln < 317: return "sincos"
        if ln < 370: return "atan2"
        return "tanh"
    if f == "r2p2_kernels.cuh":
        if ln<160: return "march"     # hot_word, floor_lo, lds
        if 218<=ln<=292: return "plain-march"
        if 292<ln<=345: return "dist-march"
        if 394<=ln<=540: return "march"
        if 540<ln<700: return "ring"
        if 1070<=ln<=1170: return "mlp"
        if ln>=1170: return "k-body:%d" % (ln//50*50)
        return "k-other"
    if f == "r2p2_kernel_ws.cuh":
        if ln<176: return "ws-head"
        if ln<=208: return "ws-window"
        if ln<=246: return "ws-sense"
        if ln<=340: return "ws-control"
        if ln<=430: return "ws-integrate"
        if ln<=450: return "ws-ring"
        return "ws-tail"
    if "sm_32_intr" in f: return "march"   # funnelshift / ldg
    return "other:"+f
agg=collections.defaultdict(lambda: collections.Counter())
tot=collections.Counter()
for (loc,s),r in zip(lines,data):
    b=bucket(*loc)
    for x in reasons:
        v=int(r[ci[x]] or 0)
        agg[b][x]+=v; tot[x]+=v
    agg[b]['inst']+=int(r[ci["Instructions Executed"]] or 0)
T=sum(tot.values()); TI=sum(a['inst'] for a in agg.values())
print("all: " + " ".join("%s %.1f" % (x[6:], 100.0*v/T) for x,v in tot.most_common(10)))
for b,a in sorted(agg.items(), key=lambda kv:-sum(v for k,v in kv[1].items() if k!='inst')):
    s=sum(v for k,v in a.items() if k!='inst')
    if s < 0.005*T: continue
    print("%-14s inst %5.1f%% smp %5.1f%% | " % (b, 100.0*a['inst']/TI, 100.0*s/T) + " ".join("%s %.1f" % (x[6:], 100.0*a[x]/T) for x,_ in tot.most_common(9)))

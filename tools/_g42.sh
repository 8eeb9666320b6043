set -x
( time python bench.py ) > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 200 gpurun_out/bench_r02_n1.err
( time python bench.py --impl reference ) > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
tail -c 200 gpurun_out/bench_r02_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02_bench_launches.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]; ci = {x: i for i, x in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h+1:]:
    if len(r) < len(hdr) or not r[ci['Metric Value']]: continue
    v = float(r[ci['Metric Value']].replace(',', '')); u = r[ci['Metric Unit']]
    ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v if u in ('ms', 'msecond') else v * 1e3
    a = agg[r[ci['Kernel Name']][:110]]; a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print('total %.1f ms over %d launches' % (tot, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('%6.2f%% %9.3f ms %4d  %s' % (100 * a[1] / tot, a[1], a[0], k))
PY

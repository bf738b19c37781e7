# r2x: ncu launch list of the bench command itself at the last code commit (one segment in flight; the CUDA-event shares of the bench
# line must agree with these per-launch shares), summarised per kernel
set -x
mkdir -p gpurun_out
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2x_launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline --no-extras > gpurun_out/r2x_ncu_launches.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/r2x_ncu_launches.log | cut -c1-600
python - <<'PY'
import csv, collections, gzip, io
rows = [l for l in open('gpurun_out/r2x_launches_bench.csv') if not l.startswith('==')]
rd = csv.DictReader(io.StringIO(''.join(rows)))
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = r['Kernel Name'].split('(')[0]
    v = float(r['Metric Value'].replace(',', ''))
    if r.get('Metric Unit', 'ns') in ('us', 'usecond'):
        v *= 1e3
    tot[k][0] += 1; tot[k][1] += v
allns = sum(v[1] for v in tot.values())
with open('gpurun_out/r2x_launches_bench_summary.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline --no-extras (first 4000 launches)\n')
    f.write('kernel,launches,total_ns,share\n')
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write('%s,%d,%d,%.4f\n' % (k, n, ns, ns / allns))
print(open('gpurun_out/r2x_launches_bench_summary.csv').read()[:2500])
PY
gzip -f gpurun_out/r2x_launches_bench.csv

# r2t: segments in flight per GPU after the launch-count reductions of round 2 (--streams 2 / 3 / 4 / 5)
set -x
mkdir -p gpurun_out
for s in 3 4 5 2; do
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 6 --warmup 3 --streams $s > gpurun_out/r2t_bench_streams$s.json 2> gpurun_out/r2t_bench_streams$s.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2t_bench_streams$s.json'))
print('streams $s value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms/step', round(d['ms_per_step'],1))
PY
done

# r2w: the record at the round's last code commit — the driver's own commands: full GPU suite, smoke, bench (--steps 20 --warmup 5).
# (the reference arm was measured by r2z and does not depend on the device code)
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_pytest_gpu.log 2>&1 ) 2>&1 | grep real; tail -3 gpurun_out/r2w_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2w_smoke.log 2>&1; tail -2 gpurun_out/r2w_smoke.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench.json'))
f=d['roofline']['families']
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), d['e2e'].get('mode','')[:30], 'host', round(d['e2e_host_traces']['value'],3), 'fin', round(d['e2e_finish_on_device']['value'],3), 'launches', d['gpu_launches'], 'one-stream ms', round(d['single_segment_latency_ms'],1))
print({k: (round(v['ms_per_step'],2), v['launches_per_step']) for k,v in f.items()})
print('roofline', {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac','traffic')}, d['roofline'].get('issue'))
print('clocks', d['clocks'])
for k in ('config2_cpu_table','config3_b3_b6','config5_stream'):
    print(k, json.dumps(d[k])[-300:])
PY

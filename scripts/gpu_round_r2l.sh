# r2l: fused Merkle tail (one launch per tree for the top levels): parity, stage spans, then the driver's own commands for both arms
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2l_pytest_gpu.log
timeout 300 python tools/stage_report.py > gpurun_out/r2l_stage_report.txt 2>&1; grep "^total" gpurun_out/r2l_stage_report.txt
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
f=d['roofline']['families']
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'fin', round(d['e2e_finish_on_device']['value'],3), 'launches/proof', d['gpu_launches']/d['steps']/3, 'one-stream ms', round(d['single_segment_latency_ms'],1))
print({k: (round(v['ms_per_step'],2), v['launches_per_step']) for k,v in f.items()})
print('cpu_baseline', json.dumps(d['cpu_baseline'])[:600])
for k in ('config2_cpu_table','config3_b3_b6','config5_stream'):
    print(k, json.dumps(d[k])[-420:])
PY
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err ) 2>&1 | grep real
cut -c1-200 gpurun_out/r2l_bench_reference.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_reference.json'))
print(d['steps'], d['ms_per_step'], d['steps_note']); print(d['cpu_baseline'])
PY

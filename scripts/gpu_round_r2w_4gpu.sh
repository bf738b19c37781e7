# r2w (4 GPUs): the driver's scaling command at the last code commit — the N = 4 point of both layouts (segments per GPU; tables of one
# segment over the GPUs), which rounds r2f / r2o (8 GPUs) and r2z / r2w (2 GPUs) do not have
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2w_bench_4gpu.json 2> gpurun_out/r2w_bench_4gpu.err ) 2>&1 | grep real
grep -v "^\[W\|^W1" gpurun_out/r2w_bench_4gpu.err | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench_4gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'host', d['e2e_host_traces']['value'], 'fin', d['e2e_finish_on_device']['value'], 'cpu_baseline', d['cpu_baseline']); print(json.dumps(d['table_sharded'])[:900]); print(json.dumps(d['config5_stream'])[-300:])
PY

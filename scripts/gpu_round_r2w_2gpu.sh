# r2w (2 GPUs): the driver's scaling command at the last code commit (default line with its table_sharded sub-record), and the
# bit-equality of the table-sharded proofs with the one-GPU proofs
set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2w_bench_2gpu.json 2> gpurun_out/r2w_bench_2gpu.err ) 2>&1 | grep real
grep -v "^\[W\|^W1" gpurun_out/r2w_bench_2gpu.err | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench_2gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'host', d['e2e_host_traces']['value'], 'fin', d['e2e_finish_on_device']['value'], 'cpu_baseline', d['cpu_baseline']); print(json.dumps(d['table_sharded'])[:900])
PY
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_check.py --full > gpurun_out/r2w_sharded_check_2gpu.log 2>&1 ) 2>&1 | grep real
grep -i "identical\|equal\|mismatch\|error\|ok" gpurun_out/r2w_sharded_check_2gpu.log | cut -c1-300 | head -8

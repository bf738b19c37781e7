# r2z (2 GPUs): the strong-scaling mode as the main workload, and the default line at N=2
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --parallelism tables --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_2gpu_tables.json 2> gpurun_out/r2z_bench_2gpu_tables.err; cut -c1-700 gpurun_out/r2z_bench_2gpu_tables.json; grep -v "^\[W\|^W1" gpurun_out/r2z_bench_2gpu_tables.err | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench_2gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'fin', d['e2e_finish_on_device']['value']); print(json.dumps(d['table_sharded'])[:700])
PY

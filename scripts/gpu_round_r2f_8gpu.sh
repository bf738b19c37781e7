# r2f (8 GPUs): parity of the table-sharded layout at N=8 + phase breakdown, then the bench line at N=8 (segments per GPU + sub-records)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full --time 12 > gpurun_out/r2f_sharded_8gpu.log 2>&1; grep "sharded_check" gpurun_out/r2f_sharded_8gpu.log | grep -v "rank [1-7] \|between marks" | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_8gpu.json 2> gpurun_out/r2f_bench_8gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_8gpu.json'))
for k in ('value','e2e','e2e_finish_on_device','single_segment_latency_ms','config2_cpu_table','config3_b3_b6','config5_stream','table_sharded'):
    print(k, json.dumps(d.get(k))[:900])
PY
grep -v "^\[W\|^W1" gpurun_out/r2f_bench_8gpu.err | tail -5

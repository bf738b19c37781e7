# r2q: partial-round constants collapsed (one vector before the partial rounds + one scalar per round): pbench, parity, bench
set -x
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk_evm_b200/csrc -I tools tools/pbench.cu -o tools/pbench && timeout 300 tools/pbench) > gpurun_out/r2q_pbench.txt 2>&1; grep "variant\|mismatch" gpurun_out/r2q_pbench.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2q_pytest_gpu.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench.json'))
f=d['roofline']['families']
print('value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'one-stream ms', round(d['single_segment_latency_ms'],1), {k: round(v['ms_per_step'],2) for k,v in f.items()})
PY

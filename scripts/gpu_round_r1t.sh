# one memory pool per context: parity, then the default bench three times
# in each stream-sync mode (variance check)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1t_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1t_pytest_gpu.log
for i in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1t_bench_run_$i.json 2> gpurun_out/r1t_bench_run_$i.err; cat gpurun_out/r1t_bench_run_$i.json | cut -c1-400; done
for i in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline --stream-sync step > gpurun_out/r1t_bench_step_$i.json 2> gpurun_out/r1t_bench_step_$i.err; cat gpurun_out/r1t_bench_step_$i.json | cut -c1-400; done

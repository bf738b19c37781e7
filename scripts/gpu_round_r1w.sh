# device-side upload gate + pinned staging for the small transfers: parity, then e2e with and without upload-ahead at 1 / 2 / 3 streams
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1w_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1w_pytest_gpu.log
for pf in 0 1; do for s in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline --no-kernel-events --streams $s --prefetch $pf > gpurun_out/r1w_bench_pf${pf}_${s}streams.json 2> gpurun_out/r1w_bench_pf${pf}_${s}streams.err; cat gpurun_out/r1w_bench_pf${pf}_${s}streams.json | cut -c1-250; tail -2 gpurun_out/r1w_bench_pf${pf}_${s}streams.err; done; done

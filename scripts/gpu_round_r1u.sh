# paired-challenge CTL evaluation (quotient + helper columns): parity, bench, and the number of segments in flight (2, 3, 4)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1u_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1u_bench.json 2> gpurun_out/r1u_bench.err; cat gpurun_out/r1u_bench.json | cut -c1-300; tail -2 gpurun_out/r1u_bench.err
for s in 3 4; do timeout 600 python bench.py --no-cpu-baseline --streams $s --no-kernel-events > gpurun_out/r1u_bench_${s}streams.json 2> gpurun_out/r1u_bench_${s}streams.err; cat gpurun_out/r1u_bench_${s}streams.json | cut -c1-300; done

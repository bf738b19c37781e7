# upload gate (one H2D chain at a time), plain-cell descriptor fast path: parity + bench x2 + one-stream bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1p_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1p_pytest_gpu.log
for i in a b; do timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1p_bench_$i.json 2> gpurun_out/r1p_bench_$i.err; cat gpurun_out/r1p_bench_$i.json; tail -3 gpurun_out/r1p_bench_$i.err; done
timeout 900 python bench.py --no-cpu-baseline --streams 1 > gpurun_out/r1p_bench_1stream.json 2> gpurun_out/r1p_bench_1stream.err; cat gpurun_out/r1p_bench_1stream.json
timeout 900 ncu --set full --clock-control none -k regex:"quotient_kernel|helper_kernel" -c 30 -o gpurun_out/r1p_prof_quot python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1p_ncu_quot.log 2>&1
ncu -i gpurun_out/r1p_prof_quot.ncu-rep --page raw --csv > gpurun_out/r1p_prof_quot.raw.csv
rm -f gpurun_out/r1p_prof_quot.ncu-rep

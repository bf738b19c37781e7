# after reverting the quotient kernels to 128-thread blocks (big blocks starved with two segments in flight): parity + bench twice
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1n_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1n_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; cat gpurun_out/r1n_bench.json; tail -3 gpurun_out/r1n_bench.err
timeout 900 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r1n_bench_b.json 2> gpurun_out/r1n_bench_b.err; cat gpurun_out/r1n_bench_b.json
timeout 900 ncu --set full --clock-control none -k regex:quotient_kernel -c 9 -o gpurun_out/r1n_prof_quot python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1n_ncu_quot.log 2>&1
ncu -i gpurun_out/r1n_prof_quot.ncu-rep --page raw --csv > gpurun_out/r1n_prof_quot.raw.csv
rm -f gpurun_out/r1n_prof_quot.ncu-rep

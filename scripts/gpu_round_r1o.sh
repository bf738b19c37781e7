# leaf-hash wave balancing (resident blocks per SM sweep), lock-step CTL loops + 256-thread Memory quotient blocks, full-size verification tests
set -x
mkdir -p gpurun_out
for b in 9 8 7 6 5 ""; do ZKGPU_LEAF_BLOCKS=$b python tools/leafbench.py >> gpurun_out/r1o_leafbench.jsonl 2>&1; done; cat gpurun_out/r1o_leafbench.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1o_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1o_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; cat gpurun_out/r1o_bench.json; tail -3 gpurun_out/r1o_bench.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1o_bench_b.json 2> gpurun_out/r1o_bench_b.err; cat gpurun_out/r1o_bench_b.json
timeout 900 ncu --set full --clock-control none -k regex:quotient_kernel -c 9 -o gpurun_out/r1o_prof_quot python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1o_ncu_quot.log 2>&1
ncu -i gpurun_out/r1o_prof_quot.ncu-rep --page raw --csv > gpurun_out/r1o_prof_quot.raw.csv
rm -f gpurun_out/r1o_prof_quot.ncu-rep

# parity + bench after: per-table register budgets restored, batched inversions in the aux kernels, lazy-accumulator fri_combine,
# NTT first/last rounds fused with the global loads/stores
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1l_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1l_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; cat gpurun_out/r1l_bench.json; tail -3 gpurun_out/r1l_bench.err
timeout 900 python bench.py --streams 3 --no-cpu-baseline > gpurun_out/r1l_bench_3streams.json 2> gpurun_out/r1l_bench_3streams.err; cat gpurun_out/r1l_bench_3streams.json
python tools/microbench.py > gpurun_out/r1l_microbench.jsonl 2>&1; cat gpurun_out/r1l_microbench.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 4 -c 4 -o gpurun_out/r1l_prof_ntt python tools/microbench.py > gpurun_out/r1l_ncu_ntt.log 2>&1
ncu -i gpurun_out/r1l_prof_ntt.ncu-rep --page raw --csv > gpurun_out/r1l_prof_ntt.raw.csv
ncu -i gpurun_out/r1l_prof_ntt.ncu-rep --page details > gpurun_out/r1l_prof_ntt.details.txt
rm -f gpurun_out/r1l_prof_ntt.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"fri_combine|helper_kernel|eval_columns" -c 12 -o gpurun_out/r1l_prof_misc python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1l_ncu_misc.log 2>&1
ncu -i gpurun_out/r1l_prof_misc.ncu-rep --page raw --csv > gpurun_out/r1l_prof_misc.raw.csv
rm -f gpurun_out/r1l_prof_misc.ncu-rep
ls -la gpurun_out

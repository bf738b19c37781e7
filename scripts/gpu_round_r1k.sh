# parity tests + bench after: index-addressed Keccak quotient, upload/commit overlap, radix-16 power-of-two NTT rounds
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1k_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1k_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err; cat gpurun_out/r1k_bench.json; tail -3 gpurun_out/r1k_bench.err
timeout 900 python bench.py --streams 1 --no-cpu-baseline > gpurun_out/r1k_bench_1stream.json 2> gpurun_out/r1k_bench_1stream.err; cat gpurun_out/r1k_bench_1stream.json
python tools/microbench.py > gpurun_out/r1k_microbench.jsonl 2>&1; cat gpurun_out/r1k_microbench.jsonl
timeout 900 ncu --set full --clock-control none -k regex:quotient_kernel -c 9 -o gpurun_out/r1k_prof_quot python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1k_ncu_quot.log 2>&1
ncu -i gpurun_out/r1k_prof_quot.ncu-rep --page raw --csv > gpurun_out/r1k_prof_quot.raw.csv
rm -f gpurun_out/r1k_prof_quot.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 4 -c 4 -o gpurun_out/r1k_prof_ntt python tools/microbench.py > gpurun_out/r1k_ncu_ntt.log 2>&1
ncu -i gpurun_out/r1k_prof_ntt.ncu-rep --page raw --csv > gpurun_out/r1k_prof_ntt.raw.csv
ncu -i gpurun_out/r1k_prof_ntt.ncu-rep --page details > gpurun_out/r1k_prof_ntt.details.txt
rm -f gpurun_out/r1k_prof_ntt.ncu-rep
ls -la gpurun_out

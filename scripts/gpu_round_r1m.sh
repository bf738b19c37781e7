# parity + bench after: limb-form Poseidon partial rounds, 384-thread lock-step Cpu/Arithmetic quotient blocks, NTT plan (8-bit strided
# passes) + one item per thread, eval_columns prefetch, smaller PoW batch
set -x
mkdir -p gpurun_out
./tools/pbench 4 > gpurun_out/r1m_pbench.txt 2>&1; cat gpurun_out/r1m_pbench.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1m_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1m_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; cat gpurun_out/r1m_bench.json; tail -3 gpurun_out/r1m_bench.err
python tools/microbench.py > gpurun_out/r1m_microbench.jsonl 2>&1; cat gpurun_out/r1m_microbench.jsonl
timeout 900 ncu --set full --clock-control none -k regex:quotient_kernel -c 9 -o gpurun_out/r1m_prof_quot python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1m_ncu_quot.log 2>&1
ncu -i gpurun_out/r1m_prof_quot.ncu-rep --page raw --csv > gpurun_out/r1m_prof_quot.raw.csv
rm -f gpurun_out/r1m_prof_quot.ncu-rep
ls -la gpurun_out

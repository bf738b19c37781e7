# full record of the round: parity tests, smoke, both bench arms, register-budget sweep of the Cpu/Arithmetic quotient kernels,
# ncu launch list of the bench command, --set full captures of the leaf hash, the NTT passes and the nine quotient kernels
set -x
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1q_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1q_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r1q_smoke.log 2>&1; tail -2 gpurun_out/r1q_smoke.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1q_bench_reference.json 2> gpurun_out/r1q_bench_reference.err; cat gpurun_out/r1q_bench_reference.json
timeout 900 python bench.py > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; cat gpurun_out/r1q_bench.json; tail -3 gpurun_out/r1q_bench.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1q_bench_b.json 2> gpurun_out/r1q_bench_b.err; cat gpurun_out/r1q_bench_b.json
timeout 900 python bench.py --streams 1 --no-cpu-baseline > gpurun_out/r1q_bench_1stream.json 2> gpurun_out/r1q_bench_1stream.err; cat gpurun_out/r1q_bench_1stream.json
for mb in 5 8; do ZKGPU_QUOT_MINB=$mb timeout 600 python bench.py --streams 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1q_bench_minb$mb.json 2> gpurun_out/r1q_bench_minb$mb.err; cat gpurun_out/r1q_bench_minb$mb.json; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r1q_launches_segment.csv python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1q_ncu_launches.log 2>&1
tail -2 gpurun_out/r1q_ncu_launches.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"leaf_hash|ntt_pass|quotient_kernel" -c 60 -o gpurun_out/r1q_prof python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1q_ncu_prof.log 2>&1
ncu -i gpurun_out/r1q_prof.ncu-rep --page raw --csv > gpurun_out/r1q_prof.raw.csv
ncu -i gpurun_out/r1q_prof.ncu-rep --page details -k regex:leaf_hash > gpurun_out/r1q_prof_leaf_hash.details.txt
ncu -i gpurun_out/r1q_prof.ncu-rep --page details -k regex:ntt_pass > gpurun_out/r1q_prof_ntt.details.txt
rm -f gpurun_out/r1q_prof.ncu-rep
ls -la gpurun_out | tail -30

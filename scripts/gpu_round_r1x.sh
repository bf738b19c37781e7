# final record of round 1: parity, smoke, both bench arms (default = three segments in flight), ncu launch list of the bench command,
# --set full of the leaf hash / NTT passes / quotient kernels
set -x
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1x_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1x_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r1x_smoke.log 2>&1; tail -2 gpurun_out/r1x_smoke.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1x_bench_reference.json 2> gpurun_out/r1x_bench_reference.err; cat gpurun_out/r1x_bench_reference.json
timeout 900 python bench.py > gpurun_out/r1x_bench.json 2> gpurun_out/r1x_bench.err; cat gpurun_out/r1x_bench.json; tail -3 gpurun_out/r1x_bench.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r1x_bench_b.json 2> gpurun_out/r1x_bench_b.err; cat gpurun_out/r1x_bench_b.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r1x_launches_segment.csv python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1x_ncu_launches.log 2>&1
tail -2 gpurun_out/r1x_ncu_launches.log
timeout 900 ncu --set full --clock-control none -k regex:"leaf_hash|ntt_pass|quotient_kernel" -c 45 -o gpurun_out/r1x_prof python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1x_ncu_prof.log 2>&1
ncu -i gpurun_out/r1x_prof.ncu-rep --page raw --csv > gpurun_out/r1x_prof.raw.csv
rm -f gpurun_out/r1x_prof.ncu-rep
ls -la gpurun_out | tail -12

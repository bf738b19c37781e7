# One gpurun call: parity tests, smoke, both bench arms, the ncu launch list of the bench command and a --set full capture
# of the dominant kernel (leaf_hash). Outputs under gpurun_out/r1j_*.
set -x
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1j_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r1j_smoke.log 2>&1; tail -2 gpurun_out/r1j_smoke.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1j_bench_reference.json 2> gpurun_out/r1j_bench_reference.err; cat gpurun_out/r1j_bench_reference.json
timeout 900 python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; cat gpurun_out/r1j_bench.json
timeout 900 python bench.py --streams 1 --no-cpu-baseline > gpurun_out/r1j_bench_1stream.json 2> gpurun_out/r1j_bench_1stream.err; cat gpurun_out/r1j_bench_1stream.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/r1j_launches_segment.csv python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1j_ncu_launches.log 2>&1
tail -2 gpurun_out/r1j_ncu_launches.log
for k in leaf_hash merkle_level; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 4 -o gpurun_out/r1j_prof_$k python bench.py --steps 1 --warmup 0 --streams 1 --no-cpu-baseline > gpurun_out/r1j_ncu_$k.log 2>&1
ncu -i gpurun_out/r1j_prof_$k.ncu-rep --page raw --csv > gpurun_out/r1j_prof_$k.raw.csv
ncu -i gpurun_out/r1j_prof_$k.ncu-rep --page details > gpurun_out/r1j_prof_$k.details.txt
rm -f gpurun_out/r1j_prof_$k.ncu-rep
done
ls -la gpurun_out

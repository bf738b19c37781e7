# first GPU call of round 2: everything written after round 1's last GPU run (trace finishing, scheduler, C++ mirror GPU leg, pinned
# memory, stage spans) meets a B200 for the first time.  No -x: every failure is wanted in one pass.
set -x
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"; lscpu | grep -o "avx512[a-z]*" | sort -u | tr '\n' ' '
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -30 gpurun_out/r2a_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk_evm_b200/csrc -I tools tools/pbench.cu -o tools/pbench && timeout 300 tools/pbench) > gpurun_out/r2a_pbench.txt 2>&1; tail -12 gpurun_out/r2a_pbench.txt
timeout 900 python bench.py --finish-on-device --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
timeout 600 ncu --set full --clock-control none -k regex:"keccak_trace_kernel|logic_trace_kernel|arith_rc|memory_finish" -c 8 -o gpurun_out/r2a_prof_tracegen python -m pytest tests/test_gpu_trace_gen.py -m gpu -q -k "170 or 5000 or 17 or 900" > gpurun_out/r2a_ncu_tracegen.log 2>&1
ncu -i gpurun_out/r2a_prof_tracegen.ncu-rep --page raw --csv > gpurun_out/r2a_prof_tracegen.raw.csv
rm -f gpurun_out/r2a_prof_tracegen.ncu-rep
ls -la gpurun_out | tail -12

# first GPU call of round 2 (everything below was written after round 1's GPU budget ran out and is checked on the CPU only: host
# builds, the PTX interpreter).  Order: parity (incl. tests/test_gpu_trace_gen.py, never run on a GPU yet) -> smoke -> the limb-form
# Poseidon variant (tools/pbench variant 4: one run decides whether it goes into poseidon_fast.cuh) -> bench with the device-side
# trace finishing leg -> launch list + ncu of the two trace-finishing kernels.
set -x
mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk_evm_b200/csrc -I tools tools/pbench.cu -o tools/pbench && timeout 300 tools/pbench) > gpurun_out/r2a_pbench.txt 2>&1; tail -12 gpurun_out/r2a_pbench.txt
timeout 900 python bench.py --finish-on-device > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err; cut -c1-300 gpurun_out/r2a_bench_reference.json
timeout 600 ncu --set full --clock-control none -k regex:"keccak_trace_kernel|logic_trace_kernel|arith_rc|memory_finish" -c 8 -o gpurun_out/r2a_prof_tracegen python -m pytest tests/test_gpu_trace_gen.py -m gpu -q -k "170 or 5000 or 17 or 900" > gpurun_out/r2a_ncu_tracegen.log 2>&1
ncu -i gpurun_out/r2a_prof_tracegen.ncu-rep --page raw --csv > gpurun_out/r2a_prof_tracegen.raw.csv
rm -f gpurun_out/r2a_prof_tracegen.ncu-rep
ls -la gpurun_out | tail -12

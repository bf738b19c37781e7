# r2b: limb-form partial rounds in the production permutation (parity), stage spans of one segment, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2b_pytest_gpu.log
timeout 300 python tools/stage_report.py > gpurun_out/r2b_stage_report.txt 2>&1; grep -A40 "by stage name" gpurun_out/r2b_stage_report.txt | head -40
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cut -c1-400 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err

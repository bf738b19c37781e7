# upload-ahead (zkgpu_segment_upload / zkgpu_prove_segment_uploaded): parity, then the bench with 2 and 3 segments in flight
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1v_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1v_pytest_gpu.log
for s in 2 3; do for i in a b; do timeout 600 python bench.py --no-cpu-baseline --streams $s > gpurun_out/r1v_bench_${s}streams_$i.json 2> gpurun_out/r1v_bench_${s}streams_$i.err; cat gpurun_out/r1v_bench_${s}streams_$i.json | cut -c1-300; tail -2 gpurun_out/r1v_bench_${s}streams_$i.err; done; done
timeout 600 python bench.py --no-cpu-baseline --streams 1 > gpurun_out/r1v_bench_1stream.json 2> gpurun_out/r1v_bench_1stream.err; cat gpurun_out/r1v_bench_1stream.json | cut -c1-300

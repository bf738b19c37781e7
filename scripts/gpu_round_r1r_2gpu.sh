# two GPUs of one box: the driver's launch line for N = 2 (segments, weak scaling) and the table-sharded layout (strong scaling)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1r_bench_2gpu_segments.json 2> gpurun_out/r1r_bench_2gpu_segments.err; cat gpurun_out/r1r_bench_2gpu_segments.json; tail -3 gpurun_out/r1r_bench_2gpu_segments.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --parallelism tables --no-cpu-baseline > gpurun_out/r1r_bench_2gpu_tables.json 2> gpurun_out/r1r_bench_2gpu_tables.err; cat gpurun_out/r1r_bench_2gpu_tables.json; tail -3 gpurun_out/r1r_bench_2gpu_tables.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r1r_bench_2gpu_reference.json 2> gpurun_out/r1r_bench_2gpu_reference.err; cat gpurun_out/r1r_bench_2gpu_reference.json

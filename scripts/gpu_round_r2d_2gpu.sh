# r2d (2 GPUs): bench at N=2 with the table_sharded sub-record
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; tail -c 2500 gpurun_out/r2d_bench_2gpu.json; grep -v "^\[W\|^W1" gpurun_out/r2d_bench_2gpu.err | tail -8

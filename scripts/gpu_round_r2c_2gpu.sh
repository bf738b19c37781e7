# r2c (2 GPUs): split commitments — single-device simulation test, 2-rank parity (sharded == one GPU, word for word), bench at N=2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -q -x > gpurun_out/r2c_pytest_shard.log 2>&1; tail -5 gpurun_out/r2c_pytest_shard.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full > gpurun_out/r2c_sharded_check.log 2>&1; grep "sharded_check\|Error\|error" gpurun_out/r2c_sharded_check.log | head -20; tail -5 gpurun_out/r2c_sharded_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; tail -c 3000 gpurun_out/r2c_bench_2gpu.json; tail -5 gpurun_out/r2c_bench_2gpu.err

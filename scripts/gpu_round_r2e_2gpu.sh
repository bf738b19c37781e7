set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full --time 5 > gpurun_out/r2e_sharded_time.log 2>&1; grep "sharded_check" gpurun_out/r2e_sharded_time.log

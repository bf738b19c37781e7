set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full --time 12 > gpurun_out/r2n_sharded.log 2>&1; grep "sharded_check" gpurun_out/r2n_sharded.log | grep "identical\|MISMATCH\|back-to-back\|per proof" | grep -v "rank 1" | cut -c1-1300
grep -i "error\|Traceback" gpurun_out/r2n_sharded.log | head -5

set -x
mkdir -p gpurun_out
for mode in lazy upfront; do
ZK_SHARD_BEGIN=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full --time 24 > gpurun_out/r2i_$mode.log 2>&1; echo "== $mode"; grep "back-to-back\|identical\|MISMATCH" gpurun_out/r2i_$mode.log | grep -v "rank 1" | cut -c1-400
done

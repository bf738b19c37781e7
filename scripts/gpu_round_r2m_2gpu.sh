# r2m (2 GPUs): alpha-independent constraint values precomputed in phase 2 — parity, then the sharded timing with and without
set -x
mkdir -p gpurun_out
true
for pre in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py --full --time 8 --precompute $pre > gpurun_out/r2m_sharded_pre$pre.log 2>&1; echo "== precompute $pre"; grep "sharded_check" gpurun_out/r2m_sharded_pre$pre.log | grep "back-to-back\|walls" | grep "rank 0" | cut -c1-400
done

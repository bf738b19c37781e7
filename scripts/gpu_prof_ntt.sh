set -x
cat > /tmp/nttb.py <<'PY'
import sys; sys.path.insert(0, ".")
import zk_evm_b200 as zk
ctx = zk.Context(0)
for (nc, lg) in ((2431, 17), (85, 19), (30, 21)):
    ms, l = ctx.bench_ntt(nc, 1 << lg, 3)
    print(nc, lg, "ms", ms, "launches", l, "GB/s algorithmic", 16.0 * nc * (1 << lg) / ms / 1e6, flush=True)
PY
python /tmp/nttb.py
ncu --set full --clock-control none -k regex:ntt_dif_pass -s 2 -c 2 -o gpurun_out/prof_ntt_r1g python /tmp/nttb.py > gpurun_out/ncu_ntt.log 2>&1
ncu -i gpurun_out/prof_ntt_r1g.ncu-rep --page raw --csv > gpurun_out/prof_ntt_r1g.raw.csv
ncu -i gpurun_out/prof_ntt_r1g.ncu-rep --page details > gpurun_out/prof_ntt_r1g.details.txt
rm -f gpurun_out/prof_ntt_r1g.ncu-rep

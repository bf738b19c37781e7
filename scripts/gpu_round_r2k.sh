# r2k: NTT tiles staged by the TMA unit (cp.async.bulk + mbarrier): parity, A/B of the NTT family time, ncu of the new kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2k_pytest_gpu.log
for tma in 1 0; do
ZKGPU_NTT_TMA=$tma timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2k_bench_tma$tma.json 2> gpurun_out/r2k_bench_tma$tma.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_tma$tma.json'))
f=d['roofline']['families']
print('TMA=$tma value', round(d['value'],3), 'one-stream ms', round(d['single_segment_latency_ms'],1), {k: round(v['ms_per_step'],2) for k,v in f.items()}, 'ntt GB/s', round(f['ntt']['achieved_GBps']))
PY
done
timeout 600 ncu --set full --clock-control none -k regex:ntt_pass --launch-skip 120 -c 12 -o gpurun_out/r2k_prof_ntt python tools/stage_report.py --reps 1 > gpurun_out/r2k_ncu_ntt.log 2>&1
ncu -i gpurun_out/r2k_prof_ntt.ncu-rep --page raw --csv > gpurun_out/r2k_prof_ntt.raw.csv
rm -f gpurun_out/r2k_prof_ntt.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2k_prof_ntt.raw.csv')))
h=rows[0]
want=['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active']
idx=[h.index(w) for w in want if w in h]
for r in rows[:14]:
    print([r[i][:40] for i in idx])
PY

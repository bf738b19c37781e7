# r2p: GPU parity of the new executing segments + ncu of the production leaf-hash kernel (instructions per permutation, DRAM traffic) and the launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_segment.py tests/test_gpu_shard.py -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r2p_launches_segment.csv python tools/stage_report.py --reps 1 > gpurun_out/r2p_ncu_launches.log 2>&1; tail -2 gpurun_out/r2p_ncu_launches.log
timeout 600 ncu --set full --clock-control none -k regex:"leaf_hash_kernel|merkle_tail|merkle_level" --launch-skip 0 -c 60 -o gpurun_out/r2p_prof_leaf python tools/stage_report.py --reps 1 > gpurun_out/r2p_ncu_leaf.log 2>&1
ncu -i gpurun_out/r2p_prof_leaf.ncu-rep --page raw --csv > gpurun_out/r2p_prof_leaf.raw.csv
rm -f gpurun_out/r2p_prof_leaf.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_prof_leaf.raw.csv')))
h=rows[0]
want=['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct']
idx=[h.index(w) for w in want if w in h]
best=sorted(rows[2:], key=lambda r: -float(r[h.index('gpu__time_duration.sum')]))[:6]
for r in [rows[0], rows[1]]+best:
    print([r[i][:34] for i in idx])
PY

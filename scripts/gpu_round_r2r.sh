# r2r: ncu --set full of THE dominant launch (leaf hash of the Keccak trace LDE: 2^18 leaves x 2431 columns) with the round-2 permutation; bench line with the e2e selection
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:leaf_hash_kernel --launch-skip 8 -c 1 -o gpurun_out/r2r_prof_leaf_keccak python tools/stage_report.py --reps 1 > gpurun_out/r2r_ncu.log 2>&1
ncu -i gpurun_out/r2r_prof_leaf_keccak.ncu-rep --page raw --csv > gpurun_out/r2r_prof_leaf_keccak.raw.csv
ncu -i gpurun_out/r2r_prof_leaf_keccak.ncu-rep --page details > gpurun_out/r2r_prof_leaf_keccak.details.txt
rm -f gpurun_out/r2r_prof_leaf_keccak.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2r_prof_leaf_keccak.raw.csv')))
h=rows[0]
for w in ['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct']:
    if w in h: print(w, rows[1][h.index(w)], rows[2][h.index(w)])
PY
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print({k: d[k] for k in ('value','e2e','e2e_host_traces')}); print(d.get('e2e_finish_on_device'))
PY
tail -3 gpurun_out/r2r_bench.err

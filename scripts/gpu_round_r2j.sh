# r2j: quotient kernels with the row / next-row co-location: parity, stage spans, DRAM traffic of the nine quotient kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2j_pytest_gpu.log
timeout 300 python tools/stage_report.py > gpurun_out/r2j_stage_report.txt 2>&1; grep -A24 "by stage name" gpurun_out/r2j_stage_report.txt; grep "^total" gpurun_out/r2j_stage_report.txt
timeout 600 ncu --set full --clock-control none -k regex:quotient_kernel --launch-skip 18 -c 9 -o gpurun_out/r2j_prof_quot python tools/stage_report.py --reps 1 > gpurun_out/r2j_ncu_quot.log 2>&1
ncu -i gpurun_out/r2j_prof_quot.ncu-rep --page raw --csv > gpurun_out/r2j_prof_quot.raw.csv
rm -f gpurun_out/r2j_prof_quot.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2j_prof_quot.raw.csv')))
h=rows[0]
def col(n): return h.index(n)
for r in rows[2:]:
    try:
        print(r[col('Kernel Name')][:60], 'ms', float(r[col('gpu__time_duration.sum')])/1e6, 'dram rd GB', float(r[col('dram__bytes_read.sum')])/1e9, 'wr GB', float(r[col('dram__bytes_write.sum')])/1e9)
    except Exception as e:
        print('row', e)
PY

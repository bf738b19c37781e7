set -x
ncu --set full --clock-control none -k regex:quotient_kernel -c 9 -o gpurun_out/prof_quot_r1h python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_quot.log 2>&1
ncu -i gpurun_out/prof_quot_r1h.ncu-rep --page raw --csv > gpurun_out/prof_quot_r1h.raw.csv
rm -f gpurun_out/prof_quot_r1h.ncu-rep
tail -2 gpurun_out/ncu_quot.log

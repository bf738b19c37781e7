set -x
./tools/pbench 4 > gpurun_out/pbench.log 2>&1; cat gpurun_out/pbench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
for k in leaf_hash ntt_dif quotient; do
ncu --set full --clock-control none -k regex:$k -c 1 -o gpurun_out/prof_${k}_r1b python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
ncu -i gpurun_out/prof_${k}_r1b.ncu-rep --page raw --csv > gpurun_out/prof_${k}_r1b.raw.csv
ncu -i gpurun_out/prof_${k}_r1b.ncu-rep --page details > gpurun_out/prof_${k}_r1b.details.txt
rm -f gpurun_out/prof_${k}_r1b.ncu-rep
done
ls -la gpurun_out; du -sh gpurun_out

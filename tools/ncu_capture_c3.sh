# ncu evidence for the 4K10 inter clip (c3): launch list + --set full of K2 (run on the GPU box through gpurun)
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload c3_4k10_inter"
timeout 300 $B > gpurun_out/plain_r1d_c3.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 800 --csv --log-file gpurun_out/launches_r1d_c3.csv $B > gpurun_out/ncu_r1d_c3_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inter_pred -s 12 -c 2 -f -o gpurun_out/prof_k2_r1d $B > gpurun_out/ncu_r1d_c3_b.log 2>&1
ls -la gpurun_out/prof_k2_r1d.ncu-rep gpurun_out/launches_r1d_c3.csv

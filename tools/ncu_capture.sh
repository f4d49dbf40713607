# ncu evidence for the round-1 "c" state (run on the GPU box through gpurun): launch list + --set full of the top kernels.
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain_r1d.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 700 --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/ncu_r1d_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:intra_unit -s 65 -c 2 -f -o gpurun_out/prof_k3_r1d $B > gpurun_out/ncu_r1d_b.log 2>&1


ls -la gpurun_out/*.ncu-rep

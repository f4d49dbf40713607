# Round-2 ncu evidence (run on the GPU box through gpurun, after the plain command has exited 0): launch lists of c3 and c2,
# then `--set full` captures of the dominant kernels.
TAG=${1:-r2}
bash tools/ncu_launches.sh $TAG c3_4k10_inter 1700 1000
bash tools/ncu_launches.sh $TAG c2_intra_1080p8 1300 700
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload c3_4k10_inter"
for k in inter_pred intra_unit cdef_kernel lr_kernel deblock_kernel itx_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 14 -c 3 -f -o gpurun_out/prof_${TAG}_c3_$k $B > gpurun_out/ncu_${TAG}_c3_$k.log 2>&1
done
B2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload c2_intra_1080p8"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:intra_unit -s 65 -c 2 -f -o gpurun_out/prof_${TAG}_c2_intra_unit $B2 > gpurun_out/ncu_${TAG}_c2_intra.log 2>&1
ls -la gpurun_out/prof_${TAG}_*.ncu-rep

# ncu --set full captures of the two dominant kernel families of the final build on the 60-frame c3 sequence (run through gpurun, after
# the same command has exited 0 without ncu): usage bash tools/ncu_final.sh TAG
TAG=${1:-r2z}
export AV1R_BENCH_REPEAT=1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload c3_4k10_inter"
timeout 300 $B > gpurun_out/plain_${TAG}_c3x1.log 2>&1 || exit 1
for k in inter_pred intra_unit; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 14 -c 3 -f -o gpurun_out/prof_${TAG}_c3_$k $B > gpurun_out/ncu_${TAG}_c3_$k.log 2>&1
done
ls -la gpurun_out/prof_${TAG}_*.ncu-rep

# ncu --set full capture of the wavefront kernel on the all-key-frame clip (after the plain command has exited 0)
TAG=${1:-r2k}
B2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload c2_intra_1080p8"
$B2 > gpurun_out/ncu_${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:intra_unit -s 65 -c 2 -f -o gpurun_out/prof_${TAG}_c2_intra_unit $B2 > gpurun_out/ncu_${TAG}_c2_intra.log 2>&1
ls -la gpurun_out/prof_${TAG}_*.ncu-rep

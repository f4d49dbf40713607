for w in c2_intra_1080p8 c2_small; do
  AV1R_PROFILE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/probe_$w.json 2> gpurun_out/probe_$w.err
  grep "engine prof" gpurun_out/probe_$w.err | head -2
done
nvidia-smi --query-gpu=utilization.gpu --format=csv -lms 100 > gpurun_out/probe_util.csv &
SMI=$!
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload c2_intra_1080p8 > /dev/null 2>&1
kill $SMI

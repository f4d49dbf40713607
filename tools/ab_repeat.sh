# e2e / value of the 4K clips against the number of back-to-back repetitions of the 60-frame sequence (bench.py CLIP_REPEAT)
for W in c3_4k10_inter c4_4k10_grain; do
for r in 1 2 4 8; do
  AV1R_BENCH_REPEAT=$r python bench.py --workload $W --steps 5 --warmup 3 --no-per-config 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('repeat $r $W value',round(d['value']),'fenced',round(d.get('value_fenced_steps') or 0),'e2e',round(d['e2e']['value']),'dav1d',round(d['cpu_baseline']['value']),'parse_ms/step',round(d['e2e']['host_parse_ms_per_step']))" >> gpurun_out/ab_repeat.txt
done
done
cat gpurun_out/ab_repeat.txt

# on-box A/B of the host parser (parse-only, no kernels): base build vs current, 1 thread and all threads; then c2 bench standalone
cp av1-go_b200/lib/libav1r.so /tmp/cur.so
for round in 1 2; do
for v in base cur; do
  if [ $v = cur ]; then cp /tmp/cur.so av1-go_b200/lib/libav1r.so; else cp av1-go_b200/lib/libav1r_$v.so av1-go_b200/lib/libav1r.so; fi
  for c in c2 c3 c1; do
    echo -n "$v $c 1-thread: " >> gpurun_out/ab_parse.txt; python -m tools.parse_bench $c --reps 3 --no-tiles --threads 1 2>/dev/null | head -1 >> gpurun_out/ab_parse.txt
    echo -n "$v $c all: " >> gpurun_out/ab_parse.txt; python -m tools.parse_bench $c --reps 3 2>/dev/null | head -1 >> gpurun_out/ab_parse.txt
  done
done
done
cp /tmp/cur.so av1-go_b200/lib/libav1r.so
python bench.py --workload c2_intra_1080p8 --steps 5 --warmup 3 --no-per-config 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 standalone value',round(d['value']),'e2e',round(d['e2e']['value']),'parse_ms',round(d['e2e']['host_parse_ms_per_step']))" >> gpurun_out/ab_parse.txt
cat gpurun_out/ab_parse.txt

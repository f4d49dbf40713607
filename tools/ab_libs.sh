# on-box comparison of library builds (av1-go_b200/lib/libav1r_<tag>.so) on the stage table of a workload
# usage: bash tools/ab_libs.sh WORKLOAD tag...
W=$1; shift
cp av1-go_b200/lib/libav1r.so /tmp/cur.so
for round in $(seq ${ROUNDS:-2}); do
for v in cur "$@"; do
  if [ $v = cur ]; then cp /tmp/cur.so av1-go_b200/lib/libav1r.so; else cp av1-go_b200/lib/libav1r_$v.so av1-go_b200/lib/libav1r.so; fi
  python bench.py --workload $W --steps 5 --warmup 3 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stages']
print('$v $W value',round(d['value']),'fenced',round(d.get('value_fenced_steps') or 0),'resident',round(d.get('value_hbm_resident') or 0),'e2e',round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in st.items()})" >> gpurun_out/ab_libs.txt
done
done
cp /tmp/cur.so av1-go_b200/lib/libav1r.so
cat gpurun_out/ab_libs.txt

# on-box comparison of environment switches of one library build on the stage table of a workload
# usage: bash tools/ab_env.sh WORKLOAD "VAR=val [VAR2=val]" ...      ("-" = no switch)
W=$1; shift
for round in $(seq ${ROUNDS:-2}); do
for v in "-" "$@"; do
  e=""; [ "$v" != "-" ] && e="$v"
  env $e python bench.py --workload $W --steps 5 --warmup 3 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stages']
print('[$v] $W value',round(d['value']),'fenced',round(d.get('value_fenced_steps') or 0),'e2e',round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in st.items()})" >> gpurun_out/ab_env.txt
done
done
cat gpurun_out/ab_env.txt

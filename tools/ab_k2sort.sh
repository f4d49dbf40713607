# A/B of the K2 work-item order (AV1R_K2_NOSORT=1: decode order) on the stage table of c3 / c1 / c4
for round in 1 2; do
for v in NOSORT SORT; do
  for w in c3_4k10_inter c4_4k10_grain c1_1080p8; do
    if [ $v = NOSORT ]; then export AV1R_K2_NOSORT=1; else unset AV1R_K2_NOSORT; fi
    python bench.py --workload $w --steps 5 --warmup 3 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stages']
print('$v $w value',round(d['value']),'inter',round(st['inter']['ms_per_step'],2),'intra',round(st['intra']['ms_per_step'],2))" >> gpurun_out/ab_k2sort.txt
  done
done
done
cat gpurun_out/ab_k2sort.txt

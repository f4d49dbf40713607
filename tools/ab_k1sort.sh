# A/B of the K1 work order (AV1R_K1_NOSORT=1: decode order) on the stage table of c3 / c2 / c1
for round in 1 2; do
for v in NOSORT SORT; do
  for w in c3_4k10_inter c2_intra_1080p8 c1_1080p8; do
    if [ $v = NOSORT ]; then export AV1R_K1_NOSORT=1; else unset AV1R_K1_NOSORT; fi
    python bench.py --workload $w --steps 5 --warmup 3 --no-per-config --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stages']
print('$v $w value',round(d['value']),'itx',round(st['itx']['ms_per_step'],2),'intra',round(st['intra']['ms_per_step'],2))" >> gpurun_out/ab_k1sort.txt
  done
done
done
cat gpurun_out/ab_k1sort.txt

# K3 operating-point sweep on the all-intra clip (c2): warps per CTA and CTAs per frame
for cfg in "8 0" "4 0" "4 32" "4 48" "8 32" "8 16"; do
  set -- $cfg
  AV1R_K3_CTAS=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload c2_intra_1080p8 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('warps $1 ctas $2: value %.0f resident %.0f e2e %.0f intra_ms %.1f'%(d['value'],d['value_hbm_resident'],d['e2e']['value'],d['roofline']['stages']['intra']['ms_per_step']))" >> gpurun_out/sweep_c2.txt
done
cat gpurun_out/sweep_c2.txt

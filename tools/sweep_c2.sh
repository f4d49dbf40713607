# K3 operating-point sweep on the all-intra clip (c2): CTAs per frame (0 = default rule) and hand-over mode
for cfg in "0 2" "16 2" "20 2" "28 2" "36 2" "0 0" "0 1"; do
  set -- $cfg
  AV1R_K3_CTAS=$1 AV1R_K3_PROGRESSIVE=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-per-config --workload c2_intra_1080p8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas $1 progressive $2: value %.0f intra_ms %.1f'%(d['value'],d['roofline']['stages']['intra']['ms_per_step']))" >> gpurun_out/sweep_c2.txt
done
cat gpurun_out/sweep_c2.txt

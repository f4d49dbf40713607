#!/bin/bash
# sweep the K3 ticket window (persistent CTAs per frame): device throughput with 16 streams vs serial per-frame latency
for g in 37 74 148 296 592 1184 2368; do
  AV1R_K3_CTAS=$g timeout 200 python bench.py --steps 3 --warmup 3 --workload ${1:-c2_intra_1080p8} --no-cpu-baseline > gpurun_out/sweep_$g.json 2> gpurun_out/sweep_$g.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$g.json"))
    print("ctas", $g, "value", round(d["value"]), "intra ms/clip", round(d["roofline"]["stages"]["intra"]["ms_per_clip"],1))
except Exception as e:
    print("ctas", $g, "failed", e)
PY
done

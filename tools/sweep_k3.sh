# K3 tuning sweep (GPU box): persistent CTAs per frame (x warps per CTA), device-path fps of a clip
for w in ${WARPS:-8}; do for c in ${CTAS:-8 10 12 14 17 24}; do
  echo "== warps $w ctas $c"; AV1R_K3_WARPS=$w AV1R_K3_CTAS=$c timeout 120 python tools/k3_prof.py ${1:-c2} 2>&1 | grep -v "^\[av1r"
done; done

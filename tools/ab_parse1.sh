# on-box A/B of the host parser, single thread, parse only (no kernels): av1-go_b200/lib/libav1r_base.so against the current build
cp av1-go_b200/lib/libav1r.so /tmp/cur.so
for round in 1 2; do
for v in base cur; do
  if [ $v = cur ]; then cp /tmp/cur.so av1-go_b200/lib/libav1r.so; else cp av1-go_b200/lib/libav1r_$v.so av1-go_b200/lib/libav1r.so; fi
  for c in c2 c3 c1; do
    echo -n "$v $c 1-thread: " >> gpurun_out/ab_parse1.txt; python -m tools.parse_bench $c --reps 3 --no-tiles --threads 1 2>/dev/null | head -1 | cut -c1-110 >> gpurun_out/ab_parse1.txt
  done
done
done
cp /tmp/cur.so av1-go_b200/lib/libav1r.so
cat gpurun_out/ab_parse1.txt

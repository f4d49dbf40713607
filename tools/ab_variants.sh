# on-box comparison of several builds of the library (av1-go_b200/lib/libav1r_<tag>.so): c2 device path at 16 streams, twice each
cp av1-go_b200/lib/libav1r.so /tmp/cur.so
for round in 1 2; do
for v in "$@" cur; do
  if [ $v = cur ]; then cp /tmp/cur.so av1-go_b200/lib/libav1r.so; else cp av1-go_b200/lib/libav1r_$v.so av1-go_b200/lib/libav1r.so; fi
  echo -n "$v: " >> gpurun_out/ab_variants.txt
  python tools/stream_sweep.py c2 16 2>/dev/null >> gpurun_out/ab_variants.txt
  python tools/stream_sweep.py c3 16 2>/dev/null >> gpurun_out/ab_variants.txt
  python tools/stream_sweep.py c1 16 2>/dev/null >> gpurun_out/ab_variants.txt
done
done
cp /tmp/cur.so av1-go_b200/lib/libav1r.so
cat gpurun_out/ab_variants.txt

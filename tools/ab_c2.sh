# A/B of two builds of the library on the same box: c2 (all key frames) operating-point sweep with the base build, then the current one
cp av1-go_b200/lib/libav1r.so /tmp/cur.so
for v in base cur base cur; do
  if [ $v = base ]; then cp av1-go_b200/lib/libav1r_base.so av1-go_b200/lib/libav1r.so; else cp /tmp/cur.so av1-go_b200/lib/libav1r.so; fi
  echo "== $v" >> gpurun_out/ab_c2.txt
  python tools/stream_sweep.py c2 16 32 2>/dev/null >> gpurun_out/ab_c2.txt
  python tools/stream_sweep.py c3 16 2>/dev/null >> gpurun_out/ab_c2.txt
done
cp /tmp/cur.so av1-go_b200/lib/libav1r.so
cat gpurun_out/ab_c2.txt

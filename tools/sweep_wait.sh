for cfg in "0 800" "100 800" "200 800" "400 800" "200 3200" "400 3200" "800 6400" "0 3200"; do
  set -- $cfg
  echo "== wait_ns $1 poll_ns_max $2" >> gpurun_out/sweep_wait.txt
  AV1R_K3_WAIT_NS=$1 AV1R_K3_POLL_NS=$2 python tools/stream_sweep.py c2 4 16 32 2>/dev/null >> gpurun_out/sweep_wait.txt
done
cat gpurun_out/sweep_wait.txt

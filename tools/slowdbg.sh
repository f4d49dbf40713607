# where does the consumer thread's issue time go in the e2e path?  (AV1R_SLOWDBG prints every get_frame / launch of the CDEF and LR
# sections that takes more than 100 us)
for ht in 0 12 8; do
AV1R_SLOWDBG=1 python - $ht > gpurun_out/slowdbg_$ht.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "av1-go_b200"); sys.path.insert(0, ".")
import av1recon
from tools.make_streams import clip_path
ht = int(sys.argv[1])
blob = open(clip_path("c3"), "rb").read()
dec = av1recon.Decoder(streams=16, frames_in_flight=32, host_threads=ht)
dec.verify_buffer(blob)
sys.stderr.write("=== warm-up done\n"); sys.stderr.flush()
best = None
for _ in range(3):
    t0 = time.perf_counter(); rc, rep, d = dec.verify_buffer(blob); dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print(f"host_threads {ht}: e2e {rep.frames / best:.1f} fps")
PY
echo "== host_threads $ht"; grep "e2e" gpurun_out/slowdbg_$ht.txt; sed -n '/warm-up done/,$p' gpurun_out/slowdbg_$ht.txt | grep -c slow; sed -n '/warm-up done/,$p' gpurun_out/slowdbg_$ht.txt | grep slow | sort | uniq -c | sort -rn | head -12
done

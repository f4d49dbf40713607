# Host-side (no kernels timed) profile of the verify path on the GPU box's idle cores: parse-only wall time with tile threads on/off,
# then the e2e call with the engine's and the parser's phase counters (AV1R_PROFILE=1).  usage: bash tools/host_prof.sh TAG clip...
TAG=$1; shift
for c in "$@"; do
  echo "== $c parse-only, tile threads on" >> gpurun_out/hostprof_$TAG.log
  AV1R_PROFILE=1 python -m tools.parse_bench $c --reps 3 >> gpurun_out/hostprof_$TAG.log 2>&1
  echo "== $c parse-only, tile threads off" >> gpurun_out/hostprof_$TAG.log
  AV1R_PROFILE=1 python -m tools.parse_bench $c --reps 3 --no-tiles >> gpurun_out/hostprof_$TAG.log 2>&1
  echo "== $c parse-only, 1 thread" >> gpurun_out/hostprof_$TAG.log
  AV1R_PROFILE=1 python -m tools.parse_bench $c --reps 2 --no-tiles --threads 1 >> gpurun_out/hostprof_$TAG.log 2>&1
  echo "== $c e2e verify" >> gpurun_out/hostprof_$TAG.log
  AV1R_PROFILE=1 python - $c >> gpurun_out/hostprof_$TAG.log 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "av1-go_b200"); sys.path.insert(0, ".")
import av1recon
from tools.make_streams import clip_path
p = sys.argv[1]
blob = open(clip_path(p) if not os.path.exists(p) else p, "rb").read()
dec = av1recon.Decoder(streams=16, frames_in_flight=32)
dec.verify_buffer(blob)
import ctypes as C
l = av1recon.lib()
out6 = (C.c_double * 20)(); out5 = (C.c_double * 5)()
l.av1r_debug_engine_prof(out6, 1); l.av1r_debug_parse_prof(out5, 1)
best = None
for _ in range(3):
    t0 = time.perf_counter(); rc, rep, d = dec.verify_buffer(blob); dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
l.av1r_debug_engine_prof(out6, 0); l.av1r_debug_parse_prof(out5, 0)
print(f"e2e {rep.frames / best:.1f} fps ({best * 1e3:.1f} ms), summed host parse {rep.host_parse_ms:.1f} ms, device_ms sum {rep.device_ms:.1f}")
print("engine prof (3 runs, ms): acquire %.1f prepare %.1f fill %.1f issue %.1f wait_parse %.1f drain %.1f" % tuple(out6[:6]))
print("  workers (3 runs, summed over segment threads, ms): wait_budget %.1f parse_tu %.1f" % (out6[18], out6[19]))
print("  issue detail (3 runs, ms): getframe %.1f itx %.1f inter %.1f intra %.1f deblock %.1f cdef %.1f lr+sr %.1f emit %.1f | arena_ensure %.1f h2d_enqueue %.1f" % tuple(list(out6[8:16]) + [out6[16], out6[17]]))
print("parse prof (3 runs, ms): tiles %.1f merge %.1f lf %.1f wrap %.1f begin %.1f" % tuple(out5))
PY
done

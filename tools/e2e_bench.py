"""End-to-end verify timing of one cached clip: python -m tools.e2e_bench c3 [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))
import av1recon  # noqa: E402
from tools.make_streams import clip_path  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
blob = open(clip_path(name), "rb").read()
dec = av1recon.Decoder(streams=16, frames_in_flight=32)
dec.verify_buffer(blob)
import ctypes as C
l = av1recon.lib()
ep, pp = (C.c_double * 6)(), (C.c_double * 5)()
l.av1r_debug_engine_prof(ep, 1)
l.av1r_debug_parse_prof(pp, 1)
for _ in range(reps):
    t0 = time.perf_counter()
    rc, rep, digs = dec.verify_buffer(blob)
    dt = time.perf_counter() - t0
    l.av1r_debug_engine_prof(ep, 1)
    l.av1r_debug_parse_prof(pp, 1)
    print("   consumer: acquire %.0f prepare %.0f fill %.0f issue %.0f wait_parse %.0f drain %.0f | parser: tiles %.0f merge %.0f lf %.0f wrap %.0f begin %.0f (ms, summed over threads)" % (tuple(ep) + tuple(pp)))
    print(f"{name}: rc {rc} {rep.frames} frames in {dt * 1e3:.1f} ms = {rep.frames / dt:.1f} fps; parse {rep.host_parse_ms:.0f} ms device {rep.device_ms:.0f} ms; {rep.message.decode()}")
dec.close()

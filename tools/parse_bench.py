"""Host parse throughput (no GPU): python -m tools.parse_bench c3 [--threads N] [--no-tiles]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))
import av1recon  # noqa: E402
from tools.make_streams import clip_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("clip")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--no-tiles", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
l = av1recon.lib()
l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
data = open(clip_path(a.clip) if not os.path.exists(a.clip) else a.clip, "rb").read()
best = None
for _ in range(a.reps):
    rep = av1recon.Report()
    rc = l.av1r_parse_buffer(data, len(data), a.threads, 0 if a.no_tiles else 1, C.byref(rep))
    assert rc == 0, rep.message
    if best is None or rep.wall_ms < best.wall_ms:
        best = rep
print(f"{a.clip}: {best.frames} frames, wall {best.wall_ms:.1f} ms = {best.frames_per_sec:.1f} fps, summed parse {best.host_parse_ms:.1f} ms ({best.host_parse_ms / best.frames:.2f} ms/frame); {best.message.decode()}")

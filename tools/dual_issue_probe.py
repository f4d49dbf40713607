"""Is the clip replay bound by the GPU or by the single host thread that issues it?  Replays one clip from K engines (contexts of the
same device) driven by K Python threads at once (ctypes releases the GIL inside av1r_clip_decode) and compares the aggregate rate.
usage: python tools/dual_issue_probe.py c2 [threads...]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import av1recon  # noqa: E402
from tools.make_streams import get_clip  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
tus = get_clip(name)
for K in [int(x) for x in sys.argv[2:]] or [1, 2, 4]:
    decs = [av1recon.Decoder(streams=32 // K if K <= 4 else 8, frames_in_flight=64 // K if K <= 4 else 16) for _ in range(K)]
    clips = [av1recon.Clip(d, tus) for d in decs]
    for c in clips:
        c.decode(); c.decode()
    passes = 8
    def run(c):
        for _ in range(passes):
            c.decode()
    th = [threading.Thread(target=run, args=(c,)) for c in clips]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    nfr = int(clips[0].info.frames_shown)
    print(f"{name}: {K} issuing thread(s): {K * passes * nfr / dt:.0f} frames/s aggregate (wall clock)", flush=True)
    for c in clips: c.free()
    for d in decs: d.close()

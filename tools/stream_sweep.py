"""Clip rate of one replayed clip against the number of CUDA streams / frames in flight (how many frames' kernels may overlap).
usage: python tools/stream_sweep.py c2 [S ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import av1recon  # noqa: E402
from tools.make_streams import get_clip  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
tus = get_clip(name)
for S in [int(x) for x in sys.argv[2:]] or [2, 4, 8, 12, 16, 24, 32]:
    dec = av1recon.Decoder(streams=S, frames_in_flight=2 * S)
    clip = av1recon.Clip(dec, tus)
    clip.decode(); clip.decode()
    ms = 0.0
    for _ in range(5):
        m, _c = clip.decode()
        ms += m
    print(f"{name}: streams {S:2d} frames_in_flight {2 * S:2d}: {5 * int(clip.info.frames_shown) / (ms / 1e3):.0f} frames/s", flush=True)
    clip.free()
    dec.close()

"""Which stage bounds the multi-stream device path?  python tools/conc_test2.py [clip]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'av1-go_b200'))
import av1recon
from tools.make_streams import get_clip
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
tus = get_clip(name)
for filt in (7, 1, 0):
    dec = av1recon.Decoder(streams=16, frames_in_flight=32, inloop_filters=filt)
    clip = av1recon.Clip(dec, tus)
    clip.decode(); clip.decode()
    ms = min(clip.decode()[0] for _ in range(3))
    print(name, 'inloop_filters', filt, 'ms/clip %.1f' % ms, 'fps %.0f' % (len(tus) / ms * 1e3), flush=True)
    clip.free(); dec.close()

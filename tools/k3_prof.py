"""K3 phase profile on the GPU box: AV1R_K3_PROF=1 python tools/k3_prof.py [clip] (cycle totals printed by av1r_close)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'av1-go_b200'))
import av1recon
from tools.make_streams import get_clip
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
tus = get_clip(name)
for streams, fif in ((1, 2), (16, 32)):
    dec = av1recon.Decoder(streams=streams, frames_in_flight=fif)
    clip = av1recon.Clip(dec, tus)
    clip.decode(); clip.decode()
    ms = min(clip.decode()[0] for _ in range(3))
    prof = clip.profile()
    print(name, 'streams', streams, 'fif', fif, 'ms/clip %.1f' % ms, 'fps %.0f' % (len(tus) / ms * 1e3), flush=True)
    clip.free(); dec.close()

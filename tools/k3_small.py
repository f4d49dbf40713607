"""Small-clip decode through the clip API (debug aid): python tools/k3_small.py [clip] [streams]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'av1-go_b200'))
import av1recon
from tools.make_streams import get_clip
name = sys.argv[1] if len(sys.argv) > 1 else 'c2_small'
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tus = get_clip(name)
print('clip', name, len(tus), 'TUs', flush=True)
dec = av1recon.Decoder(streams=streams, frames_in_flight=2 * streams)
clip = av1recon.Clip(dec, tus)
print('decode ms', clip.decode()[0], flush=True)
clip.free(); dec.close()
print('done', flush=True)

import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/av1-go_b200')
import av1recon
from tools.make_streams import get_clip
tus=get_clip('c2')
for streams,fif in ((1,2),(2,4),(4,8),(8,16),(16,32),(32,64)):
    dec=av1recon.Decoder(streams=streams, frames_in_flight=fif)
    clip=av1recon.Clip(dec,tus)
    clip.decode(); clip.decode()
    ms=min(clip.decode()[0] for _ in range(3))
    print('streams',streams,'fif',fif,'ms/clip %.1f'%ms,'fps %.0f'%(60/ms*1e3), flush=True)
    clip.free(); dec.close()

"""Broader oracle-vs-dav1d sweep (all tools on, loop filters + grain on)."""
from tools import dbg_inter as D

D.BASE_OFF.clear()
cases = [
    ("sb128-8b-640x360", dict(w=640, h=360, n=20, lag=19, kf=30, bpc=8, src="panzoom", cpu="5", filters=7, grain=1), {"sb-size": "128"}),
    ("tiles-10b-704x400", dict(w=704, h=400, n=16, lag=19, bpc=10, src="panzoom", cpu="4", filters=7, grain=1), {"tile-columns": "1", "tile-rows": "1"}),
    ("grain-10b-352x288", dict(w=352, h=288, n=10, lag=8, bpc=10, src="noise", cpu="6", filters=7, grain=1, cq="20"), {"film-grain-test": "3"}),
    ("lowq-8b-416x240", dict(w=416, h=240, n=16, lag=16, bpc=8, src="panzoom", cpu="3", filters=7, grain=1, cq="55", seed=9), {}),
    ("hiq-8b-416x240", dict(w=416, h=240, n=12, lag=0, bpc=8, src="panzoom", cpu="6", filters=7, grain=1, cq="8", seed=11), {}),
    ("odd-8b-250x170", dict(w=250, h=170, n=12, lag=10, bpc=8, src="panzoom", cpu="2", filters=7, grain=1, seed=12), {}),
]
for name, kw, o in cases:
    D.run(name, o, **kw)

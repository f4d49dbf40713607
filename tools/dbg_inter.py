"""Debug helper: encode a small inter clip with a chosen libaom tool set, decode with the CPU oracle and dav1d, compare."""
import sys
import time

import numpy as np

from oracle import dav1d_ref, oracle_lib
from tools import aomenc, sources

BASE_OFF = {"enable-obmc": "0", "enable-warped-motion": "0", "enable-global-motion": "0", "enable-ref-frame-mvs": "0",
            "enable-masked-comp": "0", "enable-interintra-comp": "0", "enable-dist-wtd-comp": "0", "enable-diff-wtd-comp": "0",
            "enable-onesided-comp": "0", "enable-dual-filter": "0", "enable-interinter-wedge": "0", "enable-interintra-wedge": "0",
            "enable-smooth-interintra": "0", "enable-restoration": "0", "enable-cdef": "0", "enable-palette": "0", "enable-intrabc": "0"}


def run(name, opts, w=192, h=128, n=6, bpc=8, src="panzoom", lag=0, kf=9999, filters=0, cpu="6", cq="32", seed=5, grain=0):
    o = dict(BASE_OFF)
    o.update({"cpu-used": cpu, "cq-level": cq})
    o.update(opts)
    frames = list(sources.SOURCES[src](w, h, n, bpc=bpc, seed=seed))
    tus = aomenc.encode(frames, w, h, bpc=bpc, opts=o, cfg={14: lag, 48: kf}, threads=1)
    ref = dav1d_ref.decode(tus, inloop_filters=filters, apply_grain=grain)
    t0 = time.time()
    try:
        got, info = oracle_lib.decode_stream(tus, inloop_filters=filters, apply_grain=grain)
    except RuntimeError as e:
        print(f"[{name}] FAIL decode: {e}")
        return False, tus
    ok = len(ref) == len(got)
    first = None
    for i in range(min(len(ref), len(got))):
        for p in range(3):
            a, b = ref[i][4][p].astype(np.int32), got[i][p].astype(np.int32)
            bad = np.argwhere(a != b)
            if len(bad):
                ok = False
                if first is None:
                    y, x = bad[0]
                    first = f"frame {i} plane {p}: {len(bad)} px differ; first (y={y},x={x}) ref={a[y, x]} got={b[y, x]}"
    print("   tools:", {k: v for k, v in oracle_lib.LAST_TOOL_HIST.items() if v})
    print(f"[{name}] {'OK' if ok else 'MISMATCH'} frames={len(got)}/{len(ref)} bytes={sum(len(t) for t in tus)} {first or ''} ({time.time()-t0:.1f}s)")
    return ok, tus


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "base"
    if which == "base":
        run("base", {})
    CASES = {
        "big": ({}, dict(w=352, h=288, n=8)),
        "lag": ({}, dict(w=352, h=288, n=12, lag=8)),
        "refmvs": ({"enable-ref-frame-mvs": "1"}, dict(w=352, h=288, n=12, lag=8)),
        "dual": ({"enable-dual-filter": "1"}, dict(w=352, h=288, n=8, lag=4)),
        "onesided": ({"enable-onesided-comp": "1"}, dict(w=352, h=288, n=12, lag=8)),
        "distwtd": ({"enable-dist-wtd-comp": "1"}, dict(w=352, h=288, n=12, lag=8)),
        "masked": ({"enable-masked-comp": "1", "enable-interinter-wedge": "1", "enable-diff-wtd-comp": "1"}, dict(w=352, h=288, n=12, lag=8, cpu="3")),
        "interintra": ({"enable-interintra-comp": "1", "enable-interintra-wedge": "1", "enable-smooth-interintra": "1"}, dict(w=352, h=288, n=10, lag=4, cpu="2")),
        "obmc": ({"enable-obmc": "1"}, dict(w=352, h=288, n=10, lag=4, cpu="3")),
        "warp": ({"enable-warped-motion": "1"}, dict(w=352, h=288, n=10, lag=4, cpu="3")),
        "global": ({"enable-global-motion": "1"}, dict(w=352, h=288, n=10, lag=4, cpu="3")),
        "filters": ({"enable-cdef": "1", "enable-restoration": "1"}, dict(w=352, h=288, n=10, lag=4, filters=7)),
        "tenbit": ({}, dict(w=352, h=288, n=8, lag=4, bpc=10)),
    }
    if which in CASES:
        o, kw = CASES[which]
        run(which, o, **kw)
    elif which == "all":
        for k, (o, kw) in CASES.items():
            run(k, o, **kw)
    if which == "full":
        import itertools
        BASE_OFF.clear()
        for cpu, src, bpc, sz in [("2", "panzoom", 8, (352, 288)), ("1", "testsrc2", 8, (320, 192)), ("4", "noise", 10, (352, 288)), ("0", "panzoom", 10, (208, 144))]:
            run(f"full-cpu{cpu}-{src}-{bpc}", {}, w=sz[0], h=sz[1], n=12, lag=10, bpc=bpc, src=src, cpu=cpu, filters=7, grain=1)
    if which == "iso":
        BASE_OFF.clear()
        for name, o in [("no-ii", {"enable-interintra-comp": "0"}), ("no-dist", {"enable-dist-wtd-comp": "0"}), ("no-dual", {"enable-dual-filter": "0"}),
                        ("no-ii-dist", {"enable-interintra-comp": "0", "enable-dist-wtd-comp": "0"})]:
            run(name, o, w=208, h=144, n=4, lag=10, bpc=10, src="panzoom", cpu="0", filters=0, grain=0)
    if which == "pal":
        BASE_OFF.clear()
        run("pal-intra", {"tune-content": "screen"}, w=320, h=192, n=3, lag=0, kf=0, src="testsrc2", cpu="4", filters=7)
        run("pal-inter", {"tune-content": "screen"}, w=320, h=192, n=6, lag=0, src="testsrc2", cpu="4", filters=7)
        run("pal-auto", {}, w=320, h=192, n=6, lag=4, src="testsrc2", cpu="1", filters=7)
        run("pal-10b", {"tune-content": "screen", "enable-intrabc": "0"}, w=320, h=192, n=4, lag=0, src="testsrc2", cpu="4", filters=7, bpc=10)

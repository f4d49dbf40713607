"""Deterministic synthetic AV1 clips for the BASELINE configs.  Clips are cached under
streams_cache/ (git-ignored, shipped to the GPU box by gpurun); regenerate with
    python -m tools.make_streams c2 [--frames N]
No ffmpeg / lavfi exists in the image, so `testsrc2` is re-synthesised in numpy ("testsrc2-like")."""
import argparse
import os
import sys

from . import aomenc, obuio, sources

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "streams_cache")

# name: (source, w, h, bpc, frames, opts, cfg)   cfg[14] = lag_in_frames, cfg[48] = kf_max_dist
CONFIGS = {
    # configs[1]: 1080p 8-bit intra-only (every frame KEY), inverse transform + intra + deblock/CDEF only
    "c2": ("panzoom", 1920, 1080, 8, 60, {"cpu-used": "8", "cq-level": "32", "enable-restoration": "0", "enable-cdef": "1"}, {14: 0, 48: 0}),
    # configs[0]: 1080p 8-bit Main-profile clip, testsrc2-like, 60 frames, 2 closed GOPs, 2 tile columns, all default tools
    "c1": ("testsrc2", 1920, 1080, 8, 60, {"cpu-used": "8", "cq-level": "32", "tile-columns": "1"}, {14: 19, 48: 30}),
    # configs[2]: 4K 10-bit, full inter tools (compound, OBMC, warped / global motion) + loop restoration, 4x2 tiles
    # (round 2: `occluders` source + cpu-used 1 -- at cpu-used 6 libaom never picked OBMC / masked compound / inter-intra and switched
    # loop restoration off at this size; the two closed GOPs are encoded as two independent parts, see get_clip)
    "c3": ("occluders", 3840, 2160, 10, 60, {"cpu-used": "1", "cq-level": "32", "tile-columns": "2", "tile-rows": "1", "enable-obmc": "1",
                                            "enable-warped-motion": "1", "enable-global-motion": "1", "enable-restoration": "1"}, {14: 19, 48: 30}, 3, 2),
    # configs[3]: as c3 on a noise-heavy source with film grain synthesis
    "c4": ("noise", 3840, 2160, 10, 60, {"cpu-used": "6", "cq-level": "32", "tile-columns": "2", "tile-rows": "1", "enable-restoration": "1",
                                        "film-grain-test": "5"}, {14: 19, 48: 30}),
    "c3_small": ("panzoom", 960, 544, 10, 20, {"cpu-used": "6", "cq-level": "32", "tile-columns": "1", "tile-rows": "1", "enable-restoration": "1"},
                 {14: 19, 48: 10}),
    # configs[4]: 32 files as c3 (seeds 100..131), 16 frames each in two closed GOPs of 8 => 64 independent GOP segments
    **{f"c5_{i:02d}": ("occluders", 3840, 2160, 10, 16, {"cpu-used": "2", "cq-level": "32", "tile-columns": "2", "tile-rows": "1", "enable-obmc": "1",
                                                            "enable-warped-motion": "1", "enable-global-motion": "1", "enable-restoration": "1"},
                       {14: 7, 48: 8}, 100 + i) for i in range(32)},
    # screen content at 1080p, every frame a key frame with palette + intra block copy (not a BASELINE config: the full-size parity
    # case for the decode-order unit table of block-copy frames, 30 x 17 units and ~60 CTAs in flight per frame)
    "screen1080": ("screen", 1920, 1080, 8, 6, {"cpu-used": "2", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0",
                                             "enable-cdef": "0"}, {14: 0, 48: 0}),
    "c2_small": ("panzoom", 640, 360, 8, 8, {"cpu-used": "8", "cq-level": "32", "enable-restoration": "0", "enable-cdef": "1"}, {14: 0, 48: 0}),
}


def clip_path(name, frames=None):
    src, w, h, bpc, n, opts, cfg = CONFIGS[name][:7]
    n = frames or n
    return os.path.join(CACHE, f"{name}_{w}x{h}_{bpc}b_{n}f.ivf")


def _encode_part(args):
    name, n, a, b, threads = args
    src, w, h, bpc, _, opts, cfg = CONFIGS[name][:7]
    seed = CONFIGS[name][7] if len(CONFIGS[name]) > 7 else 3
    fr = (f for i, f in enumerate(sources.SOURCES[src](w, h, n, bpc=bpc, seed=seed)) if a <= i < b)
    return aomenc.encode(fr, w, h, bpc=bpc, opts=opts, cfg=cfg, threads=threads)


def get_clip(name, frames=None, verbose=False, threads=None):
    """Returns the list of temporal units of a config clip, generating (and caching) it if needed.  A config with `parts` > 1 is
    encoded as that many independent encodes of consecutive frame ranges run side by side (each range = one closed GOP starting
    with its own key frame and sequence header) and concatenated: same stream structure as one encode with kf_max_dist = range
    length, at half the wall time on this 8-core build box."""
    path = clip_path(name, frames)
    if os.path.exists(path):
        return obuio.read_ivf(path)
    src, w, h, bpc, n, opts, cfg = CONFIGS[name][:7]
    parts = CONFIGS[name][8] if len(CONFIGS[name]) > 8 else 1
    n = frames or n
    threads = threads or os.cpu_count() or 8
    if verbose:
        print(f"encoding {name}: {w}x{h} {bpc}-bit {n} frames with libaom ({parts} part(s)) ...", file=sys.stderr)
    if parts > 1:
        import multiprocessing as mp
        step = (n + parts - 1) // parts
        jobs = [(name, n, a, min(n, a + step), max(1, threads // parts)) for a in range(0, n, step)]
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            tus = [tu for part in pool.map(_encode_part, jobs) for tu in part]
    else:
        tus = _encode_part((name, n, 0, n, threads))
    os.makedirs(CACHE, exist_ok=True)
    obuio.write_ivf(path + ".tmp", tus, w, h)
    os.replace(path + ".tmp", path)
    return tus


def _c5_one(i):
    t = get_clip(f"c5_{i:02d}", verbose=True, threads=max(1, (os.cpu_count() or 8) // 2))
    return len(t), sum(len(x) for x in t)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--frames", type=int, default=None)
    a = ap.parse_args()
    if a.name == "c5":   # two files side by side, 4 encoder threads each
        import multiprocessing as mp
        with mp.get_context("fork").Pool(2) as pool:
            for i, t in enumerate(pool.imap(_c5_one, range(32))):
                print(clip_path(f"c5_{i:02d}"), t[0], "TUs", t[1], "bytes", flush=True)
        sys.exit(0)
    t = get_clip(a.name, a.frames, verbose=True)
    print(clip_path(a.name, a.frames), len(t), "TUs", sum(len(x) for x in t), "bytes")

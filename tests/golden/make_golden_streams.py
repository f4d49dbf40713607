"""Generates the small golden AV1 streams under tests/golden/streams/ plus their per-frame per-plane
MD5 (libdav1d 1.5.3, cross-checked against the libaom decoder).  MD5 domain = visible samples, rows
tightly packed, 8-bit as bytes, >8-bit as little-endian uint16 (what `ffmpeg -f framemd5` hashes).
    python tests/golden/make_golden_streams.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dav1d_ref  # noqa: E402
from tools import aomenc, obuio, sources  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "streams")

# name: (source, w, h, bpc, frames, opts, cfg)
CASES = {
    "intra_8b_200x136": ("panzoom", 200, 136, 8, 3, {"cpu-used": "5", "cq-level": "30", "enable-restoration": "0"}, {14: 0, 48: 0}),
    "intra_10b_192x128": ("panzoom", 192, 128, 10, 2, {"cpu-used": "4", "cq-level": "24", "enable-restoration": "0"}, {14: 0, 48: 0}),
    "intra_8b_tiles_320x192": ("noise", 320, 192, 8, 2, {"cpu-used": "6", "cq-level": "12", "enable-restoration": "0", "tile-columns": "1", "tile-rows": "1"}, {14: 0, 48: 0}),
    "intra_8b_sb128_264x200": ("panzoom", 264, 200, 8, 2, {"cpu-used": "3", "cq-level": "45", "enable-restoration": "0", "sb-size": "128"}, {14: 0, 48: 0}),
    "intra_8b_lr_480x272": ("panzoom", 480, 272, 8, 2, {"cpu-used": "1", "cq-level": "40", "enable-restoration": "1"}, {14: 0, 48: 0}),
    "intra_8b_lr_tiles_616x376": ("panzoom", 616, 376, 8, 2, {"cpu-used": "3", "cq-level": "50", "enable-restoration": "1", "tile-columns": "1"}, {14: 0, 48: 0}),
    "intra_8b_grain_160x96": ("noise", 160, 96, 8, 2, {"cpu-used": "8", "cq-level": "30", "enable-restoration": "0", "film-grain-test": "7"}, {14: 0, 48: 0}),
    # super-resolution (K6): cfg[19] = rc_superres_mode (1 = fixed), cfg[20] / cfg[21] = denominator for inter / key frames (9..16, 8 = off)
    "intra_8b_superres_lr_328x200": ("panzoom", 328, 200, 8, 3, {"cpu-used": "4", "cq-level": "30", "enable-restoration": "1"}, {14: 0, 48: 0, 19: 1, 20: 12, 21: 12}),
    "intra_10b_superres16_264x136": ("panzoom", 264, 136, 10, 2, {"cpu-used": "5", "cq-level": "40", "enable-restoration": "0"}, {14: 0, 48: 0, 19: 1, 20: 16, 21: 16}),
    # quantiser matrices (K1): per-frame qm levels between qm-min and qm-max
    # monochrome (cfg[52] = monochrome): luma only, one MD5 per frame
    "intra_8b_mono_264x200": ("panzoom", 264, 200, 8, 2, {"cpu-used": "4", "cq-level": "30"}, {14: 0, 48: 0, 52: 1}),
    # 10-bit key frames with loop restoration, CDEF, 128x128 superblocks and a non-zero deblocking sharpness
    "intra_10b_lr_sb128_sharp3_520x296": ("panzoom", 520, 296, 10, 2, {"cpu-used": "2", "cq-level": "44", "enable-restoration": "1", "sb-size": "128", "sharpness": "3"}, {14: 0, 48: 0}),
    # key frames with segmentation + delta_q + delta_lf; 10-bit key frames with 2 x 2 tiles, loop restoration and super-resolution 8/13
    "intra_8b_aq_deltaq_264x200": ("panzoom", 264, 200, 8, 2, {"cpu-used": "3", "cq-level": "36", "aq-mode": "1", "deltaq-mode": "1", "delta-lf-mode": "1"}, {14: 0, 48: 0}),
    "intra_10b_tiles_lr_superres_520x296": ("panzoom", 520, 296, 10, 2, {"cpu-used": "3", "cq-level": "40", "enable-restoration": "1", "tile-columns": "1", "tile-rows": "1"},
                                            {14: 0, 48: 0, 19: 1, 20: 13, 21: 13}),
    # 10-bit screen content with palette, CDEF and loop restoration on
    "intra_10b_screen_cdef_320x192": ("screen", 320, 192, 10, 2, {"cpu-used": "3", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "0",
                                      "enable-restoration": "1"}, {14: 0, 48: 0}),
    "intra_8b_qm_264x200": ("panzoom", 264, 200, 8, 2, {"cpu-used": "4", "cq-level": "30", "enable-qm": "1", "qm-min": "2", "qm-max": "10"}, {14: 0, 48: 0}),
    # screen content (tune-content=screen on a source of flat colours and recurring glyphs): palette mode, and intra block copy (K3:
    # the predictor is the frame being decoded displaced by a block vector; libaom only picks it with CDEF off or at low cpu-used).
    # The 322x182 case also has chroma-from-luma blocks whose luma transform block straddles the coded frame edge.
    "intra_8b_palette_320x192": ("screen", 320, 192, 8, 2, {"cpu-used": "3", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "0", "enable-restoration": "0"}, {14: 0, 48: 0}),
    "intra_8b_intrabc_320x192": ("screen", 320, 192, 8, 2, {"cpu-used": "2", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0", "enable-cdef": "0"}, {14: 0, 48: 0}),
    "intra_10b_intrabc_328x200": ("screen", 328, 200, 10, 2, {"cpu-used": "1", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0", "enable-cdef": "0"}, {14: 0, 48: 0}),
    "intra_8b_intrabc_sb128_456x264": ("screen", 456, 264, 8, 2, {"cpu-used": "2", "cq-level": "20", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0", "enable-cdef": "0", "sb-size": "128"}, {14: 0, 48: 0}),
    "intra_8b_intrabc_edge_322x182": ("screen", 322, 182, 8, 2, {"cpu-used": "2", "cq-level": "50", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0", "enable-cdef": "0"}, {14: 0, 48: 0}),
}

# inter streams (index_inter.json): cfg[14] = lag_in_frames, cfg[48] = kf_max_dist.  The low cpu-used cases make libaom use
# every inter tool (see the tool histogram printed by tools/dbg_inter.py): compound average / distance / wedge / diff-weighted,
# inter-intra (smooth + wedge), OBMC, local + global warp, skip mode, dual filters, temporal MVs, sub-8x8 chroma, var-tx.
INTER_CASES = {
    "inter_8b_base_192x128": ("panzoom", 192, 128, 8, 8, {"cpu-used": "6", "cq-level": "32", "enable-restoration": "0", "enable-cdef": "0",
                                                         "enable-obmc": "0", "enable-warped-motion": "0", "enable-global-motion": "0"}, {14: 0, 48: 9999}),
    "inter_8b_alltools_352x288": ("panzoom", 352, 288, 8, 12, {"cpu-used": "2", "cq-level": "32"}, {14: 10, 48: 9999}),
    "inter_10b_alltools_208x144": ("panzoom", 208, 144, 10, 12, {"cpu-used": "0", "cq-level": "32"}, {14: 10, 48: 9999}),
    "inter_8b_sb128_tiles_640x360": ("panzoom", 640, 360, 8, 10, {"cpu-used": "5", "cq-level": "40", "sb-size": "128", "tile-columns": "1", "tile-rows": "1"},
                                     {14: 9, 48: 6}),
    "inter_10b_grain_208x144": ("noise", 208, 144, 10, 8, {"cpu-used": "4", "cq-level": "40", "film-grain-test": "3"}, {14: 6, 48: 9999}),
    # key frame coded at 8/11 of the width and upscaled; the inter frames (full width) predict from the upscaled reference
    "inter_8b_kfsuperres_320x192": ("panzoom", 320, 192, 8, 6, {"cpu-used": "5", "cq-level": "32"}, {14: 0, 48: 9999, 19: 1, 20: 8, 21: 11}),
    # scaled references (spec 7.11.3.3): fixed spatial resize with different denominators for key and inter frames (cfg[16..18] =
    # rc_resize_mode, rc_resize_denominator, rc_resize_kf_denominator), and super-resolution on inter frames (references have the
    # upscaled width, the frame is predicted at the coded width: horizontal scaling only)
    "inter_8b_refscale_352x288": ("panzoom", 352, 288, 8, 10, {"cpu-used": "3", "cq-level": "34"}, {14: 4, 48: 9999, 16: 1, 17: 12, 18: 10}),
    "inter_10b_refscale_208x144": ("panzoom", 208, 144, 10, 8, {"cpu-used": "2", "cq-level": "30"}, {14: 4, 48: 9999, 16: 1, 17: 14, 18: 9}),
    "inter_8b_superres_inter_352x288": ("panzoom", 352, 288, 8, 10, {"cpu-used": "3", "cq-level": "34"}, {14: 4, 48: 9999, 19: 1, 20: 12, 21: 9}),
    # screen content with inter frames: block-copy + palette key frame, then integer motion vectors (force_integer_mv) and palette
    # blocks inside inter frames
    "inter_8b_screen_352x288": ("screen", 352, 288, 8, 6, {"cpu-used": "2", "cq-level": "30", "tune-content": "screen", "enable-intrabc": "1", "enable-restoration": "0", "enable-cdef": "0"}, {14: 4, 48: 9999}),
    "inter_10b_mono_208x144": ("panzoom", 208, 144, 10, 6, {"cpu-used": "3", "cq-level": "36"}, {14: 4, 48: 9999, 52: 1}),
    # syntax paths off libaom's defaults (each stream differs from the all-tools ones in one header switch):
    # lossless (base_q_idx 0: Walsh-Hadamard transform only, no in-loop filters)
    "inter_8b_lossless_128x96": ("panzoom", 128, 96, 8, 3, {"cpu-used": "4", "lossless": "1"}, {14: 0, 48: 9999}),
    # delta_q + delta_lf per superblock (per-block quantiser index and per-block deblocking levels)
    "inter_8b_deltaq_lf_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "3", "cq-level": "36", "deltaq-mode": "1", "delta-lf-mode": "1"}, {14: 4, 48: 9999}),
    # segmentation: variance AQ (spatially coded map, per-segment quantiser) and cyclic refresh (temporally predicted map)
    "inter_8b_aq1_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "4", "cq-level": "36", "aq-mode": "1"}, {14: 4, 48: 9999}),
    "inter_8b_aq3_256x160": ("panzoom", 256, 160, 8, 8, {"cpu-used": "5", "cq-level": "36", "aq-mode": "3"}, {14: 0, 48: 9999}),
    # enable_order_hint = 0 (no temporal motion vectors, no skip mode, no distance-weighted compound)
    "inter_8b_nohint_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "3", "cq-level": "36", "enable-order-hint": "0"}, {14: 4, 48: 9999}),
    # disable_cdf_update = 1 / disable_frame_end_update_cdf = 1 with two tile columns / error_resilient_mode = 1 (cfg[12])
    "inter_8b_nocdfupd_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "4", "cq-level": "36", "cdf-update-mode": "0"}, {14: 4, 48: 9999}),
    "inter_8b_frameparallel_352x288": ("panzoom", 352, 288, 8, 6, {"cpu-used": "4", "cq-level": "36", "frame-parallel": "1", "tile-columns": "1"}, {14: 4, 48: 9999}),
    "inter_8b_errres_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "4", "cq-level": "36"}, {14: 4, 48: 9999, 12: 1}),
    # reduced_tx_set = 1
    "inter_8b_reducedtx_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "3", "cq-level": "36", "reduced-tx-type-set": "1"}, {14: 4, 48: 9999}),
    # deblocking with loop_filter_sharpness != 0
    "inter_8b_sharp5_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "4", "cq-level": "40", "sharpness": "5"}, {14: 4, 48: 9999}),
    # every frame another size: dynamic (random) spatial resize = references both larger and smaller than the frame, odd sizes
    # (cfg[16] = rc_resize_mode 2), and a random super-resolution denominator per frame (cfg[19] = rc_superres_mode 2)
    "inter_8b_resize_dyn_352x288": ("panzoom", 352, 288, 8, 10, {"cpu-used": "4", "cq-level": "36"}, {14: 0, 48: 9999, 16: 2}),
    "inter_8b_superres_rand_352x288": ("panzoom", 352, 288, 8, 10, {"cpu-used": "4", "cq-level": "36"}, {14: 0, 48: 9999, 19: 2}),
    # film grain: monochrome 10-bit (lag 2), lag 3 with overlap, and the 8-point scaling function of libaom's test vector 16
    "inter_10b_grain_mono_208x144": ("noise", 208, 144, 10, 6, {"cpu-used": "4", "cq-level": "40", "film-grain-test": "9"}, {14: 4, 48: 9999, 52: 1}),
    "inter_8b_grain11_208x144": ("noise", 208, 144, 8, 6, {"cpu-used": "4", "cq-level": "40", "film-grain-test": "11"}, {14: 4, 48: 9999}),
    "inter_8b_grain16_208x144": ("noise", 208, 144, 8, 6, {"cpu-used": "4", "cq-level": "40", "film-grain-test": "16"}, {14: 4, 48: 9999}),
    # separate chroma quantiser deltas; a frame header OBU followed by several tile group OBUs (2 x 2 tiles in 3 groups) instead of one
    # OBU_FRAME; most sequence-level tool switches off at once (the symbols those tools would read are absent from the bitstream)
    "inter_8b_chromadq_256x160": ("panzoom", 256, 160, 8, 6, {"cpu-used": "4", "cq-level": "36", "enable-chroma-deltaq": "1"}, {14: 4, 48: 9999}),
    "inter_8b_tilegroups_352x288": ("panzoom", 352, 288, 8, 6, {"cpu-used": "4", "cq-level": "36", "tile-columns": "1", "tile-rows": "1", "num-tile-groups": "3"}, {14: 4, 48: 9999}),
    "inter_8b_seqflags_off_256x160": ("panzoom", 256, 160, 8, 8, {"cpu-used": "2", "cq-level": "36", "enable-dual-filter": "0", "enable-dist-wtd-comp": "0",
                                      "enable-ref-frame-mvs": "0", "enable-masked-comp": "0", "enable-interintra-comp": "0", "enable-intra-edge-filter": "0",
                                      "enable-filter-intra": "0", "enable-cdef": "0", "enable-restoration": "0", "enable-warped-motion": "0",
                                      "enable-cfl-intra": "0", "enable-palette": "0"}, {14: 6, 48: 9999}),
    # a hidden SWITCH_FRAME (frame_type 3: error-resilient, refreshes all eight slots, explicit reference order hints) between inter
    # frames (cfg[49] = sframe_dist 2, cfg[50] = sframe_mode 1, error resilient, lag 8)
    "inter_8b_sframe_256x160": ("panzoom", 256, 160, 8, 12, {"cpu-used": "5", "cq-level": "36"}, {14: 8, 48: 9999, 49: 2, 50: 1, 12: 1}),
    # 10-bit versions of the rarer paths: lossless; complexity AQ + delta_q + delta_lf with 128x128 superblocks; a new frame size on
    # every frame
    "inter_10b_lossless_128x96": ("panzoom", 128, 96, 10, 3, {"cpu-used": "4", "lossless": "1"}, {14: 0, 48: 9999}),
    "inter_10b_aq2_deltaq_sb128_384x224": ("panzoom", 384, 224, 10, 6, {"cpu-used": "3", "cq-level": "36", "aq-mode": "2", "deltaq-mode": "1",
                                           "delta-lf-mode": "1", "sb-size": "128"}, {14: 4, 48: 9999}),
    "inter_10b_resize_dyn_304x208": ("panzoom", 304, 208, 10, 8, {"cpu-used": "4", "cq-level": "36"}, {14: 0, 48: 9999, 16: 2}),
    # extreme geometry: one 16 x 16 block per frame; a frame much taller than wide (one superblock column)
    "inter_8b_smallest_16x16": ("panzoom", 16, 16, 8, 4, {"cpu-used": "4", "cq-level": "30"}, {14: 0, 48: 9999}),
    "inter_8b_tall_72x520": ("panzoom", 72, 520, 8, 4, {"cpu-used": "4", "cq-level": "34"}, {14: 0, 48: 9999}),
    # odd frame size (410 x 230: the last mi column / row is half outside the picture) with 4 x 2 tiles at 10 bits
    "inter_10b_tiles4x2_odd_410x230": ("panzoom", 410, 230, 10, 6, {"cpu-used": "4", "cq-level": "36", "tile-columns": "2", "tile-rows": "1"}, {14: 4, 48: 9999}),
    "inter_10b_qm_208x144": ("panzoom", 208, 144, 10, 6, {"cpu-used": "3", "cq-level": "36", "enable-qm": "1", "qm-min": "0", "qm-max": "15"}, {14: 4, 48: 9999}),
}


def main():
    """all | inter | intra | only NAME...  (`only` adds / refreshes the named cases and keeps the rest of the index)"""
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    only = set(sys.argv[2:]) if which == "only" else None
    if which in ("all", "inter", "only"):
        build(INTER_CASES, "index_inter.json", only)
    if which in ("all", "intra", "only"):
        build(CASES, "index.json", only)


def build(cases, index_name, only=None):
    index = {}
    if only is not None:
        index = json.load(open(os.path.join(OUT, index_name)))
    for name, (src, w, h, bpc, n, opts, cfg) in cases.items():
        if only is not None and name not in only:
            continue
        frames = list(sources.SOURCES[src](w, h, n, bpc=bpc, seed=7 if "lr" in name else len(name)))
        tus = aomenc.encode(frames, w, h, bpc=bpc, opts=opts, cfg=cfg, threads=1)
        ref = dav1d_ref.decode(tus)
        aom = aomenc.decode(tus)
        assert len(ref) == len(aom) == n
        md5 = []
        for i in range(n):
            for p in range(len(ref[i][4])):   # one plane for monochrome
                assert np.array_equal(ref[i][4][p], aom[i][p]), "dav1d and libaom disagree"
            md5.append(dav1d_ref.plane_md5(ref[i][4]))
        obuio.write_ivf(os.path.join(OUT, name + ".ivf"), tus, w, h)
        index[name] = dict(w=w, h=h, bpc=bpc, frames=n, md5=md5, bytes=sum(len(t) for t in tus))
        print(name, index[name]["bytes"], "bytes")
    json.dump(index, open(os.path.join(OUT, index_name), "w"), indent=1)


if __name__ == "__main__":
    main()

"""Synthetic AV1 bitstream generator: libaom 3.13.1 encoder driven through ctypes.

The library is the one bundled with opencv-python-headless in this image (there is no
aomenc / ffmpeg binary and no network).  ABI facts: SURVEY.md Appendix B.
This is input preparation for tests and the bench, not part of the product path.
"""
import ctypes as C
import glob
import os

import numpy as np

_LIB = None

AOM_IMG_FMT_I420 = 0x102
AOM_IMG_FMT_I42016 = 0x902
AOM_CODEC_USE_HIGHBITDEPTH = 0x40000
ENC_ABI = 25
DEC_ABI = 22


def lib_path():
    import cv2  # noqa: F401
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    hits = sorted(glob.glob(os.path.join(base, "libaom-*.so*")))
    if not hits:
        raise RuntimeError("libaom not found")
    return hits[0]


def lib():
    global _LIB
    if _LIB is None:
        # locate without importing cv2 (slow) when possible
        import importlib.util
        spec = importlib.util.find_spec("cv2")
        base = os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs")
        hits = sorted(glob.glob(os.path.join(base, "libaom-*.so*")))
        l = C.CDLL(hits[0])
        l.aom_codec_av1_cx.restype = C.c_void_p
        l.aom_codec_av1_dx.restype = C.c_void_p
        l.aom_codec_enc_config_default.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        l.aom_codec_enc_init_ver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        l.aom_codec_dec_init_ver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        l.aom_codec_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        l.aom_codec_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_ulong, C.c_long]
        l.aom_codec_get_cx_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.aom_codec_get_cx_data.restype = C.c_void_p
        l.aom_codec_destroy.argtypes = [C.c_void_p]
        l.aom_img_alloc.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_uint]
        l.aom_img_alloc.restype = C.c_void_p
        l.aom_img_free.argtypes = [C.c_void_p]
        l.aom_codec_error.argtypes = [C.c_void_p]
        l.aom_codec_error.restype = C.c_char_p
        l.aom_codec_error_detail.argtypes = [C.c_void_p]
        l.aom_codec_error_detail.restype = C.c_char_p
        l.aom_codec_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        l.aom_codec_get_frame.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.aom_codec_get_frame.restype = C.c_void_p
        _LIB = l
    return _LIB


def _img_fill(l, img, planes, bpc):
    raw = (C.c_uint8 * 104).from_address(img)
    hdr = np.frombuffer(raw, dtype=np.uint8)
    ptrs = hdr[64:88].view(np.uint64)
    strides = hdr[88:100].view(np.int32)
    bps = 1 if bpc == 8 else 2
    for i, p in enumerate(planes):
        h, w = p.shape
        st = int(strides[i])
        buf = (C.c_uint8 * (st * h)).from_address(int(ptrs[i]))
        dst = np.frombuffer(buf, dtype=np.uint8).reshape(h, st)
        src = np.ascontiguousarray(p.astype(np.uint8 if bpc == 8 else "<u2")).view(np.uint8).reshape(h, w * bps)
        dst[:, : w * bps] = src


def encode(frames, w, h, bpc=8, opts=None, cfg=None, threads=8, usage=0):
    """frames: iterable of [Y,U,V] numpy planes (4:2:0).  opts: dict of string options for
    aom_codec_set_option.  cfg: dict {index: value} poked into aom_codec_enc_cfg_t (uint32 view).
    Returns list of temporal units (bytes)."""
    l = lib()
    iface = l.aom_codec_av1_cx()
    cfgbuf = (C.c_uint32 * 1024)()
    rc = l.aom_codec_enc_config_default(iface, cfgbuf, usage)
    if rc:
        raise RuntimeError(f"config_default {rc}")
    cfgbuf[1] = threads
    cfgbuf[2] = 0  # profile 0
    cfgbuf[3] = w
    cfgbuf[4] = h
    cfgbuf[8] = bpc
    cfgbuf[9] = bpc
    cfgbuf[24] = 3  # AOM_Q
    for k, v in (cfg or {}).items():
        cfgbuf[k] = v
    ctx = (C.c_uint8 * 256)()
    flags = AOM_CODEC_USE_HIGHBITDEPTH if bpc > 8 else 0
    rc = l.aom_codec_enc_init_ver(ctx, iface, cfgbuf, flags, ENC_ABI)
    if rc:
        raise RuntimeError(f"enc_init {rc}: {l.aom_codec_error(ctx)}")
    o = {"cpu-used": "8", "cq-level": "32", "row-mt": "1"}
    o.update(opts or {})
    for k, v in o.items():
        rc = l.aom_codec_set_option(ctx, k.encode(), str(v).encode())
        if rc:
            raise RuntimeError(f"set_option {k}={v} -> {rc}: {l.aom_codec_error_detail(ctx)}")
    img = l.aom_img_alloc(None, AOM_IMG_FMT_I420 if bpc == 8 else AOM_IMG_FMT_I42016, w, h, 32)
    tus = []

    def drain():
        it = C.c_void_p(None)
        while True:
            pkt = l.aom_codec_get_cx_data(ctx, C.byref(it))
            if not pkt:
                break
            kind = C.c_int.from_address(pkt).value
            if kind != 0:
                continue
            buf = C.c_void_p.from_address(pkt + 8).value
            sz = C.c_size_t.from_address(pkt + 16).value
            tus.append(C.string_at(buf, sz))

    pts = 0
    for planes in frames:
        _img_fill(l, img, planes, bpc)
        rc = l.aom_codec_encode(ctx, img, pts, 1, 0)
        if rc:
            raise RuntimeError(f"encode {rc}: {l.aom_codec_error(ctx)} {l.aom_codec_error_detail(ctx)}")
        drain()
        pts += 1
    while True:
        n = len(tus)
        rc = l.aom_codec_encode(ctx, None, pts, 1, 0)
        if rc:
            raise RuntimeError(f"flush {rc}")
        drain()
        if len(tus) == n:
            break
    l.aom_img_free(img)
    l.aom_codec_destroy(ctx)
    return tus


def decode(tus):
    """libaom decoder cross-check: returns list of [Y,U,V] planes per output frame."""
    l = lib()
    ctx = (C.c_uint8 * 256)()
    rc = l.aom_codec_dec_init_ver(ctx, l.aom_codec_av1_dx(), None, 0, DEC_ABI)
    if rc:
        raise RuntimeError(f"dec_init {rc}")
    out = []
    for tu in tus:
        rc = l.aom_codec_decode(ctx, tu, len(tu), None)
        if rc:
            raise RuntimeError(f"aom decode {rc}: {l.aom_codec_error(ctx)}")
        it = C.c_void_p(None)
        while True:
            img = l.aom_codec_get_frame(ctx, C.byref(it))
            if not img:
                break
            raw = (C.c_uint8 * 104).from_address(img)
            hdr = np.frombuffer(raw, dtype=np.uint8)
            fmt = int(hdr[0:4].view(np.uint32)[0])
            dw, dh = int(hdr[40:44].view(np.uint32)[0]), int(hdr[44:48].view(np.uint32)[0])
            xs, ys = int(hdr[56:60].view(np.uint32)[0]), int(hdr[60:64].view(np.uint32)[0])
            ptrs = hdr[64:88].view(np.uint64)
            strides = hdr[88:100].view(np.int32)
            hb = bool(fmt & 0x800)
            bps = 2 if hb else 1
            planes = []
            for i in range(3):
                pw = dw if i == 0 else (dw + xs) >> xs
                ph = dh if i == 0 else (dh + ys) >> ys
                st = int(strides[i])
                buf = (C.c_uint8 * (st * ph)).from_address(int(ptrs[i]))
                a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, st)[:, : pw * bps]
                planes.append(np.ascontiguousarray(a).view(np.dtype("<u2") if hb else np.uint8).reshape(ph, pw).copy())
            out.append(planes)
    l.aom_codec_destroy(ctx)
    return out

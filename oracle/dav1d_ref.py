"""TEST INFRASTRUCTURE ONLY -- the reference decoder, driven directly.

The reference daemon (IONIQ6000/av1-go) has no decode code of its own: every
pixel would be produced by libdav1d inside the FFmpeg build it downloads
(/root/reference/internal/config/config.go:33, internal/ffmpeg/binary.go:104-211).
That library (dav1d 1.5.3, API 7.0) is exported by Pillow's bundled libavif in
this image, so this module drives it through ctypes and is the parity ORACLE
(kind "reference") for every stage of the path.  It must only be imported from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

Struct layouts are hand-declared (no headers on disk); see SURVEY.md Appendix B.
"""
import ctypes as C
import glob
import hashlib
import os
import sys

import numpy as np

_LIB = None


def lib_path():
    import PIL  # noqa: F401  (only to locate site-packages/pillow.libs)
    base = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    hits = sorted(glob.glob(os.path.join(base, "libavif-*.so*")))
    if not hits:
        raise RuntimeError("dav1d oracle unavailable: no libavif in pillow.libs")
    return hits[0]


class Dav1dSettings(C.Structure):
    _fields_ = [
        ("n_threads", C.c_int), ("max_frame_delay", C.c_int), ("apply_grain", C.c_int),
        ("operating_point", C.c_int), ("all_layers", C.c_int), ("frame_size_limit", C.c_uint),
        ("alloc_cookie", C.c_void_p), ("alloc_cb", C.c_void_p), ("release_cb", C.c_void_p),
        ("log_cookie", C.c_void_p), ("log_cb", C.c_void_p),
        ("strict_std_compliance", C.c_int), ("output_invisible_frames", C.c_int),
        ("inloop_filters", C.c_int), ("decode_frame_type", C.c_int),
        ("reserved", C.c_uint8 * 16),
    ]


class Dav1dUserData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("ref", C.c_void_p)]


class Dav1dDataProps(C.Structure):
    _fields_ = [("timestamp", C.c_int64), ("duration", C.c_int64), ("offset", C.c_int64),
                ("size", C.c_size_t), ("user_data", Dav1dUserData)]


class Dav1dData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("sz", C.c_size_t), ("ref", C.c_void_p),
                ("m", Dav1dDataProps)]


class Dav1dPicParams(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("layout", C.c_int), ("bpc", C.c_int)]


class Dav1dPicture(C.Structure):
    _fields_ = [
        ("seq_hdr", C.c_void_p), ("frame_hdr", C.c_void_p),
        ("data", C.c_void_p * 3), ("stride", C.c_ssize_t * 2),
        ("p", Dav1dPicParams), ("m", Dav1dDataProps),
        ("content_light", C.c_void_p), ("mastering_display", C.c_void_p),
        ("itut_t35", C.c_void_p), ("n_itut_t35", C.c_size_t),
        ("reserved", C.c_size_t * 4),
        ("frame_hdr_ref", C.c_void_p), ("seq_hdr_ref", C.c_void_p),
        ("content_light_ref", C.c_void_p), ("mastering_display_ref", C.c_void_p),
        ("itut_t35_ref", C.c_void_p), ("reserved_ref", C.c_size_t * 4),
        ("ref", C.c_void_p), ("allocator_data", C.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        l = C.CDLL(lib_path())
        l.dav1d_version.restype = C.c_char_p
        l.dav1d_default_settings.argtypes = [C.POINTER(Dav1dSettings)]
        l.dav1d_open.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Dav1dSettings)]
        l.dav1d_close.argtypes = [C.POINTER(C.c_void_p)]
        l.dav1d_data_create.argtypes = [C.POINTER(Dav1dData), C.c_size_t]
        l.dav1d_data_create.restype = C.c_void_p
        l.dav1d_send_data.argtypes = [C.c_void_p, C.POINTER(Dav1dData)]
        l.dav1d_get_picture.argtypes = [C.c_void_p, C.POINTER(Dav1dPicture)]
        l.dav1d_picture_unref.argtypes = [C.POINTER(Dav1dPicture)]
        l.dav1d_data_unref.argtypes = [C.POINTER(Dav1dData)]
        _LIB = l
    return _LIB


def version():
    return lib().dav1d_version().decode()


def _grab(pic, keep):
    w, h, bpc, layout = pic.p.w, pic.p.h, pic.p.bpc, pic.p.layout
    if not keep:
        return (w, h, bpc, layout, None)
    bps = 1 if bpc == 8 else 2
    dt = np.uint8 if bpc == 8 else np.dtype("<u2")
    ssx = 1 if layout in (1, 2) else 0
    ssy = 1 if layout == 1 else 0
    planes = []
    for i in range(3 if layout != 0 else 1):
        pw = w if i == 0 else (w + ssx) >> ssx
        ph = h if i == 0 else (h + ssy) >> ssy
        st = pic.stride[0 if i == 0 else 1]
        buf = (C.c_uint8 * (st * ph)).from_address(pic.data[i])
        a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, st)[:, : pw * bps]
        planes.append(np.ascontiguousarray(a).view(dt).reshape(ph, pw).copy())
    return (w, h, bpc, layout, planes)


def decode(tus, n_threads=1, apply_grain=1, inloop_filters=7, keep=True, max_frame_delay=0):
    """Decode a list of temporal units (bytes).  Returns a list of
    (w, h, bpc, layout, [Y,U,V] numpy planes or None) in display order."""
    l = lib()
    s = Dav1dSettings()
    l.dav1d_default_settings(C.byref(s))
    s.n_threads = n_threads
    s.max_frame_delay = max_frame_delay
    s.apply_grain = apply_grain
    s.inloop_filters = inloop_filters
    ctx = C.c_void_p()
    rc = l.dav1d_open(C.byref(ctx), C.byref(s))
    if rc:
        raise RuntimeError(f"dav1d_open {rc}")
    out = []
    pic = Dav1dPicture()

    def pull():
        got = False
        while True:
            C.memset(C.byref(pic), 0, C.sizeof(pic))
            r = l.dav1d_get_picture(ctx, C.byref(pic))
            if r != 0:
                return got
            out.append(_grab(pic, keep))
            l.dav1d_picture_unref(C.byref(pic))
            got = True

    try:
        for tu in tus:
            d = Dav1dData()
            p = l.dav1d_data_create(C.byref(d), len(tu))
            C.memmove(p, tu, len(tu))
            while d.sz > 0:
                r = l.dav1d_send_data(ctx, C.byref(d))
                if r not in (0, -11):
                    l.dav1d_data_unref(C.byref(d))
                    raise RuntimeError(f"dav1d_send_data {r}")
                pull()
        # drain
        while pull():
            pass
    finally:
        l.dav1d_close(C.byref(ctx))
    return out


def decode_md5(tus, n_threads=1, apply_grain=1, inloop_filters=7):
    """Per-frame per-plane MD5 (framemd5 domain: visible rows tightly packed, >8-bit as LE uint16) of libdav1d's output, computed
    picture by picture so that BASELINE-size clips need not be held in memory.  -> list of [md5(Y), md5(U), md5(V)] hex strings."""
    l = lib()
    s = Dav1dSettings()
    l.dav1d_default_settings(C.byref(s))
    s.n_threads = n_threads
    s.max_frame_delay = 0
    s.apply_grain = apply_grain
    s.inloop_filters = inloop_filters
    ctx = C.c_void_p()
    rc = l.dav1d_open(C.byref(ctx), C.byref(s))
    if rc:
        raise RuntimeError(f"dav1d_open {rc}")
    out = []
    pic = Dav1dPicture()

    def pull():
        got = False
        while True:
            C.memset(C.byref(pic), 0, C.sizeof(pic))
            if l.dav1d_get_picture(ctx, C.byref(pic)) != 0:
                return got
            out.append(plane_md5(_grab(pic, True)[4]))
            l.dav1d_picture_unref(C.byref(pic))
            got = True

    try:
        for tu in tus:
            d = Dav1dData()
            p = l.dav1d_data_create(C.byref(d), len(tu))
            C.memmove(p, tu, len(tu))
            while d.sz > 0:
                r = l.dav1d_send_data(ctx, C.byref(d))
                if r not in (0, -11):
                    l.dav1d_data_unref(C.byref(d))
                    raise RuntimeError(f"dav1d_send_data {r}")
                pull()
        while pull():
            pass
    finally:
        l.dav1d_close(C.byref(ctx))
    return out


def plane_md5(planes):
    """md5 over tightly packed visible rows; >8-bit as LE uint16 (framemd5 domain)."""
    return [hashlib.md5(np.ascontiguousarray(p).tobytes()).hexdigest() for p in planes]


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.obuio import read_ivf
    tus = read_ivf(sys.argv[1])
    print("dav1d", version(), len(tus), "TUs")
    for i, (w, h, bpc, layout, pl) in enumerate(decode(tus, n_threads=0)):
        print(i, w, h, bpc, *plane_md5(pl))

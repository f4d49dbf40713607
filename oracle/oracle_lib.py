"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_build/liboracle.so (the plain-C/C++
restatement of the reconstruction stages).  Import only from tests/, smoke() and bench.py's
cpu_baseline leg."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("oracle library not built: run `make`")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def film_grain(fg_params, planes, bpc, subx=1, suby=1, mono=0, mc_identity=0):
    """fg_params: any ctypes struct with the av1r_film_grain_params layout.  planes: [Y,U,V] numpy."""
    l = lib()
    h, w = planes[0].shape
    dt = np.uint8 if bpc == 8 else np.uint16
    src = [np.ascontiguousarray(p.astype(dt)) for p in planes]
    dst = [np.zeros_like(p) for p in src]
    sp = (C.c_void_p * 3)(*[p.ctypes.data for p in src])
    dp = (C.c_void_p * 3)(*[p.ctypes.data for p in dst])
    ss = (C.c_int * 3)(*[p.strides[0] for p in src])
    ds = (C.c_int * 3)(*[p.strides[0] for p in dst])
    l.orc_film_grain(C.byref(fg_params), bpc, w, h, subx, suby, mono, mc_identity, sp, ss, dp, ds)
    return dst


TOOL_NAMES = ["inter_blocks", "compound_avg", "compound_dist", "compound_wedge", "compound_diffwtd", "interintra", "interintra_wedge",
              "obmc", "local_warp", "global_warp", "skip_mode", "dual_filter", "temporal_mv", "intra_in_inter", "sub8x8_chroma", "newmv",
              "vartx_split", "switchable_filter", "palette", "intrabc"]
LAST_TOOL_HIST = {}


def decode_stream(tus, inloop_filters=7, apply_grain=1):
    """Whole-stream CPU decode (product host parser + scalar oracle reconstruction).
    Returns (frames, info) with frames = list of [Y,U,V] uint16 arrays."""
    l = lib()
    l.orc_stream_open.restype = C.c_void_p
    l.orc_stream_error.restype = C.c_char_p
    l.orc_stream_error.argtypes = [C.c_void_p]
    l.orc_stream_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    l.orc_stream_close.argtypes = [C.c_void_p]
    l.orc_stream_num_frames.argtypes = [C.c_void_p]
    l.orc_stream_frame_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    l.orc_stream_frame_copy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    h = l.orc_stream_open(inloop_filters, apply_grain)
    try:
        for i, tu in enumerate(tus):
            rc = l.orc_stream_decode(h, tu, len(tu))
            if rc < 0:
                raise RuntimeError(f"oracle decode failed at TU {i}: rc={rc} {l.orc_stream_error(h).decode()}")
        n = l.orc_stream_num_frames(h)
        hist = (C.c_uint64 * 24)()
        l.orc_stream_tool_hist.argtypes = [C.c_void_p, C.c_void_p]
        l.orc_stream_tool_hist(h, hist)
        global LAST_TOOL_HIST
        LAST_TOOL_HIST = dict(zip(TOOL_NAMES, list(hist)))
        frames, info = [], []
        for i in range(n):
            w, hh, bd = C.c_int(), C.c_int(), C.c_int()
            pm = C.c_double()
            st = (C.c_uint64 * 3)()
            l.orc_stream_frame_info(h, i, C.byref(w), C.byref(hh), C.byref(bd), C.byref(pm), st)
            planes = []
            for p in range(3):
                pw = w.value if p == 0 else (w.value + 1) // 2
                ph = hh.value if p == 0 else (hh.value + 1) // 2
                a = np.zeros((ph, pw), dtype=np.uint16)
                l.orc_stream_frame_copy(h, i, p, a.ctypes.data, pw)
                planes.append(a)
            frames.append(planes)
            info.append(dict(w=w.value, h=hh.value, bd=bd.value, parse_ms=pm.value, coded_samples=st[0], coef_tokens=st[1], tx_blocks=st[2]))
        return frames, info
    finally:
        l.orc_stream_close(h)

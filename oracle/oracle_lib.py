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

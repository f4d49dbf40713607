"""The intra kernel (K3) hands 64x64 units to persistent CTAs in a host-built order and lets records wait only for records / units
that come earlier in it; av1r_debug_k3_check rebuilds that plan for every frame of a stream on the host and checks the invariants
(csrc/k3_plan.h: ranges tile the order, records lie in their unit in decode order, dependencies are lower-index neighbours, every
neighbour a record can read is a dependency, inter-intra residuals follow their blend record)."""
import ctypes as C
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STREAMS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "streams", "*.ivf")))


@pytest.mark.parametrize("path", STREAMS, ids=[os.path.basename(p)[:-4] for p in STREAMS])
def test_k3_plan_invariants(built, path):
    import av1recon
    l = av1recon.lib()
    l.av1r_debug_k3_check.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_char_p, C.c_size_t]
    data = open(path, "rb").read()
    frames, units = C.c_longlong(0), C.c_longlong(0)
    msg = C.create_string_buffer(512)
    rc = l.av1r_debug_k3_check(data, len(data), C.byref(frames), C.byref(units), msg, 512)
    assert rc == 0, msg.value.decode()
    assert frames.value > 0 and units.value > 0

"""Debug helper: decode a stream with the CPU oracle and with dav1d, report the first mismatch."""
import sys

import numpy as np

from oracle import dav1d_ref, oracle_lib


def compare(tus, inloop_filters=0, apply_grain=0, verbose=True):
    ref = dav1d_ref.decode(tus, inloop_filters=inloop_filters, apply_grain=apply_grain)
    got, info = oracle_lib.decode_stream(tus, inloop_filters=inloop_filters, apply_grain=apply_grain)
    ok = len(ref) == len(got)
    if not ok and verbose:
        print("frame count", len(ref), len(got))
    for i in range(min(len(ref), len(got))):
        for p in range(3):
            a, b = ref[i][4][p].astype(np.int32), got[i][p].astype(np.int32)
            if a.shape != b.shape:
                print("shape", i, p, a.shape, b.shape)
                ok = False
                continue
            bad = np.argwhere(a != b)
            if len(bad):
                ok = False
                if verbose:
                    y, x = bad[0]
                    # first mismatching 4x4 in raster-by-64x64 order helps locate the block
                    sb = sorted(set((int(yy) // (64 >> (p > 0)), int(xx) // (64 >> (p > 0))) for yy, xx in bad[:5000]))[:3]
                    print(f"frame {i} plane {p}: {len(bad)} px differ; first (y={y},x={x}) ref={a[y, x]} got={b[y, x]}; SBs {sb}")
    return ok

"""Generates tests/golden/filmgrain.npz: inputs (grain-free dav1d planes + the stream's temporal
units) and outputs (dav1d planes with apply_grain=1) for every libaom film-grain test vector,
so that the K8 parity test does not depend on the encoder being present.  Run from repo root:
    python tests/golden/make_golden_filmgrain.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dav1d_ref  # noqa: E402
from tools import aomenc, sources  # noqa: E402

W, H = 136, 72


def main():
    out = {}
    for bpc in (8, 10):
        for tv in range(1, 17):
            fr = list(sources.noise_gradient(W, H, 1, bpc=bpc, seed=tv))
            tus = aomenc.encode(fr, W, H, bpc=bpc, opts={"film-grain-test": str(tv), "cpu-used": "9"}, cfg={14: 0, 48: 30}, threads=1)
            d0 = dav1d_ref.decode(tus, apply_grain=0)
            d1 = dav1d_ref.decode(tus, apply_grain=1)
            key = f"b{bpc}_tv{tv}"
            out[key + "_tus"] = np.frombuffer(b"".join(tus), dtype=np.uint8)
            out[key + "_tulens"] = np.array([len(t) for t in tus], dtype=np.int64)
            for i in range(len(d0)):
                for p in range(3):
                    out[f"{key}_f{i}_in{p}"] = d0[i][4][p]
                    out[f"{key}_f{i}_out{p}"] = d1[i][4][p]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "filmgrain.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()

"""A verify step sees whatever the transcoder left on disk: truncated files, flipped bits, zeroed ranges.  The host parser must
come back (0 or an AV1R_E* code) on every one of them -- never crash -- and the CUDA path must neither hang nor fault on the
work-lists such streams produce (the intra kernel carries watchdogs for exactly this)."""
import ctypes as C
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
NAMES = ["intra_8b_200x136", "inter_8b_alltools_352x288", "intra_8b_lr_480x272", "inter_10b_grain_208x144",
         "intra_8b_superres_lr_328x200", "inter_8b_sb128_tiles_640x360"]


def _mutants(seed, n):
    rng = random.Random(seed)
    for _ in range(n):
        name = rng.choice(NAMES)
        data = bytearray(open(os.path.join(GOLD, name + ".ivf"), "rb").read())
        kind = rng.choice(["flip", "flip", "flip", "trunc", "zero"])
        if kind == "flip":
            for _ in range(rng.randint(1, 4)):
                data[rng.randrange(32, len(data))] ^= 1 << rng.randrange(8)
        elif kind == "trunc":
            data = data[:rng.randrange(40, len(data))]
        else:
            p = rng.randrange(44, len(data) - 16)
            data[p:p + 16] = bytes(16)
        yield name, kind, bytes(data)


def test_host_parser_survives_corrupted_streams(built):
    import av1recon
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    codes = {}
    for name, kind, data in _mutants(20261018, 60):
        rep = av1recon.Report()
        rc = l.av1r_parse_buffer(data, len(data), 1, 0, C.byref(rep))
        assert rc <= 0, (name, kind, rc)
        codes[rc] = codes.get(rc, 0) + 1
    assert codes.get(0, 0) > 0 and len(codes) > 1, codes   # some mutants still parse, some are rejected


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_cuda_verify_survives_corrupted_streams(built):
    import av1recon
    dec = av1recon.Decoder(streams=4, frames_in_flight=8, host_threads=2)
    for name, kind, data in _mutants(7, 24):
        rc, rep, digests = dec.verify_buffer(data)
        assert rc <= 0, (name, kind, rc)
    # the engine is still usable afterwards: a clean stream decodes to its golden frame count
    good = open(os.path.join(GOLD, "intra_8b_200x136.ivf"), "rb").read()
    rc, rep, digests = dec.verify_buffer(good)
    assert rc == 0 and rep.frames == 3, rep.message
    dec.close()

"""Matroska demux: the artefact at the daemon's seam is `<base>.av1-tmp.mkv`
(/root/reference/internal/daemon/daemon.go:86, `-f matroska` /root/reference/internal/ffmpeg/transcode.go:143).  The test muxer
(tools/mkvmux.py) wraps golden streams like FFmpeg does; probe / host parse (CPU) and the CUDA verify path (GPU) must see the same
frames as through IVF."""
import ctypes as C
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
INDEX = dict(json.load(open(os.path.join(GOLD, "index.json"))))
INDEX.update(json.load(open(os.path.join(GOLD, "index_inter.json"))))
NAMES = ["intra_8b_200x136", "inter_8b_sb128_tiles_640x360", "intra_10b_superres16_264x136"]


def _mkv(name, **kw):
    from tools import mkvmux
    from tools.obuio import read_ivf
    tus = read_ivf(os.path.join(GOLD, name + ".ivf"))
    return mkvmux.mux(tus, INDEX[name]["w"], INDEX[name]["h"], **kw)


@pytest.mark.parametrize("unknown", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_probe_and_host_parse_matroska(built, name, unknown):
    import av1recon
    data = _mkv(name, frames_per_cluster=3, unknown_size_clusters=unknown)
    info = av1recon.probe_buffer(data)
    meta = INDEX[name]
    assert info.is_av1 == 1 and info.width == meta["w"] and info.height == meta["h"] and info.bit_depth == meta["bpc"]
    assert info.temporal_units == meta["frames"]
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    rep = av1recon.Report()
    rc = l.av1r_parse_buffer(data, len(data), 2, 1, C.byref(rep))
    assert rc == 0 and rep.frames == meta["frames"], rep.message


def test_matroska_without_av1_track_is_rejected(built):
    import av1recon
    data = _mkv("intra_8b_200x136").replace(b"V_AV1", b"V_VP9")
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    rep = av1recon.Report()
    assert l.av1r_parse_buffer(data, len(data), 1, 0, C.byref(rep)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_verify_matroska_equals_ivf(built, name):
    import av1recon
    ivf = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
    rc0, rep0, d0 = av1recon.verify_buffer(ivf)
    rc1, rep1, d1 = av1recon.verify_buffer(_mkv(name, frames_per_cluster=4))
    assert rc0 == 0 and rc1 == 0, (rep0.message, rep1.message)
    assert rep1.frames == rep0.frames == INDEX[name]["frames"]
    assert [list(map(int, x)) for x in d0] == [list(map(int, x)) for x in d1]

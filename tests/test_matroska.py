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


# ---- layouts FFmpeg's muxer really writes (transcode.go:143), beyond the minimal test muxer ---------------------------------
def _parse(data):
    import av1recon
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    rep = av1recon.Report()
    rc = l.av1r_parse_buffer(data, len(data), 2, 0, C.byref(rep))
    return rc, rep


@pytest.mark.parametrize("kw", [dict(ffmpeg_like=True), dict(ffmpeg_like=True, block_groups=True),
                                dict(ffmpeg_like=True, unknown_size_clusters=True), dict(block_groups=True)],
                         ids=["seekhead_void_audio_first_tags_cues", "block_groups", "unknown_size", "plain_block_groups"])
@pytest.mark.parametrize("name", NAMES)
def test_ffmpeg_shaped_matroska(built, name, kw):
    """SeekHead / Void / Tags before the clusters, Cues after them, an audio track *before* the AV1 track (so the AV1 track is
    number 2), interleaved audio blocks with EBML / Xiph / fixed lacing, BlockGroups: the frames found are those of the IVF file."""
    import av1recon
    data = _mkv(name, frames_per_cluster=3, **kw)
    info = av1recon.probe_buffer(data)
    meta = INDEX[name]
    assert info.is_av1 == 1 and (info.width, info.height, info.bit_depth) == (meta["w"], meta["h"], meta["bpc"])
    assert info.temporal_units == meta["frames"]
    rc, rep = _parse(data)
    assert rc == 0 and rep.frames == meta["frames"], rep.message


@pytest.mark.parametrize("lacing", ["xiph", "ebml"])
def test_laced_video_blocks_are_split_into_temporal_units(built, lacing):
    """Lacing is legal EBML on any track: a laced block of the AV1 track yields one temporal unit per laced frame."""
    name = "inter_8b_sb128_tiles_640x360"
    data = _mkv(name, frames_per_cluster=4, video_lacing=lacing)
    rc, rep = _parse(data)
    assert rc == 0 and rep.frames == INDEX[name]["frames"], rep.message


def test_demuxed_payload_is_byte_identical_to_the_ivf_units(built):
    """Independent check of the demuxer (not just 'it parses'): concatenated temporal-unit payloads it reports for the FFmpeg-shaped
    file equal the IVF units with their temporal delimiters removed -- compared through the parse of both (same frame count and
    the same summed host statistics), since offsets are internal."""
    import av1recon
    l = av1recon.lib()
    l.av1r_parse_stats.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(av1recon.ClipInfo)]
    name = "inter_8b_sb128_tiles_640x360"
    ivf = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
    a, b = av1recon.ClipInfo(), av1recon.ClipInfo()
    assert l.av1r_parse_stats(ivf, len(ivf), C.byref(a)) == 0
    mk = _mkv(name, ffmpeg_like=True, block_groups=True, video_lacing="ebml")
    assert l.av1r_parse_stats(mk, len(mk), C.byref(b)) == 0
    for f in ("frames_decoded", "frames_shown", "coded_samples", "coef_tokens", "tx_blocks", "inter_blocks", "intra_samples"):
        assert getattr(a, f) == getattr(b, f), f
    assert list(a.tool_hist) == list(b.tool_hist)

"""bench.py's workload definitions, checked without a GPU: the repeated 4K clips really are the encoded sequence played CLIP_REPEAT
times (same temporal units for both arms), the container built from them parses to that many frames in twice that many closed GOP
segments, and the config object of the JSON line says so."""
import ctypes as C
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    return bench


def test_repeated_clip_is_the_same_stream_for_both_arms(built):
    import av1recon
    from av1recon import shard
    from tools.make_streams import clip_path
    from tools.obuio import ivf_header, read_ivf
    bench = _bench()
    if not os.path.exists(clip_path("c3")):
        pytest.skip("streams_cache/c3 not generated on this box")
    rep = bench.CLIP_REPEAT["c3"]
    assert rep >= 1
    base = read_ivf(clip_path("c3"))
    tus = bench.bench_clip("c3")
    assert len(tus) == rep * len(base) and tus[:len(base)] == base and tus[-len(base):] == base
    # the reference arm decodes what WORKLOADS[...][2]() returns: the same list
    assert bench.WORKLOADS["c3_4k10_inter"][2]() == tus
    cfg = bench.workload_config("c3_4k10_inter", 1)
    assert cfg["frames_per_step"] == 60 * rep
    if rep > 1:
        assert f"{60 * rep} frames in {2 * rep} closed GOPs" in cfg["workload"]
    # configs[0] names 60 frames: not repeated
    assert bench.workload_config("c1_1080p8", 1)["frames_per_step"] == 60 and bench.CLIP_REPEAT.get("c1", 1) == 1
    # the container the e2e call gets: every repetition is its own pair of key-frame-delimited segments
    hdr = ivf_header(clip_path("c3"))
    blob = shard.ivf_bytes(tus, hdr["w"], hdr["h"])
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    out = av1recon.Report()
    assert l.av1r_parse_buffer(blob, len(blob), 0, 1, C.byref(out)) == 0, out.message
    assert out.frames == 60 * rep
    assert f"{2 * rep} GOP segments".encode() in out.message

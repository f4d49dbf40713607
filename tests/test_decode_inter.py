"""Whole-path parity on inter streams (BASELINE configs[0], [2], [3] shapes at test size): hidden ARF frames,
show_existing_frame, compound / OBMC / warped / global motion, inter-intra, temporal MV prediction, loop filters and
film grain -- per-frame per-plane MD5 vs libdav1d.  CPU: oracle (host parser + scalar reconstruction) vs golden MD5
and vs live dav1d per stage.  GPU: the CUDA engine through the C ABI vs the same."""
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
INDEX = json.load(open(os.path.join(GOLD, "index_inter.json")))


def _tus(name):
    from tools.obuio import read_ivf
    return read_ivf(os.path.join(GOLD, name + ".ivf"))


def _md5(planes, bpc):
    return [hashlib.md5(np.ascontiguousarray(p.astype(np.uint8 if bpc == 8 else "<u2")).tobytes()).hexdigest() for p in planes]


@pytest.mark.parametrize("name", sorted(INDEX))
def test_oracle_matches_golden_md5(built, name):
    from oracle import oracle_lib
    meta = INDEX[name]
    frames, _ = oracle_lib.decode_stream(_tus(name))
    assert len(frames) == meta["frames"]
    for i, planes in enumerate(frames):
        assert _md5(planes, meta["bpc"])[:len(meta["md5"][i])] == meta["md5"][i], f"{name} frame {i}"


def test_golden_streams_exercise_the_inter_tools(built):
    """The committed streams really contain the tools the path claims to cover (counted by the host parser)."""
    from oracle import oracle_lib
    total = {}
    for name in INDEX:
        oracle_lib.decode_stream(_tus(name))
        for k, v in oracle_lib.LAST_TOOL_HIST.items():
            total[k] = total.get(k, 0) + v
    for tool in ("inter_blocks", "compound_avg", "compound_dist", "compound_wedge", "compound_diffwtd", "interintra", "interintra_wedge", "obmc",
                 "local_warp", "global_warp", "skip_mode", "dual_filter", "temporal_mv", "intra_in_inter", "sub8x8_chroma", "newmv", "vartx_split"):
        assert total.get(tool, 0) > 0, f"no golden stream uses {tool}: {total}"


def test_syntax_switch_streams_use_their_switch(built):
    """The golden streams named after a header switch really carry it (header scan, host only): lossless, delta_q / delta_lf,
    segmentation with spatial and temporal map coding, enable_order_hint = 0, disable_cdf_update, disable_frame_end_update_cdf with
    several tiles, error_resilient_mode, reduced_tx_set, an odd frame size under 4 x 2 tiles."""
    import av1recon

    def hdrs(name):
        return [h for h in av1recon.scan_headers(_tus(name)) if not h.show_existing_frame]

    assert all(h.coded_lossless and h.base_q_idx == 0 for h in hdrs("inter_8b_lossless_128x96"))
    dq = hdrs("inter_8b_deltaq_lf_256x160")
    assert any(h.delta_q_present and h.delta_lf_present for h in dq)
    aq1 = hdrs("inter_8b_aq1_256x160")
    assert all(h.segmentation_enabled for h in aq1) and any(h.segmentation_temporal_update for h in aq1)
    assert any(h.segmentation_enabled and h.segmentation_update_map for h in hdrs("inter_8b_aq3_256x160"))
    assert all(h.enable_order_hint == 0 for h in hdrs("inter_8b_nohint_256x160"))
    assert all(h.disable_cdf_update for h in hdrs("inter_8b_nocdfupd_256x160"))
    fp = hdrs("inter_8b_frameparallel_352x288")
    assert all(h.disable_frame_end_update_cdf and not h.disable_cdf_update and h.tile_cols == 2 for h in fp)
    er = hdrs("inter_8b_errres_256x160")
    assert all(h.error_resilient_mode and h.primary_ref_frame == 7 for h in er)
    assert all(h.reduced_tx_set for h in hdrs("inter_8b_reducedtx_256x160"))
    assert any(h.frame_type == 3 and not h.show_frame and h.refresh_frame_flags == 255 for h in hdrs("inter_8b_sframe_256x160"))   # SWITCH_FRAME
    odd = hdrs("inter_10b_tiles4x2_odd_410x230")
    assert all(h.tile_cols == 4 and h.tile_rows == 2 and h.width == 410 and h.height == 230 and h.bit_depth == 10 for h in odd)


@pytest.mark.parametrize("filters", [0, 7])
def test_oracle_stage_isolation_vs_dav1d(built, filters):
    from oracle import dav1d_ref, oracle_lib
    for name in ("inter_8b_alltools_352x288", "inter_10b_alltools_208x144"):
        tus = _tus(name)
        ref = dav1d_ref.decode(tus, inloop_filters=filters, apply_grain=0)
        got, _ = oracle_lib.decode_stream(tus, inloop_filters=filters, apply_grain=0)
        assert len(ref) == len(got)
        for i in range(len(ref)):
            for p in range(3):
                assert np.array_equal(ref[i][4][p], got[i][p]), f"{name} filters={filters} frame {i} plane {p}"


def _gpu_decode(name, **kw):
    import av1recon
    dec = av1recon.Decoder(parity_md5=1, keep_frames=1, **kw)
    for i, tu in enumerate(_tus(name)):
        dec.submit(tu, i)
    dec.flush()
    return dec


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(INDEX))
def test_cuda_matches_golden_md5(built, name):
    meta = INDEX[name]
    dec = _gpu_decode(name)
    assert len(dec.results) == meta["frames"], dec.error()
    for i, r in enumerate(dec.results):
        assert r.status == 0 and r.bpc == meta["bpc"]
        if "refscale" not in name and "resize" not in name:   # (spatially resized streams: the coded size differs from the source size the index records)
            assert r.w == meta["w"] and r.h == meta["h"]
        got = [bytes(r.md5[p]).hex() for p in range(len(meta["md5"][i]))]
        if got != meta["md5"][i]:
            from oracle import oracle_lib
            ref, _ = oracle_lib.decode_stream(_tus(name))
            planes = dec.frame_planes(r)
            msg = []
            for p in range(3):
                bad = np.argwhere(planes[p].astype(np.int32) != ref[i][p].astype(np.int32))
                if len(bad):
                    y, x = bad[0]
                    msg.append(f"plane {p}: {len(bad)} px differ, first (y={y}, x={x}) cuda={planes[p][y, x]} ref={ref[i][p][y, x]}")
            pytest.fail(f"{name} frame {i}: MD5 mismatch; " + "; ".join(msg))
    dec.close()


@pytest.mark.gpu
@pytest.mark.parametrize("filters", [0, 7])
def test_cuda_stage_isolation(built, filters):
    """CUDA engine with the in-loop filters masked vs the oracle at the same stage: isolates K2 (+K3 blends) from K4..K7."""
    from oracle import oracle_lib
    for name in sorted(INDEX):
        ref, _ = oracle_lib.decode_stream(_tus(name), inloop_filters=filters, apply_grain=0)
        dec = _gpu_decode(name, inloop_filters=filters, apply_grain=0)
        assert len(dec.results) == len(ref), dec.error()
        for i, r in enumerate(dec.results):
            planes = dec.frame_planes(r)
            for p in range(3):
                bad = np.argwhere(planes[p].astype(np.int32) != ref[i][p].astype(np.int32))
                assert len(bad) == 0, (f"{name} filters={filters} frame {i} plane {p}: {len(bad)} px differ, first (y,x)={tuple(bad[0])} "
                                       f"cuda={planes[p][tuple(bad[0])]} ref={ref[i][p][tuple(bad[0])]}")
        dec.close()


@pytest.mark.gpu
def test_cuda_verify_buffer_inter(built):
    """Whole-file entry point on an inter clip with two GOP segments (kf_max_dist 6): digests equal the streaming path's."""
    import av1recon
    name = "inter_8b_sb128_tiles_640x360"
    data = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
    rc, rep, digests = av1recon.verify_buffer(data)
    assert rc == 0 and rep.status == 0, rep.message
    assert rep.frames == INDEX[name]["frames"]
    dec = _gpu_decode(name)
    for i, r in enumerate(dec.results):
        assert [int(x) for x in r.checksum] == [int(x) for x in digests[i]], f"frame {i}"
    dec.close()


@pytest.mark.gpu
def test_cuda_clip_replay_passes_back_to_back(built):
    """av1r_clip_decode_passes (what bench.py times): several replays of a clip enqueued back to back without a drain in between --
    slots, frame buffers and reference events recycled across the pass boundary -- reproduce the digests of the single-pass replay
    and of the whole-file verify path; more in-flight frames than slots on purpose."""
    import av1recon
    name = "inter_8b_alltools_352x288"
    tus = _tus(name)
    data = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
    rc, rep, digests = av1recon.verify_buffer(data)
    assert rc == 0 and rep.status == 0, rep.message
    dec = av1recon.Decoder(streams=4, frames_in_flight=8)
    clip = av1recon.Clip(dec, tus * 2)          # the sequence twice: every repetition starts with its key frame
    ms1, one = clip.decode()
    ms5, five = clip.decode_passes(5)
    assert one == five and len(one) == 2 * rep.frames
    assert [tuple(int(x) for x in d) for d in digests] * 2 == [tuple(int(x) for x in d) for d in one]
    assert ms5 > 0
    clip.free()
    dec.close()


@pytest.mark.gpu
@pytest.mark.timeout(120)
def test_cuda_verify_buffer_many_multi_tu_segments(built):
    """A batch container with more multi-TU GOP segments than the parser look-ahead budget (the C5 shape: the workers of later
    segments must not starve the segment the consumer is issuing -- this used to deadlock).  Digests equal the streaming path's."""
    import av1recon
    from av1recon import shard
    from tools.obuio import read_ivf
    name = "inter_8b_sb128_tiles_640x360"
    tus = read_ivf(os.path.join(GOLD, name + ".ivf"))
    copies = 6                                           # 12 segments of 6 / 4 temporal units; budget = max(2 * 4 threads, 8) = 8
    blob = shard.ivf_bytes(tus * copies, INDEX[name]["w"], INDEX[name]["h"])
    for _ in range(3):
        rc, rep, digests = av1recon.verify_buffer(blob, host_threads=4)
        assert rc == 0 and rep.status == 0, rep.message
        assert rep.frames == INDEX[name]["frames"] * copies
    dec = _gpu_decode(name)
    want = [[int(x) for x in r.checksum] for r in dec.results]
    dec.close()
    for i in range(len(digests)):
        assert [int(x) for x in digests[i]] == want[i % len(want)], f"frame {i}"


FULL_SIZE = ["c1", "c2", "c3_small", "c3", "c4", "screen1080"] + [f"c5_{i:02d}" for i in range(0, 32, 4)]


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.parametrize("clip", FULL_SIZE)
def test_cuda_full_size_clip_md5_vs_dav1d(built, clip):
    """All five BASELINE configs at full size (c5: every fourth file of the batch): per-frame **MD5 of the Y, U and V planes** from
    the CUDA engine in parity mode (av1r_submit_tu / av1r_collect, parity_md5 = 1) equals the MD5 of libdav1d's output planes for the
    same frame -- north_star's bit-exactness criterion.  The clips must be in streams_cache/ (they travel with the repo snapshot)."""
    import av1recon
    from oracle import dav1d_ref
    from tools.make_streams import clip_path
    from tools.obuio import read_ivf
    path = clip_path(clip)
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    tus = read_ivf(path)
    want = dav1d_ref.decode_md5(tus, n_threads=os.cpu_count() or 4)
    dec = av1recon.Decoder(parity_md5=1, streams=8, frames_in_flight=16)
    for i, tu in enumerate(tus):
        dec.submit(tu, i)
    dec.flush()
    got = [[bytes(r.md5[p]).hex() for p in range(3)] for r in dec.results]
    dec.close()
    assert len(got) == len(want) > 0
    for i in range(len(want)):
        assert got[i] == want[i], f"{clip} frame {i}: plane MD5 differs from libdav1d"


@pytest.mark.gpu
@pytest.mark.timeout(900)
@pytest.mark.parametrize("clip", ["c1", "c2", "c3", "c4", "c5_00", "screen1080"])
def test_cuda_full_size_verify_digests_vs_dav1d(built, clip):
    """The segment-parallel verify path (what bench.py's e2e times) on the full-size clips: every plane digest of every frame from
    av1r_verify_buffer equals the digest of libdav1d's output for the same frame (position-salted 64-bit hash; host restatement in
    the library)."""
    import ctypes as C
    import av1recon
    from oracle import dav1d_ref
    from tools.make_streams import clip_path
    path = clip_path(clip)
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    data = open(path, "rb").read()
    from tools.obuio import read_ivf
    tus = read_ivf(path)
    rc, rep, digests = av1recon.verify_buffer(data)
    assert rc == 0 and rep.status == 0, rep.message
    l = av1recon.lib()
    l.av1r_plane_checksum_host.restype = C.c_uint64
    l.av1r_plane_checksum_host.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
    ref = dav1d_ref.decode(tus, n_threads=os.cpu_count() or 4)
    assert len(ref) == rep.frames == len(digests)
    bpc = rep.bit_depth
    for i, fr in enumerate(ref):
        for p in range(3):
            a = np.ascontiguousarray(fr[4][p].astype(np.uint8 if bpc == 8 else "<u2"))
            want = l.av1r_plane_checksum_host(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], bpc)
            assert int(digests[i][p]) == int(want), f"{clip} frame {i} plane {p}: digest differs from libdav1d"


def test_baseline_clips_exercise_their_configs_tools(built):
    """The clips in streams_cache/ really contain what their BASELINE config names (the 4K10 clip held no OBMC, no
    masked compound and no loop restoration).  Same gate bench.py applies before it times a clip.  Host parser only."""
    import av1recon
    import importlib.util
    from tools.make_streams import clip_path
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    checked = 0
    for key in ("c1", "c2", "c3", "c4"):
        path = clip_path(key)
        if not os.path.exists(path):
            continue
        info = av1recon.parse_stats(open(path, "rb").read())
        bench.check_clip_tools(key, info, av1recon.tool_hist(info))
        checked += 1
    # c5: gate on the sum over the files that exist
    tot, n = None, 0
    for i in range(32):
        path = clip_path(f"c5_{i:02d}")
        if not os.path.exists(path):
            continue
        info = av1recon.parse_stats(open(path, "rb").read())
        n += 1
        if tot is None:
            tot = info
        else:
            for f in ("lr_frames", "cdef_frames", "deblock_frames", "grain_frames"):
                setattr(tot, f, getattr(tot, f) + getattr(info, f))
            for k in range(24):
                tot.tool_hist[k] += info.tool_hist[k]
    if n:
        bench.check_clip_tools("c5", tot, av1recon.tool_hist(tot))
        checked += 1
    if not checked:
        pytest.skip("no BASELINE-size clips in streams_cache/")
    # and the gate does refuse: the all-intra clip cannot stand for the inter config
    path = clip_path("c2")
    if os.path.exists(path):
        info = av1recon.parse_stats(open(path, "rb").read())
        with pytest.raises(bench.ClipLacksTools):
            bench.check_clip_tools("c3", info, av1recon.tool_hist(info))

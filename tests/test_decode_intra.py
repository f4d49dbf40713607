"""Whole-path parity on intra-only streams (BASELINE configs[1] shape): per-frame per-plane MD5 vs
libdav1d.  CPU: oracle (host parser + scalar reconstruction) vs golden MD5 and vs live dav1d with the
in-loop filters toggled (stage isolation).  GPU: the CUDA engine through the C ABI vs the same."""
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
INDEX = json.load(open(os.path.join(GOLD, "index.json")))


def _tus(name):
    from tools.obuio import read_ivf
    return read_ivf(os.path.join(GOLD, name + ".ivf"))


def _md5(planes, bpc):
    out = []
    for p in planes:
        a = np.ascontiguousarray(p.astype(np.uint8 if bpc == 8 else "<u2"))
        out.append(hashlib.md5(a.tobytes()).hexdigest())
    return out


@pytest.mark.parametrize("name", sorted(INDEX))
def test_oracle_matches_golden_md5(built, name):
    from oracle import oracle_lib
    meta = INDEX[name]
    frames, info = oracle_lib.decode_stream(_tus(name))
    assert len(frames) == meta["frames"]
    for i, planes in enumerate(frames):
        assert _md5(planes, meta["bpc"])[:len(meta["md5"][i])] == meta["md5"][i], f"{name} frame {i}"


def test_screen_content_goldens_use_palette_and_block_copy(built):
    """The screen-content streams really contain palette and intra-block-copy blocks, and the wavefront kernel's plan for them
    (decode-order unit table, every block vector's source units listed before the reader) passes the host-side check."""
    import ctypes as C
    import av1recon
    l = av1recon.lib()
    for name, tools in (("intra_8b_palette_320x192", ("palette",)), ("intra_8b_intrabc_320x192", ("palette", "intrabc")),
                        ("intra_10b_intrabc_328x200", ("intrabc",)), ("intra_8b_intrabc_sb128_456x264", ("intrabc",)),
                        ("intra_8b_intrabc_edge_322x182", ("intrabc",))):
        blob = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
        hist = av1recon.tool_hist(av1recon.parse_stats(blob))
        for t in tools:
            assert hist.get(t, 0) > 50, f"{name}: {hist}"
        fr, un, msg = C.c_longlong(), C.c_longlong(), C.create_string_buffer(256)
        assert l.av1r_debug_k3_check(blob, len(blob), C.byref(fr), C.byref(un), msg, 256) == 0, msg.value
        assert fr.value == INDEX[name]["frames"]


def test_block_copy_into_the_future_is_a_bitstream_error(built):
    """A block vector whose source units are not listed before the reader's unit must fail the plan (-2), never reach the kernel:
    built here from a hand-made record list (two units side by side, the left one copying from the right one)."""
    import ctypes as C
    import av1recon
    l = av1recon.lib()
    if not hasattr(l, "av1r_debug_k3_ibc_selftest"):
        pytest.skip("self-test entry point not built")
    l.av1r_debug_k3_ibc_selftest.restype = C.c_int
    assert l.av1r_debug_k3_ibc_selftest(0) > 0      # source to the left (earlier unit): plan accepted
    assert l.av1r_debug_k3_ibc_selftest(1) == -2    # source to the right (later unit): rejected


@pytest.mark.parametrize("filters", [0, 1, 3, 7])
def test_oracle_stage_isolation_vs_dav1d(built, filters):
    """inloop_filters = 0 (recon only), 1 (+deblock), 3 (+CDEF), 7 (+LR): oracle == dav1d at every stage."""
    from oracle import dav1d_ref, oracle_lib
    for name in ("intra_8b_200x136", "intra_10b_192x128", "intra_8b_lr_480x272"):
        tus = _tus(name)
        ref = dav1d_ref.decode(tus, inloop_filters=filters, apply_grain=0)
        got, _ = oracle_lib.decode_stream(tus, inloop_filters=filters, apply_grain=0)
        assert len(ref) == len(got)
        for i in range(len(ref)):
            for p in range(3):
                assert np.array_equal(ref[i][4][p], got[i][p]), f"{name} filters={filters} frame {i} plane {p}"


def test_probe_and_headers(built):
    import av1recon
    for name, meta in INDEX.items():
        data = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
        info = av1recon.probe_buffer(data)
        assert info.is_av1 == 1 and info.width == meta["w"] and info.height == meta["h"] and info.bit_depth == meta["bpc"]
        assert info.temporal_units == meta["frames"] and info.keyframes == meta["frames"]


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
        assert r.status == 0 and r.w == meta["w"] and r.h == meta["h"] and r.bpc == meta["bpc"]
        got = [bytes(r.md5[p]).hex() for p in range(len(meta["md5"][i]))]
        if got != meta["md5"][i]:
            # locate the first differing sample to make the failure actionable
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
@pytest.mark.parametrize("filters", [0, 1, 3, 7])
def test_cuda_stage_isolation(built, filters):
    """CUDA engine with inloop_filters masked, vs the oracle at the same stage (oracle == dav1d above)."""
    from oracle import oracle_lib
    for name in ("intra_8b_200x136", "intra_10b_192x128", "intra_8b_sb128_264x200", "intra_8b_lr_480x272"):
        ref, _ = oracle_lib.decode_stream(_tus(name), inloop_filters=filters, apply_grain=0)
        dec = _gpu_decode(name, inloop_filters=filters, apply_grain=0)
        assert len(dec.results) == len(ref)
        for i, r in enumerate(dec.results):
            planes = dec.frame_planes(r)
            for p in range(3):
                bad = np.argwhere(planes[p].astype(np.int32) != ref[i][p].astype(np.int32))
                assert len(bad) == 0, (f"{name} filters={filters} frame {i} plane {p}: {len(bad)} px differ, first (y,x)={tuple(bad[0])} "
                                       f"cuda={planes[p][tuple(bad[0])]} ref={ref[i][p][tuple(bad[0])]}")
        dec.close()


@pytest.mark.gpu
def test_cuda_checksum_matches_host_digest(built):
    import av1recon
    name = "intra_8b_200x136"
    dec = _gpu_decode(name)
    l = av1recon.lib()
    for r in dec.results:
        planes = dec.frame_planes(r)
        for p in range(3):
            a = np.ascontiguousarray(planes[p])
            want = l.av1r_plane_checksum_host(a.ctypes.data, a.strides[0], a.shape[1], a.shape[0], r.bpc)
            assert r.checksum[p] == want
    dec.close()


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 4])
def test_cuda_verify_buffer_segment_parallel(built, threads):
    """av1r_verify_buffer (GOP-segment-parallel host parse) returns the same plane digests, in display
    order, as the single-threaded streaming API."""
    import av1recon
    for name in ("intra_8b_200x136", "intra_8b_tiles_320x192", "intra_8b_grain_160x96"):
        data = open(os.path.join(GOLD, name + ".ivf"), "rb").read()
        rc, rep, digs = av1recon.verify_buffer(data, host_threads=threads)
        assert rc == 0, rep.message
        assert rep.frames == INDEX[name]["frames"] and rep.width == INDEX[name]["w"]
        dec = _gpu_decode(name)
        want = [tuple(r.checksum) for r in dec.results]
        dec.close()
        assert digs == want


@pytest.mark.gpu
def test_cuda_verify_reports_corruption(built):
    """A truncated / corrupted stream must come back as an error with first_bad_frame set, like a failed
    RunTranscode sets job.Reason (/root/reference/internal/daemon/daemon.go:102-112)."""
    import av1recon
    data = bytearray(open(os.path.join(GOLD, "intra_8b_200x136.ivf"), "rb").read())
    cut = bytes(data[: len(data) // 2])
    rc, rep, _ = av1recon.verify_buffer(cut, host_threads=2)
    assert rc != 0 and rep.status == rc and len(rep.message) > 0

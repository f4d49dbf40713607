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
# host-only runs also mutate the streams of the less common syntax paths (segmentation, delta_q / delta_lf, lossless, a new frame size
# on every frame, separate tile group OBUs, screen content); the whole set of 57 was run through the sanitizer build with 6,100
# mutants without a report (tools/fuzz_goldens.py; not part of the suite for its run time)
NAMES_HOST = NAMES + ["inter_8b_aq1_256x160", "inter_8b_aq3_256x160", "inter_8b_deltaq_lf_256x160", "inter_8b_lossless_128x96",
                      "inter_8b_resize_dyn_352x288", "inter_8b_superres_rand_352x288", "inter_8b_tilegroups_352x288",
                      "inter_8b_screen_352x288", "intra_8b_intrabc_sb128_456x264", "inter_10b_tiles4x2_odd_410x230"]


def _mutants(seed, n, names=NAMES):
    rng = random.Random(seed)
    for _ in range(n):
        name = rng.choice(names)
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
    for name, kind, data in list(_mutants(20261018, 60)) + list(_mutants(99, 80, NAMES_HOST)):
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


# ---- Matroska: the container the daemon really hands over (a single mutated byte read 16 KB past the buffer) ----------
def _mkv_mutants(seed, n):
    from tools import mkvmux
    from tools.obuio import read_ivf
    import json
    idx = dict(json.load(open(os.path.join(GOLD, "index.json"))))
    idx.update(json.load(open(os.path.join(GOLD, "index_inter.json"))))
    rng = random.Random(seed)
    base = {}
    for name in ("intra_8b_200x136", "inter_8b_sb128_tiles_640x360"):
        tus = read_ivf(os.path.join(GOLD, name + ".ivf"))
        for k, kw in enumerate((dict(), dict(ffmpeg_like=True, block_groups=True), dict(ffmpeg_like=True, unknown_size_clusters=True),
                                dict(video_lacing="ebml"), dict(video_lacing="xiph"))):
            base[(name, k)] = mkvmux.mux(tus, idx[name]["w"], idx[name]["h"], frames_per_cluster=3, **kw)
    keys = sorted(base)
    for i in range(n):
        key = keys[i % len(keys)]
        data = bytearray(base[key])
        kind = rng.choice(["hdr", "hdr", "flip", "trunc", "size"])
        if kind == "hdr":      # the element structure lives in the first few hundred bytes: hit it hard
            for _ in range(rng.randint(1, 3)):
                data[rng.randrange(0, min(len(data), 600))] = rng.randrange(256)
        elif kind == "flip":
            for _ in range(rng.randint(1, 4)):
                data[rng.randrange(0, len(data))] ^= 1 << rng.randrange(8)
        elif kind == "trunc":
            data = data[:rng.randrange(8, len(data))]
        else:                  # blow up a size field: 0xA3 / 0x63A2 / cluster ids followed by a vint
            hits = [p for p in range(len(data) - 9) if data[p] in (0xA3, 0xA0, 0xA1, 0xAE) or data[p:p + 2] == b"\x63\xA2"]
            p = rng.choice(hits) + (2 if data[rng.choice(hits)] == 0x63 else 1)
            data[p:p + 8] = bytes([0x01]) + bytes(rng.randrange(256) for _ in range(7))
        yield key, kind, bytes(data)


def test_host_demux_and_parse_survive_mutated_matroska(built):
    import av1recon
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    codes = {}
    for key, kind, data in _mkv_mutants(99, 200):
        rep = av1recon.Report()
        rc = l.av1r_parse_buffer(data, len(data), 1, 0, C.byref(rep))
        assert rc <= 0, (key, kind, rc)
        try:
            av1recon.probe_buffer(data)
        except RuntimeError:
            pass
        codes[rc] = codes.get(rc, 0) + 1
    assert len(codes) > 1, codes


def test_sanitizer_build_finds_no_out_of_bounds_access(built, tmp_path):
    """The host half (demux + parser) compiled with -fsanitize=address,undefined over mutated Matroska and IVF inputs: the run must
    end with exit code 0 (an ASan / UBSan report aborts the process)."""
    import subprocess
    exe = os.path.join(ROOT, "build", "asan", "parse_fuzz")
    subprocess.check_call(["make", "build/asan/parse_fuzz"], cwd=ROOT, stdout=subprocess.DEVNULL)
    files = []
    for i, (key, kind, data) in enumerate(_mkv_mutants(4242, 120)):
        p = tmp_path / f"m{i}.mkv"
        p.write_bytes(data)
        files.append(str(p))
    for i, (name, kind, data) in enumerate(list(_mutants(555, 40)) + list(_mutants(556, 60, NAMES_HOST))):
        p = tmp_path / f"s{i}.ivf"
        p.write_bytes(data)
        files.append(str(p))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe] + files, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
    assert "ok=" in r.stdout


# ---- frames with missing tile data must not verify as "ok" (stream_parser.cpp) ------------------------------------
def _obus(tu):
    from tools import mkvmux
    return mkvmux.split_obus(tu)


def _leb(n):
    out = b""
    while True:
        b = n & 0x7F
        n >>= 7
        out += bytes([b | (0x80 if n else 0)])
        if not n:
            return out


def _frame_to_bare_header(tus, i):
    """Rewrites the OBU_FRAME of temporal unit i as an OBU_FRAME_HEADER that keeps the header bytes and drops the tile data."""
    import av1recon
    tu = tus[i]
    out = b""
    for typ, raw in _obus(tu):
        if typ != 6:
            out += raw
            continue
        nbytes = [h for h in av1recon.scan_headers(tus) if h.tu_index == i][-1].header_bytes
        # payload starts after the 1-byte header + leb128 size
        p = 1
        while raw[p] & 0x80:
            p += 1
        p += 1
        out += bytes([(3 << 3) | 2]) + _leb(nbytes) + raw[p:p + nbytes]
    return out


@pytest.mark.parametrize("which", ["first", "last"])
def test_frame_header_without_tile_data_is_an_error(built, which):
    import av1recon
    from tools.obuio import read_ivf, write_ivf  # noqa: F401
    from av1recon import shard
    tus = read_ivf(os.path.join(GOLD, "inter_8b_base_192x128.ivf"))
    i = 0 if which == "first" else len(tus) - 1
    tus = list(tus)
    tus[i] = _frame_to_bare_header(tus, i)
    blob = shard.ivf_bytes(tus, 192, 128)
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    rep = av1recon.Report()
    rc = l.av1r_parse_buffer(blob, len(blob), 1, 0, C.byref(rep))
    assert rc == -74, (rc, rep.frames, rep.message)      # AV1R_EBITSTREAM, not "ok with one frame fewer"


def test_out_of_order_tile_groups_are_an_error(built):
    """A frame coded as OBU_FRAME_HEADER + tile groups: repeating the first tile group instead of sending the second one must not
    complete the frame."""
    import av1recon
    from tools.obuio import read_ivf
    from av1recon import shard
    name = "intra_8b_tiles_320x192"
    tus = list(read_ivf(os.path.join(GOLD, name + ".ivf")))
    # split the OBU_FRAME into FRAME_HEADER + one OBU_TILE_GROUP per tile would need the tile sizes; instead duplicate the whole
    # OBU_FRAME's tile payload as an extra OBU_TILE_GROUP claiming tiles [0, n): tg_start (0) != tiles already done (n)
    obus = _obus(tus[0])
    typ, raw = [(t, r) for t, r in obus if t == 6][0]
    hdr_bytes = av1recon.scan_headers([tus[0]])[-1].header_bytes
    p = 1
    while raw[p] & 0x80:
        p += 1
    p += 1
    tile_payload = raw[p + hdr_bytes:]
    header_only = bytes([(3 << 3) | 2]) + _leb(hdr_bytes) + raw[p:p + hdr_bytes]
    tg = bytes([(4 << 3) | 2]) + _leb(len(tile_payload)) + tile_payload
    good = b"".join(r for t, r in obus if t != 6) + header_only + tg
    bad = b"".join(r for t, r in obus if t != 6) + header_only + tg + tg
    l = av1recon.lib()
    l.av1r_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(av1recon.Report)]
    for blob_tus, want_ok in (([good] + tus[1:], True), ([bad] + tus[1:], False)):
        blob = shard.ivf_bytes(blob_tus, 320, 192)
        rep = av1recon.Report()
        rc = l.av1r_parse_buffer(blob, len(blob), 1, 0, C.byref(rep))
        assert (rc == 0) == want_ok, (want_ok, rc, rep.message)

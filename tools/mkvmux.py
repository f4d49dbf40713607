"""Minimal Matroska (EBML) muxer for tests: wraps AV1 temporal units the way the daemon's FFmpeg child does (`-f matroska`,
/root/reference/internal/ffmpeg/transcode.go:143): EBML header, Segment with one V_AV1 track whose CodecPrivate is the av1C box
(4-byte header + the sequence header OBU), and Clusters of SimpleBlocks from which the temporal delimiters have been removed."""
import struct


def _vint(n):
    for length in range(1, 9):
        if n < (1 << (7 * length)) - 1:
            return bytes([(1 << (8 - length)) | (n >> (8 * (length - 1)))]) + n.to_bytes(length, "big")[1:] if length > 1 else bytes([0x80 | n])
    raise ValueError(n)


def _el(eid, payload):
    return bytes.fromhex(eid) + _vint(len(payload)) + payload


def _uint(v):
    n = max(1, (v.bit_length() + 7) // 8)
    return v.to_bytes(n, "big")


def split_obus(tu):
    """-> list of (obu_type, bytes) of a temporal unit (all OBUs carry a size field in libaom output)."""
    out, pos = [], 0
    while pos < len(tu):
        hdr = tu[pos]
        typ, ext, has_size = (hdr >> 3) & 15, (hdr >> 2) & 1, (hdr >> 1) & 1
        p = pos + 1 + ext
        assert has_size
        size, shift = 0, 0
        while True:
            b = tu[p]
            p += 1
            size |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        out.append((typ, tu[pos:p + size]))
        pos = p + size
    return out


def _ebml_lace_sizes(sizes):
    """EBML lacing: first size as a vint, the following ones as signed vint differences (the last frame's size is implied)."""
    out = _vint(sizes[0])
    for prev, cur in zip(sizes[:-2], sizes[1:-1]):
        diff = cur - prev
        for length in range(1, 9):
            bias = (1 << (7 * length - 1)) - 1
            if -bias < diff <= bias:
                v = diff + bias
                out += bytes([(1 << (8 - length)) | (v >> (8 * (length - 1)))]) + v.to_bytes(length, "big")[1:] if length > 1 else bytes([0x80 | v])
                break
    return out


def _xiph_lace_sizes(sizes):
    out = b""
    for s in sizes[:-1]:
        out += b"\xff" * (s // 255) + bytes([s % 255])
    return out


def _block(track, rel_ts, flags, frames, lacing=None):
    """SimpleBlock / Block payload.  frames: list of byte strings; lacing None | 'xiph' | 'ebml' | 'fixed'."""
    hdr = _vint(track) + struct.pack(">h", rel_ts)
    if lacing is None or len(frames) == 1:
        return hdr + bytes([flags]) + frames[0]
    code = {"xiph": 1, "fixed": 2, "ebml": 3}[lacing]
    sizes = [len(f) for f in frames]
    lace = {"xiph": _xiph_lace_sizes, "ebml": _ebml_lace_sizes, "fixed": lambda s: b""}[lacing](sizes)
    return hdr + bytes([flags | (code << 1), len(frames) - 1]) + lace + b"".join(frames)


def mux(tus, w, h, frames_per_cluster=8, unknown_size_clusters=False, strip_td=True, ffmpeg_like=False, video_lacing=None,
        block_groups=False):
    """ffmpeg_like: the layout FFmpeg's matroska muxer writes for a transcode with audio -- SeekHead, Void padding, Info, Tracks
    with an audio track *first* (track 1 A_OPUS, track 2 V_AV1), Tags, then Clusters that interleave laced audio blocks with the
    video, and Cues at the end.  video_lacing: pack pairs of temporal units of the AV1 track into one laced block (lacing is legal
    EBML even if FFmpeg never laces video).  block_groups: BlockGroup(Block, BlockDuration) instead of SimpleBlock for non-key video."""
    seq = next(o for t, o in split_obus(tus[0]) if t == 1)
    av1c = bytes([0x81, 0x00, 0x0C, 0x00]) + seq      # marker/version, profile/level, flags, no initial presentation delay
    ebml = _el("1A45DFA3", _el("4286", _uint(1)) + _el("42F7", _uint(1)) + _el("42F2", _uint(4)) + _el("42F3", _uint(8)) +
               _el("4282", b"matroska") + _el("4287", _uint(4)) + _el("4285", _uint(2)))
    info = _el("1549A966", _el("2AD7B1", _uint(1000000)) + _el("4D80", b"av1r-test") + _el("5741", b"av1r-test"))
    video = _el("E0", _el("B0", _uint(w)) + _el("BA", _uint(h)))
    vtrack_no = 2 if ffmpeg_like else 1
    track = _el("AE", _el("D7", _uint(vtrack_no)) + _el("73C5", _uint(vtrack_no)) + _el("83", _uint(1)) + _el("86", b"V_AV1") + _el("63A2", av1c) + video)
    if ffmpeg_like:
        opus_head = b"OpusHead" + bytes([1, 2]) + struct.pack("<HIhB", 312, 48000, 0, 0)
        audio = _el("AE", _el("D7", _uint(1)) + _el("73C5", _uint(77)) + _el("83", _uint(2)) + _el("86", b"A_OPUS") + _el("63A2", opus_head) +
                    _el("E1", _el("B5", struct.pack(">f", 48000.0)) + _el("9F", _uint(2))))
        tracks = _el("1654AE6B", audio + track)
    else:
        tracks = _el("1654AE6B", track)

    def payload_of(tu):
        return b"".join(o for t, o in split_obus(tu) if not (strip_td and t == 2))

    clusters = b""
    cue_points = b""
    for c0 in range(0, len(tus), frames_per_cluster):
        body = _el("E7", _uint(c0 * 33))
        group = tus[c0:c0 + frames_per_cluster]
        i = 0
        while i < len(group):
            if ffmpeg_like:   # interleaved audio: Opus packets, EBML- / Xiph- / fixed-laced in turn
                pk = [bytes([0xFC, (c0 + i + k) & 0xFF]) + bytes((7 * (c0 + i) + 3 * k) % 23 + 1) for k in range(3)]
                kind = ("ebml", "xiph", "fixed")[(c0 + i) % 3]
                if kind == "fixed":
                    pk = [p[:3].ljust(3, b"\0") for p in pk]
                body += _el("A3", _block(1, i * 33, 0x80, pk, kind))
            n = 2 if (video_lacing and i + 1 < len(group)) else 1
            frames = [payload_of(t) for t in group[i:i + n]]
            key = 0x80 if any(t == 1 for t, _ in split_obus(group[i])) else 0
            blk = _block(vtrack_no, i * 33, key, frames, video_lacing if n == 2 else None)
            if block_groups and not key:
                body += _el("A0", _el("A1", blk[:3] + bytes([blk[3] & 0x7F]) + blk[4:]) + _el("9B", _uint(33)) + _el("FB", struct.pack(">b", -33)))
            else:
                body += _el("A3", blk)
            i += n
        cue_points += _el("BB", _el("B3", _uint(c0 * 33)) + _el("B7", _el("F7", _uint(vtrack_no)) + _el("F1", _uint(len(clusters)))))
        if unknown_size_clusters:
            clusters += bytes.fromhex("1F43B675") + b"\x01\xff\xff\xff\xff\xff\xff\xff" + body
        else:
            clusters += _el("1F43B675", body)
    if ffmpeg_like:
        seek = _el("114D9B74", b"".join(_el("4DBB", _el("53AB", bytes.fromhex(i_)) + _el("53AC", _uint(4096 + k)))
                                        for k, i_ in enumerate(("1549A966", "1654AE6B", "1254C367", "1C53BB6B"))))
        void = _el("EC", bytes(87))
        tags = _el("1254C367", _el("7373", _el("63C0", _el("63C5", _uint(77))) + _el("67C8", _el("45A3", b"ENCODER") + _el("4487", b"Lavf62.3.100"))))
        cues = _el("1C53BB6B", cue_points)
        seg_payload = seek + void + info + tracks + tags + clusters + cues
    else:
        seg_payload = info + tracks + clusters
    if unknown_size_clusters:
        return ebml + bytes.fromhex("18538067") + b"\x01\xff\xff\xff\xff\xff\xff\xff" + seg_payload
    return ebml + _el("18538067", seg_payload)

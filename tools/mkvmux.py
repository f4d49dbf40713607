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


def mux(tus, w, h, frames_per_cluster=8, unknown_size_clusters=False, strip_td=True):
    seq = next(o for t, o in split_obus(tus[0]) if t == 1)
    av1c = bytes([0x81, 0x00, 0x0C, 0x00]) + seq      # marker/version, profile/level, flags, no initial presentation delay
    ebml = _el("1A45DFA3", _el("4286", _uint(1)) + _el("42F7", _uint(1)) + _el("42F2", _uint(4)) + _el("42F3", _uint(8)) +
               _el("4282", b"matroska") + _el("4287", _uint(4)) + _el("4285", _uint(2)))
    info = _el("1549A966", _el("2AD7B1", _uint(1000000)) + _el("4D80", b"av1r-test") + _el("5741", b"av1r-test"))
    video = _el("E0", _el("B0", _uint(w)) + _el("BA", _uint(h)))
    track = _el("AE", _el("D7", _uint(1)) + _el("73C5", _uint(1)) + _el("83", _uint(1)) + _el("86", b"V_AV1") + _el("63A2", av1c) + video)
    tracks = _el("1654AE6B", track)
    clusters = b""
    for c0 in range(0, len(tus), frames_per_cluster):
        body = _el("E7", _uint(c0 * 33))
        for i, tu in enumerate(tus[c0:c0 + frames_per_cluster]):
            payload = b"".join(o for t, o in split_obus(tu) if not (strip_td and t == 2))
            key = 0x80 if any(t == 1 for t, _ in split_obus(tu)) else 0
            body += _el("A3", _vint(1) + struct.pack(">hB", i * 33, key) + payload)
        if unknown_size_clusters:
            clusters += bytes.fromhex("1F43B675") + b"\x01\xff\xff\xff\xff\xff\xff\xff" + body
        else:
            clusters += _el("1F43B675", body)
    seg_payload = info + tracks + clusters
    if unknown_size_clusters:
        return ebml + bytes.fromhex("18538067") + b"\x01\xff\xff\xff\xff\xff\xff\xff" + seg_payload
    return ebml + _el("18538067", seg_payload)

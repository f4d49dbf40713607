"""GOP-segment work items and their assignment to GPUs: a shown KEY_FRAME refreshes every reference slot and resets
the CDFs, so key-frame-delimited segments (and separate files) decode independently -- ranks share nothing and no collective is on
the data path.  Mirrors what the daemon's batch loop (/root/reference/cmd/av1d/main.go:312-349) would do with one engine per GPU."""
import struct


def split_segments(tus, headers):
    """tus: temporal units of one file; headers: av1recon.scan_headers(tus).  -> list of (first_tu, last_tu_exclusive)."""
    first_hdr = {}
    for h in headers:
        first_hdr.setdefault(h.tu_index, h)
    starts = [i for i in range(len(tus)) if i in first_hdr and first_hdr[i].frame_type == 0 and first_hdr[i].show_frame
              and not first_hdr[i].show_existing_frame]
    if not starts or starts[0] != 0:
        starts = [0] + starts
    return [(s, starts[k + 1] if k + 1 < len(starts) else len(tus)) for k, s in enumerate(starts)]


def assign(items, world):
    """Longest-processing-time-first assignment of weighted items to `world` ranks.  items: list of (key, weight).
    Returns a list of `world` lists of keys; deterministic, identical on every rank (no communication needed)."""
    order = sorted(range(len(items)), key=lambda i: (-items[i][1], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(items[i][0])
        load[r] += items[i][1]
    for r in range(world):
        out[r].sort()
    return out


def ivf_bytes(tus, w, h, fps=30):
    """An IVF container holding the given temporal units (what av1r_ctx_verify_buffer takes)."""
    parts = [struct.pack("<4sHH4sHHIIII", b"DKIF", 0, 32, b"AV01", w, h, fps, 1, len(tus), 0)]
    for i, tu in enumerate(tus):
        parts.append(struct.pack("<IQ", len(tu), i))
        parts.append(tu)
    return b"".join(parts)

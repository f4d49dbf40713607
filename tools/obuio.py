"""IVF container I/O for AV1 temporal units (32-byte file header, 12-byte frame headers)."""
import struct


def write_ivf(path, tus, w, h, fps=30):
    with open(path, "wb") as f:
        f.write(struct.pack("<4sHH4sHHIIII", b"DKIF", 0, 32, b"AV01", w, h, fps, 1, len(tus), 0))
        for i, tu in enumerate(tus):
            f.write(struct.pack("<IQ", len(tu), i))
            f.write(tu)


def read_ivf(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"DKIF":
        raise ValueError("not an IVF file")
    hl = struct.unpack_from("<H", data, 6)[0]
    pos = hl
    tus = []
    while pos + 12 <= len(data):
        sz, _pts = struct.unpack_from("<IQ", data, pos)
        pos += 12
        tus.append(data[pos:pos + sz])
        pos += sz
    return tus


def ivf_header(path):
    with open(path, "rb") as f:
        h = f.read(32)
    _, _, _, fourcc, w, hh, num, den, n, _ = struct.unpack("<4sHH4sHHIIII", h)
    return dict(fourcc=fourcc, w=w, h=hh, n=n)

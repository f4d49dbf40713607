"""Procedural 4:2:0 YUV sources (numpy) standing in for ffmpeg's lavfi `testsrc2` / noise
(the idiom of /root/reference/internal/ffmpeg/binary.go:286; no ffmpeg exists in this image)."""
import numpy as np


def _to420(y, u, v, bpc):
    mx = (1 << bpc) - 1
    sc = 1 << (bpc - 8)
    Y = np.clip(y * sc, 0, mx)
    U = np.clip(u * sc, 0, mx)
    V = np.clip(v * sc, 0, mx)
    dt = np.uint8 if bpc == 8 else np.uint16
    U = (U[0::2, 0::2] + U[1::2, 0::2] + U[0::2, 1::2] + U[1::2, 1::2]) / 4
    V = (V[0::2, 0::2] + V[1::2, 0::2] + V[0::2, 1::2] + V[1::2, 1::2]) / 4
    return [np.rint(Y).astype(dt), np.rint(U).astype(dt), np.rint(V).astype(dt)]


def testsrc2_like(w, h, n, bpc=8, seed=1):
    """Colour bars + moving gradient + sweeping box + per-frame counter block + texture."""
    rng = np.random.default_rng(seed)
    xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    bars_y = np.array([235, 210, 170, 145, 106, 81, 41, 16], np.float32)
    bars_u = np.array([128, 16, 166, 54, 202, 90, 240, 128], np.float32)
    bars_v = np.array([128, 146, 16, 34, 222, 240, 110, 128], np.float32)
    bi = np.minimum((xx * 8 // w).astype(np.int32), 7)
    tex = rng.integers(0, 24, size=(h // 8 + 1, w // 8 + 1)).astype(np.float32)
    tex = np.kron(tex, np.ones((8, 8), np.float32))[:h, :w]
    for t in range(n):
        y = bars_y[bi].copy()
        u = bars_u[bi].copy()
        v = bars_v[bi].copy()
        # moving diagonal gradient in the middle band
        band = (yy > h * 0.35) & (yy < h * 0.65)
        g = ((xx + yy * 0.5 + t * 7) % 256)
        y[band] = g[band]
        u[band] = 128 + 40 * np.sin((xx[band] + 3 * t) / 37.0)
        v[band] = 128 + 40 * np.cos((yy[band] - 2 * t) / 23.0)
        # textured band at the bottom (static texture sliding slowly)
        bot = yy > h * 0.8
        y[bot] = 100 + np.roll(tex, 2 * t, axis=1)[bot] * 4
        # sweeping box
        bx = int((t * 13) % max(1, w - w // 8))
        by = int(h * 0.1 + (t * 5) % max(1, h // 2))
        y[by:by + h // 8, bx:bx + w // 8] = 200 - (t * 3) % 100
        u[by:by + h // 8, bx:bx + w // 8] = 90
        v[by:by + h // 8, bx:bx + w // 8] = 200
        # counter block: 8 binary cells
        for b in range(8):
            if (t >> b) & 1:
                y[8:40, 8 + 40 * b: 40 + 40 * b] = 235
            else:
                y[8:40, 8 + 40 * b: 40 + 40 * b] = 16
        yield _to420(y, u, v, bpc)


def noise_gradient(w, h, n, bpc=8, seed=4, sigma=6.0):
    """Smooth gradient + seeded Gaussian noise (film-grain style source)."""
    rng = np.random.default_rng(seed)
    xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    for t in range(n):
        y = 40 + 160 * (xx / w) * (0.5 + 0.5 * yy / h) + 10 * np.sin((xx + 4 * t) / 50.0)
        u = 128 + 50 * np.sin(yy / 90.0 + t / 20.0)
        v = 128 + 50 * np.cos(xx / 120.0)
        y = y + rng.normal(0, sigma, size=y.shape).astype(np.float32)
        u = u + rng.normal(0, sigma / 2, size=y.shape).astype(np.float32)
        v = v + rng.normal(0, sigma / 2, size=y.shape).astype(np.float32)
        yield _to420(y, u, v, bpc)


def pan_zoom(w, h, n, bpc=8, seed=3):
    """Textured image with global pan + slow zoom/rotation and independently moving patches."""
    rng = np.random.default_rng(seed)
    big = 1 << int(np.ceil(np.log2(max(w, h) * 1.5)))
    base = rng.normal(0, 1, size=(big // 16, big // 16)).astype(np.float32)
    tex = np.kron(base, np.ones((16, 16), np.float32))
    fine = rng.normal(0, 1, size=(big // 4, big // 4)).astype(np.float32)
    tex = tex * 30 + np.kron(fine, np.ones((4, 4), np.float32)) * 12 + 128
    # blur a little so that sub-pel motion is meaningful
    tex = (tex + np.roll(tex, 1, 0) + np.roll(tex, 1, 1) + np.roll(np.roll(tex, 1, 0), 1, 1)) / 4
    cu = rng.normal(128, 25, size=(big // 32, big // 32)).astype(np.float32)
    cu = np.kron(cu, np.ones((32, 32), np.float32))
    cv = np.roll(cu, 77, axis=1)[::-1]
    xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    cx, cy = w / 2, h / 2
    patches = [(rng.integers(0, w - w // 6), rng.integers(0, h - h // 6), rng.integers(-6, 7), rng.integers(-4, 5))
               for _ in range(6)]
    for t in range(n):
        ang = 0.002 * t
        zoom = 1.0 + 0.003 * t
        ca, sa = np.cos(ang) / zoom, np.sin(ang) / zoom
        sx = (xx - cx) * ca - (yy - cy) * sa + cx + 2.5 * t + big / 4
        sy = (xx - cx) * sa + (yy - cy) * ca + cy + 1.25 * t + big / 4
        x0 = np.floor(sx).astype(np.int32)
        y0 = np.floor(sy).astype(np.int32)
        fx = sx - x0
        fy = sy - y0
        x0 %= big
        y0 %= big
        x1 = (x0 + 1) % big
        y1 = (y0 + 1) % big

        def samp(img):
            return (img[y0, x0] * (1 - fx) * (1 - fy) + img[y0, x1] * fx * (1 - fy)
                    + img[y1, x0] * (1 - fx) * fy + img[y1, x1] * fx * fy)

        y = samp(tex)
        u = samp(cu)
        v = samp(cv)
        for (px, py, dx, dy) in patches:
            qx = int((px + dx * t) % (w - w // 6))
            qy = int((py + dy * t) % (h - h // 6))
            y[qy:qy + h // 6, qx:qx + w // 6] = tex[py:py + h // 6, px:px + w // 6] * 0.7 + 60
            u[qy:qy + h // 6, qx:qx + w // 6] = 100
        yield _to420(y, u, v, bpc)


SOURCES = {"testsrc2": testsrc2_like, "noise": noise_gradient, "panzoom": pan_zoom}


def occluders(w, h, n, bpc=8, seed=3, n_obj=14):
    """Fine-textured background under a global pan / slow zoom, with textured foreground objects that move independently, cross
    each other (occlusion edges -> wedge / difference-weighted compound, OBMC, inter-intra at uncovered borders), two of which fade
    in brightness (distance-weighted compound), plus mild sensor-like noise (keeps loop restoration worthwhile)."""
    rng = np.random.default_rng(seed)
    big = 1 << int(np.ceil(np.log2(max(w, h) * 1.5)))
    coarse = np.kron(rng.normal(0, 1, size=(big // 32, big // 32)).astype(np.float32), np.ones((32, 32), np.float32))
    mid = np.kron(rng.normal(0, 1, size=(big // 8, big // 8)).astype(np.float32), np.ones((8, 8), np.float32))
    fine = np.kron(rng.normal(0, 1, size=(big // 2, big // 2)).astype(np.float32), np.ones((2, 2), np.float32))
    tex = coarse * 28 + mid * 14 + fine * 7 + 120
    for _ in range(2):
        tex = (tex + np.roll(tex, 1, 0) + np.roll(tex, 1, 1) + np.roll(np.roll(tex, 1, 0), 1, 1)) / 4
    cu = np.kron(rng.normal(128, 25, size=(big // 64, big // 64)).astype(np.float32), np.ones((64, 64), np.float32))
    cv = np.roll(cu, 177, axis=1)[::-1]
    xx, yy = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    cx, cy = w / 2, h / 2
    objs = []
    for i in range(n_obj):
        ow = int(rng.integers(w // 14, w // 5))
        oh = int(rng.integers(h // 10, h // 4))
        objs.append(dict(x=float(rng.integers(0, w - ow)), y=float(rng.integers(0, h - oh)), w=ow, h=oh,
                         vx=float(rng.uniform(-9, 9)), vy=float(rng.uniform(-6, 6)), tx=int(rng.integers(0, big - ow)), ty=int(rng.integers(0, big - oh)),
                         gain=float(rng.uniform(0.6, 1.2)), off=float(rng.uniform(-30, 50)), fade=(i % 5 == 0), round=(i % 3 == 0),
                         cu=float(rng.uniform(70, 190)), cv=float(rng.uniform(70, 190))))
    for t in range(n):
        zoom = 1.0 + 0.002 * t
        sx = (xx - cx) / zoom + cx + 1.75 * t + big / 4
        sy = (yy - cy) / zoom + cy + 0.75 * t + big / 4
        x0 = np.floor(sx).astype(np.int32)
        y0 = np.floor(sy).astype(np.int32)
        fx = sx - x0
        fy = sy - y0
        x0 %= big
        y0 %= big
        x1 = (x0 + 1) % big
        y1 = (y0 + 1) % big

        def samp(img):
            return (img[y0, x0] * (1 - fx) * (1 - fy) + img[y0, x1] * fx * (1 - fy) + img[y1, x0] * (1 - fx) * fy + img[y1, x1] * fx * fy)

        y = samp(tex)
        u = samp(cu)
        v = samp(cv)
        for o in objs:
            px = o["x"] + o["vx"] * t
            py = o["y"] + o["vy"] * t
            # bounce inside the frame
            rx, ry = w - o["w"], h - o["h"]
            px = abs((px % (2 * rx)) - rx) if rx > 0 else 0
            py = abs((py % (2 * ry)) - ry) if ry > 0 else 0
            qx, qy = int(px), int(py)
            patch = tex[o["ty"]:o["ty"] + o["h"], o["tx"]:o["tx"] + o["w"]] * o["gain"] + o["off"]
            if o["fade"]:
                patch = patch + 40.0 * np.sin(t / 5.0)
            if o["round"]:
                ey, ex = np.ogrid[0:o["h"], 0:o["w"]]
                m = (((ex - o["w"] / 2) / (o["w"] / 2)) ** 2 + ((ey - o["h"] / 2) / (o["h"] / 2)) ** 2) <= 1.0
                ys = y[qy:qy + o["h"], qx:qx + o["w"]]
                ys[m] = patch[m]
                us = u[qy:qy + o["h"], qx:qx + o["w"]]
                us[m] = o["cu"]
                vs = v[qy:qy + o["h"], qx:qx + o["w"]]
                vs[m] = o["cv"]
            else:
                y[qy:qy + o["h"], qx:qx + o["w"]] = patch
                u[qy:qy + o["h"], qx:qx + o["w"]] = o["cu"]
                v[qy:qy + o["h"], qx:qx + o["w"]] = o["cv"]
        y = y + rng.normal(0, 2.0, size=y.shape).astype(np.float32)
        yield _to420(y, u, v, bpc)


SOURCES["occluders"] = occluders


def screen_text(w, h, n, bpc=8, seed=5):
    """Screen content: a few flat colours, rows of small glyphs drawn from a fixed font of random 5x7 bitmaps (so the same shapes
    recur all over the frame: what palette mode and intra block copy are made for), window frames, a scrolling text pane."""
    rng = np.random.default_rng(seed)
    font = rng.integers(0, 2, size=(26, 7, 5)).astype(np.uint8)
    font[:, :, 0] |= font[:, :, 4] & font[:, :, 2]
    cols = np.array([[235, 128, 128], [16, 128, 128], [81, 90, 240], [145, 54, 34], [41, 240, 110], [180, 100, 160]], np.float32)
    lines = [rng.integers(0, 26, size=w // 6 + 2) for _ in range(h // 9 + n + 4)]
    for k in range(3, len(lines), 3):
        lines[k] = lines[k - 3].copy()      # repeated lines: long exact matches for block copy
    for t in range(n):
        idx = np.zeros((h, w), np.int32)    # colour index per sample
        pane_x0 = w // 3
        idx[:, :pane_x0] = 5
        idx[:, pane_x0:pane_x0 + 2] = 1
        idx[:12, :] = 2                      # title bar
        for li in range((h - 14) // 9):
            y0 = 14 + li * 9
            txt = lines[li + t]              # the pane scrolls one text line per frame
            for ci in range((w - pane_x0 - 6) // 6):
                g = font[txt[ci]]
                x0 = pane_x0 + 4 + ci * 6
                blk = idx[y0:y0 + 7, x0:x0 + 5]
                blk[g[:blk.shape[0], :blk.shape[1]] > 0] = 1 if (li + t) % 5 else 3
            # static side bar with icons
            if li % 3 == 0:
                g = np.kron(font[(li // 3) % 26], np.ones((1, 1), np.uint8))
                for rep in range(max(1, (pane_x0 - 8) // 12)):
                    x0 = 4 + rep * 12
                    blk = idx[y0:y0 + 7, x0:x0 + 5]
                    blk[g[:blk.shape[0], :blk.shape[1]] > 0] = 4
        y = cols[idx, 0]
        u = cols[idx, 1]
        v = cols[idx, 2]
        yield _to420(y, u, v, bpc)


SOURCES["screen"] = screen_text

"""K8 film grain: oracle pinned against libdav1d (golden + live), CUDA kernel against both."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "filmgrain.npz")


def _golden_cases():
    z = np.load(GOLD)
    for bpc in (8, 10):
        for tv in range(1, 17):
            key = f"b{bpc}_tv{tv}"
            lens = z[key + "_tulens"]
            blob = z[key + "_tus"].tobytes()
            tus, pos = [], 0
            for n in lens:
                tus.append(blob[pos:pos + int(n)])
                pos += int(n)
            ins = [z[f"{key}_f0_in{p}"] for p in range(3)]
            outs = [z[f"{key}_f0_out{p}"] for p in range(3)]
            yield bpc, tv, tus, ins, outs


def test_oracle_matches_dav1d_golden(built):
    """oracle/filmgrain.c vs dav1d 1.5.3 (apply_grain=1) on all 16 libaom film-grain vectors."""
    import av1recon
    from oracle import oracle_lib
    seen = set()
    for bpc, tv, tus, ins, outs in _golden_cases():
        hdr = [h for h in av1recon.scan_headers(tus) if h.show_frame][0]
        fg = hdr.film_grain
        assert fg.apply_grain == 1
        seen.add((fg.ar_coeff_lag, fg.chroma_scaling_from_luma, fg.overlap_flag, fg.clip_to_restricted_range))
        got = oracle_lib.film_grain(fg, ins, bpc)
        for p in range(3):
            assert np.array_equal(got[p], outs[p]), f"bpc={bpc} vector={tv} plane={p}"
    assert len(seen) >= 4  # the vectors really exercise different parameter shapes


def test_oracle_matches_dav1d_live(built):
    """Same check on a freshly encoded stream with odd dimensions (ragged block/stripe tails)."""
    import av1recon
    from oracle import dav1d_ref, oracle_lib
    from tools import aomenc, sources
    for bpc, (w, h) in ((8, (150, 70)), (10, (98, 130))):
        fr = list(sources.noise_gradient(w, h, 2, bpc=bpc, seed=9))
        tus = aomenc.encode(fr, w, h, bpc=bpc, opts={"film-grain-test": "12", "cpu-used": "9"}, cfg={14: 0}, threads=1)
        hd = [x for x in av1recon.scan_headers(tus) if x.show_frame or x.show_existing_frame]
        d0 = dav1d_ref.decode(tus, apply_grain=0)
        d1 = dav1d_ref.decode(tus, apply_grain=1)
        assert len(hd) == len(d0) == len(d1) == 2
        for i in range(2):
            got = oracle_lib.film_grain(hd[i].film_grain, d0[i][4], bpc)
            for p in range(3):
                assert np.array_equal(got[p], d1[i][4][p])


@pytest.mark.gpu
def test_cuda_film_grain_golden(built):
    """CUDA K8 through the C ABI vs the dav1d golden planes and the C oracle: bit exact."""
    import torch
    import av1recon
    from oracle import oracle_lib
    from tests.gpu_util import gpu_film_grain
    for bpc, tv, tus, ins, outs in _golden_cases():
        fg = [h for h in av1recon.scan_headers(tus) if h.show_frame][0].film_grain
        got = gpu_film_grain(av1recon, torch, fg, ins, bpc)
        ref = oracle_lib.film_grain(fg, ins, bpc)
        for p in range(3):
            bad = np.argwhere(got[p] != outs[p])
            assert bad.size == 0, f"bpc={bpc} vector={tv} plane={p}: {len(bad)} px differ, first at {bad[0]}"
            assert np.array_equal(got[p], ref[p])


@pytest.mark.gpu
@pytest.mark.parametrize("bpc,w,h", [(8, 1920, 1080), (10, 3840, 2160), (10, 1000, 562), (8, 66, 34)])
def test_cuda_film_grain_full_size(built, bpc, w, h):
    """Full BASELINE sizes: CUDA vs the C oracle on seeded planes (oracle is pinned to dav1d above)."""
    import torch
    import av1recon
    from oracle import oracle_lib
    from tests.gpu_util import gpu_film_grain
    z = np.load(GOLD)
    rng = np.random.default_rng(w * 31 + h)
    mx = (1 << bpc) - 1
    planes = [rng.integers(0, mx + 1, size=(h, w)).astype(np.uint16),
              rng.integers(0, mx + 1, size=((h + 1) // 2, (w + 1) // 2)).astype(np.uint16),
              rng.integers(0, mx + 1, size=((h + 1) // 2, (w + 1) // 2)).astype(np.uint16)]
    for tv in (3, 1, 16, 7):
        key = f"b{bpc}_tv{tv}"
        lens = z[key + "_tulens"]
        blob = z[key + "_tus"].tobytes()
        tus, pos = [], 0
        for n in lens:
            tus.append(blob[pos:pos + int(n)])
            pos += int(n)
        fg = [x for x in av1recon.scan_headers(tus) if x.show_frame][0].film_grain
        got = gpu_film_grain(av1recon, torch, fg, planes, bpc)
        ref = oracle_lib.film_grain(fg, planes, bpc)
        for p in range(3):
            bad = np.argwhere(got[p] != ref[p])
            assert bad.size == 0, f"{w}x{h} bpc={bpc} vector={tv} plane={p}: {len(bad)} differ, first {bad[0]}"


@pytest.mark.gpu
def test_cuda_plane_checksum(built):
    import ctypes as C
    import torch
    import av1recon
    from tests.gpu_util import to_dev_plane
    l = av1recon.lib()
    rng = np.random.default_rng(5)
    for bpc, (w, h) in ((8, (333, 77)), (10, (1920, 1080))):
        a = rng.integers(0, 1 << bpc, size=(h, w)).astype(np.uint16)
        t, pitch = to_dev_plane(torch, a, bpc)
        out = torch.zeros(1, dtype=torch.int64, device="cuda")
        assert l.av1r_stage_plane_checksum(t.data_ptr(), pitch, w, h, bpc, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        host = np.ascontiguousarray(a.astype(np.uint8 if bpc == 8 else "<u2"))
        want = l.av1r_plane_checksum_host(host.ctypes.data, host.strides[0], w, h, bpc)
        assert (int(out.item()) & (2**64 - 1)) == want

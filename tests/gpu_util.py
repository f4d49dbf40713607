"""Helpers for -m gpu tests: device buffers via torch (plumbing only) + calls through the C ABI."""
import ctypes as C

import numpy as np


def to_dev_plane(torch, arr, bpc):
    """numpy plane -> (uint8 device tensor with 256-byte aligned pitch, pitch)."""
    h, w = arr.shape
    bps = 1 if bpc == 8 else 2
    pitch = (w * bps + 255) // 256 * 256
    host = np.zeros((h, pitch), dtype=np.uint8)
    host[:, : w * bps] = np.ascontiguousarray(arr.astype(np.uint8 if bpc == 8 else "<u2")).view(np.uint8).reshape(h, w * bps)
    return torch.from_numpy(host).cuda(), pitch


def from_dev_plane(t, w, h, bpc):
    bps = 1 if bpc == 8 else 2
    a = t.cpu().numpy()[:, : w * bps]
    return np.ascontiguousarray(a).view(np.uint8 if bpc == 8 else np.dtype("<u2")).reshape(h, w)


def gpu_film_grain(av1recon, torch, fg, planes, bpc, subx=1, suby=1, mono=0, mc_identity=0):
    l = av1recon.lib()
    h, w = planes[0].shape
    src = [to_dev_plane(torch, p, bpc) for p in planes]
    dst = [(torch.zeros_like(t), pitch) for t, pitch in src]
    scratch = torch.zeros(l.av1r_film_grain_scratch_bytes(), dtype=torch.uint8, device="cuda")
    sp = (C.c_void_p * 3)(*[t.data_ptr() for t, _ in src])
    dp = (C.c_void_p * 3)(*[t.data_ptr() for t, _ in dst])
    ss = (C.c_size_t * 3)(*[p for _, p in src])
    ds = (C.c_size_t * 3)(*[p for _, p in dst])
    torch.cuda.synchronize()
    rc = l.av1r_stage_film_grain(C.byref(fg), bpc, w, h, subx, suby, mono, mc_identity, sp, ss, dp, ds,
                                 scratch.data_ptr(), None)
    assert rc == 0, l.av1r_stage_last_error()
    torch.cuda.synchronize()
    return [from_dev_plane(dst[i][0], planes[i].shape[1], planes[i].shape[0], bpc) for i in range(3)]

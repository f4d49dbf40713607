"""Multi-GPU entry point inside the library (include/av1r.h av1r_pool_*, av1r_verify_batch): what the single-process daemon
(/root/reference/cmd/av1d/main.go:311-349) would call.  CPU tests cover the assignment rule; GPU tests compare a pool run with the
single-engine run frame by frame (digests) and check that one bad file does not spoil the others."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")
NAMES = ["inter_8b_sb128_tiles_640x360", "intra_8b_200x136", "inter_10b_alltools_208x144", "inter_8b_alltools_352x288",
         "intra_8b_lr_480x272", "inter_10b_grain_208x144"]


def test_library_assignment_matches_the_python_rule_and_is_balanced(built):
    """C++ longest-first assignment (what av1r_pool_verify_* uses) == av1recon.shard.assign (what the torchrun ranks of bench.py
    use), complete and balanced to within one item."""
    import av1recon
    from av1recon import shard
    weights = [1000 + 37 * ((f * 7 + s * 3) % 11) + (5000 if f == 3 else 0) for f in range(32) for s in range(2)]
    for nd in (1, 2, 3, 4, 8):
        a = av1recon.batch_assign(weights, nd)
        assert len(a) == len(weights) and set(a) <= set(range(nd))
        loads = [sum(w for w, d in zip(weights, a) if d == k) for k in range(nd)]
        assert max(loads) - min(loads) <= max(weights)
        parts = shard.assign([(i, w) for i, w in enumerate(weights)], nd)
        assert [sorted(i for i, d in enumerate(a) if d == k) for k in range(nd)] == parts


def test_pool_open_rejects_bad_arguments(built):
    import ctypes as C
    import av1recon
    l = av1recon.lib()
    l.av1r_pool_open.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert l.av1r_pool_open(None, 1, None, C.byref(h)) == -22
    devs = (C.c_int * 2)(0, 0)
    assert l.av1r_pool_open(devs, 2, None, C.byref(h)) == -22      # the same device twice
    assert l.av1r_pool_open(devs, 0, None, C.byref(h)) == -22


def _blobs():
    return [open(os.path.join(GOLD, n + ".ivf"), "rb").read() for n in NAMES]


@pytest.mark.gpu
def test_pool_on_one_gpu_equals_single_engine(built):
    import av1recon
    blobs = _blobs()
    want = []
    dec = av1recon.Decoder(streams=8, frames_in_flight=16, host_threads=4)
    for b in blobs:
        rc, rep, digs = dec.verify_buffer(b)
        assert rc == 0, rep.message
        want.append(digs)
    dec.close()
    pool = av1recon.Pool([0])
    rc, total, reps, digs = pool.verify_buffers(blobs, max_frames=64)
    assert rc == 0, total.message
    assert total.frames == sum(len(w) for w in want)
    for f in range(len(blobs)):
        assert reps[f].status == 0 and digs[f] == want[f], NAMES[f]
    # a second batch on the same pool (the daemon keeps it open), with one corrupt and one empty file in the middle
    bad = blobs[1][:len(blobs[1]) // 2]                    # truncated in the middle of a frame
    rc, total, reps, digs = pool.verify_buffers([blobs[0], bad, blobs[2], b"DKIF"], max_frames=64)
    assert rc != 0 and total.first_bad_frame == 1
    assert reps[0].status == 0 and digs[0] == want[0]
    assert reps[1].status != 0 and reps[3].status != 0
    assert reps[2].status == 0 and digs[2] == want[2]
    pool.close()


@pytest.mark.gpu
def test_verify_batch_on_files(built, tmp_path):
    import ctypes as C
    import av1recon
    from tools import mkvmux
    from tools.obuio import read_ivf
    paths = []
    for i, n in enumerate(NAMES[:3]):
        tus = read_ivf(os.path.join(GOLD, n + ".ivf"))
        hdr = av1recon.scan_headers(tus)[0]
        p = tmp_path / f"job{i}.av1-tmp.mkv"           # the daemon's temp name (daemon.go:86)
        p.write_bytes(mkvmux.mux(tus, hdr.width, hdr.height, ffmpeg_like=True))
        paths.append(str(p))
    paths.append(str(tmp_path / "missing.av1-tmp.mkv"))
    l = av1recon.lib()
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    reps = (av1recon.Report * n)()
    total = av1recon.Report()
    devs = (C.c_int * 1)(0)
    l.av1r_verify_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p, C.POINTER(av1recon.Report),
                                    C.POINTER(av1recon.Report)]
    rc = l.av1r_verify_batch(arr, n, devs, 1, None, reps, C.byref(total))
    assert rc == -2 and reps[3].status == -2, (rc, total.message)
    for i in range(3):
        assert reps[i].status == 0 and reps[i].frames > 0, reps[i].message


@pytest.mark.gpu
def test_two_gpus_in_one_process_equal_one_gpu(built):
    """Two engines on two devices inside this process (no torchrun, no NCCL): same digests as one device."""
    import torch
    import av1recon
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    blobs = _blobs()
    p1 = av1recon.Pool([0])
    rc1, t1, r1, d1 = p1.verify_buffers(blobs, max_frames=64)
    p1.close()
    p2 = av1recon.Pool([0, 1])
    rc2, t2, r2, d2 = p2.verify_buffers(blobs, max_frames=64)
    p2.close()
    assert rc1 == 0 and rc2 == 0, (t1.message, t2.message)
    assert d1 == d2 and t1.frames == t2.frames

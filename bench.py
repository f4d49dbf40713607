#!/usr/bin/env python
"""bench.py -- AV1 decode-verify throughput on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic input.  Workloads:
  filmgrain_4k10   K8 stage alone on 4K 10-bit frames (first kernel family that landed)
  c2_intra_1080p8  BASELINE configs[1]: 1080p 8-bit all-key-frame clip, full GPU reconstruction
The default is the most complete workload the engine currently decodes bit-exactly.
PyTorch is used only for device buffers / events / torch.distributed plumbing.
"""
import os as _os
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # frame-level overlap: one hardware queue per stream

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))

import numpy as np  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for i, nm in enumerate(names):
                    if r[2 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def golden_fg_params(tv=3, bpc=10):
    import av1recon
    z = np.load(os.path.join(ROOT, "tests", "golden", "filmgrain.npz"))
    key = f"b{bpc}_tv{tv}"
    lens = z[key + "_tulens"]
    blob = z[key + "_tus"].tobytes()
    tus, pos = [], 0
    for n in lens:
        tus.append(blob[pos:pos + int(n)])
        pos += int(n)
    return [h for h in av1recon.scan_headers(tus) if h.show_frame][0].film_grain


# ------------------------------------------------------------------------------------------
# workload: film grain stage on 4K10 frames
# ------------------------------------------------------------------------------------------
def run_filmgrain(args, torch, dist, rank, world, local):
    import av1recon
    l = av1recon.lib()
    w, h, bpc = 3840, 2160, 10
    nbuf = 8                      # 8 x 24.9 MB src + 8 x dst = 398 MB working set  > 126 MB L2
    fg = golden_fg_params(3, bpc)
    F = (w * h + 2 * (w // 2) * (h // 2)) * 2
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    pitch = [w * 2, w, w]         # bytes; already multiples of 256
    hs = [h, h // 2, h // 2]

    def mkframe(fill_random):
        planes = []
        for i in range(3):
            if fill_random:
                t = torch.randint(0, 1024, (hs[i], pitch[i] // 2), generator=g, device=dev, dtype=torch.int16)
            else:
                t = torch.zeros((hs[i], pitch[i] // 2), device=dev, dtype=torch.int16)
            planes.append(t)
        return planes

    srcs = [mkframe(True) for _ in range(nbuf)]
    dsts = [mkframe(False) for _ in range(nbuf)]
    scratch = [torch.zeros(l.av1r_film_grain_scratch_bytes(), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    cks = torch.zeros(nbuf * 3, dtype=torch.int64, device=dev)
    # host copies for the e2e leg (pinned)
    host_src = [[p.cpu().pin_memory() for p in srcs[0]]]
    host_cks = torch.zeros(3, dtype=torch.int64).pin_memory()

    def ptrs(planes):
        return (C.c_void_p * 3)(*[p.data_ptr() for p in planes])

    pit = (C.c_size_t * 3)(*pitch)
    stream = torch.cuda.current_stream(dev)
    sh = C.c_void_p(stream.cuda_stream)

    def one_frame(i):
        rc = l.av1r_stage_film_grain(C.byref(fg), bpc, w, h, 1, 1, 0, 0, ptrs(srcs[i]), pit, ptrs(dsts[i]), pit,
                                     scratch[i].data_ptr(), sh)
        if rc:
            raise RuntimeError(l.av1r_stage_last_error())

    def step():
        for i in range(nbuf):
            one_frame(i)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frames = args.steps * nbuf * world
    value = frames / (ms / 1e3)
    # kernel-only timing of the dominant kernel (fg_apply): time apply+prepare per frame; prepare is a
    # single-CTA kernel overlapping nothing here, so report the per-frame pair and the apply share from ncu.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for i in range(nbuf):
        one_frame(i)
    ev[1].record(stream)
    torch.cuda.synchronize()
    per_frame_ms = ev[0].elapsed_time(ev[1]) / nbuf
    peak, peak_src = measured_peaks()
    achieved = 2 * F / (per_frame_ms / 1e3) / 1e9
    # e2e: host planes in pinned memory -> H2D -> film grain -> checksum -> D2H of the 3 digests
    def e2e_step():
        for i in range(nbuf):
            for p in range(3):
                srcs[i][p].copy_(host_src[0][p], non_blocking=True)
            one_frame(i)
            for p in range(3):
                l.av1r_stage_plane_checksum(dsts[i][p].data_ptr(), pitch[p], w if p == 0 else w // 2, hs[p], bpc,
                                            cks[i * 3 + p:].data_ptr(), sh)
            host_cks.copy_(cks[i * 3:i * 3 + 3], non_blocking=True)
    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ksteps = max(1, args.steps // 2)
    for _ in range(ksteps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = ksteps * nbuf * world / e2e_s
    out = {
        "metric": "AV1 decode-verify frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": "filmgrain_4k10: K8 film-grain stage only, 3840x2160 10-bit 4:2:0, 8 frames/step, "
                               "libaom film-grain test vector 3; working set 398 MB > L2 (no flush needed)",
                   "frames_per_step": nbuf, "parallelism": f"replicas{world}"},
        "gpu_launches": 2 * nbuf * args.steps,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": F * nbuf, "d2h_bytes_per_step": 24 * nbuf},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "fg_apply_kernel(+fg_prepare)",
                     "algorithmic_bytes_per_launch": 2 * F},
        "clocks": clocks,
    }
    return out


def cpu_baseline_filmgrain():
    """oracle (port) film grain on host cores: bounded sample of 4 4K10 frames, 1 thread."""
    from oracle import oracle_lib
    w, h, bpc = 3840, 2160, 10
    fg = golden_fg_params(3, bpc)
    rng = np.random.default_rng(0)
    planes = [rng.integers(0, 1024, size=(h, w)).astype(np.uint16), rng.integers(0, 1024, size=(h // 2, w // 2)).astype(np.uint16),
              rng.integers(0, 1024, size=(h // 2, w // 2)).astype(np.uint16)]
    oracle_lib.film_grain(fg, planes, bpc)
    n = 4
    t0 = time.perf_counter()
    for _ in range(n):
        oracle_lib.film_grain(fg, planes, bpc)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{n} frames 3840x2160 10-bit through oracle/filmgrain.c (scalar C, 1 thread)"}


# ------------------------------------------------------------------------------------------
# workload: BASELINE configs[1] -- 1080p 8-bit intra-only clip, full GPU reconstruction
# ------------------------------------------------------------------------------------------
CLIP_DESC = {
    "c1": ("c1_1080p8: BASELINE configs[0] -- 1920x1080 8-bit 4:2:0 Main profile, 60 frames, testsrc2-like source, libaom 3.13.1 cq 32, "
           "lag_in_frames 19 (hidden ARFs + show_existing_frame), kf_max_dist 30 (2 closed GOPs), 2 tile columns, all default tools"),
    "c3": ("c3_4k10_inter: BASELINE configs[2] -- 3840x2160 10-bit 4:2:0, 60 frames, pan/zoom/rotation texture with moving patches, "
           "libaom cq 32, compound + OBMC + warped/global motion + loop restoration, 4x2 tiles, lag 19, kf_max_dist 30"),
    "c4": ("c4_4k10_grain: BASELINE configs[3] -- as c3 on a noise-heavy source with film grain synthesis (libaom film-grain-test 5)"),
    "c3_small": "c3_small: 960x544 10-bit inter clip (smoke-size version of c3)",
}
C2_DESC = ("c2_intra_1080p8: BASELINE configs[1] -- 1920x1080 8-bit 4:2:0, 60 frames, every frame KEY (libaom 3.13.1, cq 32, "
           "CDEF on, LR off, synthetic pan/zoom texture); step = one pass over the 60-frame clip; "
           "value = device path from HBM-resident work-lists (sequential host symbol parse reported separately as host_parse_ms); "
           "per-step working set (work-lists + frame buffers of 60 frames) exceeds the 126 MB L2, no flush needed")


def c2_clip(name="c2"):
    from tools.make_streams import get_clip
    return get_clip(name, verbose=True)


STEP_NOTE = ("; step = one pass over the clip; value = device path from HBM-resident work-lists (sequential host symbol parse reported "
             "separately as host_parse_ms); per-step working set (work-lists + frame buffers in flight) exceeds the 126 MB L2, no flush needed")


def run_c2(args, torch, dist, rank, world, local, name="c2"):
    import av1recon
    tus = c2_clip(name)
    torch.cuda.set_device(local)
    dec = av1recon.Decoder(device=local, streams=32, frames_in_flight=64)
    clip = av1recon.Clip(dec, tus)
    info = clip.info
    nfr = int(info.frames_shown)
    # reference digests: first replay
    ms0, cks0 = clip.decode()
    for _ in range(args.warmup):
        clip.decode()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        ms, cks = clip.decode()
        total_ms += ms
        if cks != cks0:
            raise RuntimeError("replay produced different digests: non-deterministic reconstruction")
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = nfr * args.steps * world / (total_ms / 1e3)
    # per-stage device time (serialised replay, CUDA events) -> live roofline of the dominant kernel
    prof = clip.profile()
    F = int(info.frame_bytes)
    A = int(info.coded_samples)
    ntok = int(info.coef_tokens)
    nrec = int(info.tx_blocks)
    nf = int(info.frames_decoded)
    bps = 1 if info.bit_depth == 8 else 2
    intra_s, inter_s, ref_s = int(info.intra_samples), int(info.inter_samples), int(info.inter_ref_samples)
    stage_bytes = {
        "itx": 4 * ntok + 32 * nrec + 2 * A,                 # C (tokens) + records + residual write
        "intra": bps * intra_s + 2 * A + 32 * nrec,           # intra samples written + residual read + records
        "inter": bps * (ref_s + inter_s) + 40 * int(info.inter_blocks),   # Rbar * F_inter read + F_inter written + records
        "deblock": 2 * F * int(info.deblock_frames),
        "cdef": 2 * F * int(info.cdef_frames),
        "lr": int(2.0625 * F) * int(info.lr_frames),
        "grain": 2 * F * int(info.grain_frames),
        "digest": F * nfr,
    }
    # (K6 super-resolution has no bench workload: the BASELINE configs do not use it; parity only, tests/golden/*superres*)
    stages = {}
    for k, (ms, launches) in prof.items():
        if launches:
            ent = {"ms_per_clip": ms, "launches": launches}
            if k in stage_bytes and ms > 0:
                ent["algorithmic_gbs"] = stage_bytes[k] / (ms / 1e3) / 1e9
            stages[k] = ent
    dom = max((k for k in stages if k in stage_bytes), key=lambda k: stages[k]["ms_per_clip"])
    peak, peak_src = measured_peaks()
    dom_ms = stages[dom]["ms_per_clip"] / stages[dom]["launches"]
    dom_bytes = stage_bytes[dom] / stages[dom]["launches"]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    # e2e: the call a user makes -- av1r_verify_buffer on the HOST container bytes: demux, GOP-segment-parallel
    # host symbol parse (all host cores), H2D of the work-lists, reconstruction kernels, D2H of the 24-byte
    # plane digests of every frame.  Wall clock.
    from tools.make_streams import clip_path
    blob = open(clip_path(name), "rb").read()
    vdec = av1recon.Decoder(device=local, streams=16, frames_in_flight=32,   # the daemon keeps one engine open
                            host_threads=max(1, (os.cpu_count() or 1) // world))
    vdec.verify_buffer(blob)                                                  # warm-up (allocations, first-touch)
    best = None
    for _ in range(5):   # wall-clock of a host-bound path on a shared box: best of 5
        t0 = time.perf_counter()
        rc, rep, digs = vdec.verify_buffer(blob)
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(f"av1r_verify_buffer failed: {rep.message}")
        if digs != cks0:
            raise RuntimeError("e2e digests differ from replay digests")
        best = dt if best is None else min(best, dt)
    e2e_s = best
    parse_ms = rep.host_parse_ms
    vdec.close()
    # single-threaded streaming API (av1r_submit_tu per temporal unit), for reference
    dec2 = av1recon.Decoder(device=local, streams=16, frames_in_flight=32)
    t0 = time.perf_counter()
    for i, tu in enumerate(tus):
        dec2.submit(tu, i)
    dec2.flush()
    e2e_1t = nfr / (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = nfr * world / e2e_s
    launches_per_step = sum(v["launches"] for v in stages.values())
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload (or null)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(WORKLOAD_NAMES.get(name, name), {}).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        traffic = None
    out = {
        "metric": "AV1 decode-verify frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if info.bit_depth == 8 else "u16", "data": "synthetic",
        "config": {"workload": C2_DESC if name == "c2" else CLIP_DESC[name] + STEP_NOTE, "frames_per_step": nfr, "parallelism": f"replicas{world} (independent clips per GPU, no collective)",
                   "streams": 32, "frames_in_flight": 64},
        "gpu_launches": launches_per_step * args.steps,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(info.worklist_bytes), "d2h_bytes_per_step": 24 * nfr,
                "host_parse_ms_per_step": parse_ms, "host_threads": os.cpu_count(), "single_thread_submit_tu_fps": e2e_1t,
                "note": "av1r_verify_buffer: key-frame-delimited GOP segments parsed on all host cores; host_parse_ms is the summed sequential symbol-parse time (north_star: reported separately)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": dom, "algorithmic_bytes_per_launch": dom_bytes,
                     "avg_launch_ms": dom_ms, "stages": stages,
                     "pipeline_algorithmic_gbs": sum(stage_bytes[k] for k in stages if k in stage_bytes) / (total_ms / args.steps / 1e3) / 1e9},
        "host_parse_ms_per_frame": float(info.host_parse_ms) / max(1, nf),
        "clip": {"bytes": sum(len(t) for t in tus), "frames": nfr, "coded_sample_fraction": A / max(1, nf * (F // bps)),
                 "coef_tokens_per_frame": ntok / max(1, nf), "tx_blocks_per_frame": nrec / max(1, nf),
                 "frames_decoded": nf, "inter_sample_fraction": inter_s / max(1, nf * (F // bps)),
                 "mean_refs_per_inter_sample": ref_s / max(1, inter_s),
                 "tools": {k: int(v) for k, v in zip(av1recon.TOOL_NAMES, info.tool_hist) if v}},
        "clocks": clocks,
    }
    clip.free()
    dec.close()
    dec2.close()
    return out


def cpu_baseline_c2(name="c2"):
    """libdav1d 1.5.3 (the decoder inside the reference's FFmpeg build) on the host cores, same clip."""
    from oracle import dav1d_ref
    tus = c2_clip(name)
    ncpu = os.cpu_count() or 1
    dav1d_ref.decode(tus[:8], n_threads=ncpu, keep=False)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        out = dav1d_ref.decode(tus, n_threads=ncpu, keep=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter()
    dav1d_ref.decode(tus[:20], n_threads=1, keep=False)
    dt1 = time.perf_counter() - t0
    return {"value": len(out) / best, "unit": "frames/s", "cores": ncpu, "kind": "reference", "ms_per_step": best * 1e3,
            "dtype": "u8" if name in ("c1", "c2", "c2_small") else "u16",
            "sample": f"libdav1d {dav1d_ref.version()} driven directly (no ffmpeg binary in the image), n_threads={ncpu}, whole {len(tus)}-TU clip "
                      f"preloaded in RAM, best of 3, no MD5; single-thread figure {20 / dt1:.1f} frames/s on 20 frames"}


# ------------------------------------------------------------------------------------------
# workload: BASELINE configs[4] -- batch of 32 4K 10-bit files, GOP-segment-sharded across the GPUs (no collective)
# ------------------------------------------------------------------------------------------
C5_DESC = ("c5_batch_4k10: BASELINE configs[4] -- batch of 32 synthetic 3840x2160 10-bit files (pan/zoom texture, seeds 100..131, 16 frames "
           "each, kf_max_dist 4 => 4 closed GOP segments per file => 128 independent work items), items assigned to the ranks "
           "longest-first by coded bytes, one engine per GPU, no collective; step = one pass over the whole batch (strong scaling: the "
           "batch is fixed, per-rank share shrinks with N); value = device path from HBM-resident work-lists")


def c5_items(nfiles=32):
    """-> (items [(key, weight)], tus_of {key: [tu bytes]}, (w, h))"""
    import av1recon
    from av1recon import shard
    from tools.make_streams import get_clip
    items, tus_of = [], {}
    for f in range(nfiles):
        tus = get_clip(f"c5_{f:02d}", verbose=True)
        for s_idx, (a, b) in enumerate(shard.split_segments(tus, av1recon.scan_headers(tus))):
            key = (f, s_idx)
            tus_of[key] = tus[a:b]
            items.append((key, sum(len(t) for t in tus[a:b])))
    return items, tus_of, (3840, 2160)


def run_c5(args, torch, dist, rank, world, local):
    import av1recon
    from av1recon import shard
    t_start = time.perf_counter()

    def note(msg):   # progress on stderr: this workload holds 512 4K frames and takes minutes end to end
        print(f"[c5 rank {rank} +{time.perf_counter() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)
    items, tus_of, (w, h) = c5_items(int(os.environ.get("AV1R_C5_FILES", "32")))
    note(f"{len(items)} GOP segments scanned")
    mine = shard.assign(items, world)[rank]
    my_tus = [t for k in mine for t in tus_of[k]]
    torch.cuda.set_device(local)
    dec = av1recon.Decoder(device=local, streams=32, frames_in_flight=64)
    clip = av1recon.Clip(dec, my_tus)
    info = clip.info
    nfr_local = int(info.frames_shown)
    note(f"clip loaded: {nfr_local} frames")
    ms0, cks0 = clip.decode()
    note(f"first decode {ms0:.1f} ms")
    for _ in range(args.warmup):
        clip.decode()
    note("warm-up done")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        ms, cks = clip.decode()
        total_ms += ms
        if cks != cks0:
            raise RuntimeError("replay produced different digests: non-deterministic reconstruction")
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    nfr = nfr_local
    if world > 1:
        t = torch.tensor([total_ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        n = torch.tensor([nfr_local], device=f"cuda:{local}")
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        nfr = int(n.item())
    value = nfr * args.steps / (total_ms / 1e3)
    note(f"timed steps done: {total_ms / args.steps:.1f} ms per step")
    prof = clip.profile()
    note("stage profile done")
    stages = {k: {"ms_per_step": ms, "launches": n} for k, (ms, n) in prof.items() if n}
    # e2e: the rank's share of the batch as one container through av1r_ctx_verify_buffer (host parse of the segments on this
    # rank's share of the host cores, H2D, kernels, D2H of the digests)
    blob = shard.ivf_bytes(my_tus, w, h)
    host_threads = max(1, (os.cpu_count() or 1) // world)
    vdec = av1recon.Decoder(device=local, streams=16, frames_in_flight=32, host_threads=host_threads)
    vdec.verify_buffer(blob)
    note("e2e warm-up done")
    best = None
    for _ in range(2):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rc, rep, digs = vdec.verify_buffer(blob)
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(f"av1r_verify_buffer failed: {rep.message}")
        if digs != cks0:
            raise RuntimeError("e2e digests differ from replay digests")
        best = dt if best is None else min(best, dt)
    e2e_s = best
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    vdec.close()
    peak, peak_src = measured_peaks()
    F = int(info.frame_bytes)
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    out = {
        "metric": "AV1 decode-verify frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": C5_DESC, "frames_per_step": nfr, "items": len(items), "items_this_rank": len(mine),
                   "parallelism": f"gop-segment sharding over {world} GPU(s), no collective", "streams": 32, "frames_in_flight": 64},
        "gpu_launches": sum(v["launches"] for v in stages.values()) * args.steps,
        "e2e": {"value": nfr / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(info.worklist_bytes), "d2h_bytes_per_step": 24 * nfr_local,
                "host_threads_per_rank": host_threads, "host_parse_ms_per_step_rank0": rep.host_parse_ms},
        "roofline": {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                     "kernel": dom, "stages": stages, "note": "per-stage algorithmic GB/s are reported by the single-clip workloads (c3_4k10_inter)"},
        "clip": {"frames": nfr, "frame_bytes": F, "tools": {k: int(v) for k, v in zip(av1recon.TOOL_NAMES, info.tool_hist) if v}},
        "clocks": clocks,
    }
    clip.free()
    dec.close()
    return out


def cpu_baseline_c5():
    """libdav1d on all host cores over a bounded sample of the batch (8 of the 32 files, one after the other)."""
    from oracle import dav1d_ref
    from tools.make_streams import get_clip
    ncpu = os.cpu_count() or 1
    files = [get_clip(f"c5_{f:02d}") for f in range(8)]
    dav1d_ref.decode(files[0], n_threads=ncpu, keep=False)
    t0 = time.perf_counter()
    n = 0
    for tus in files:
        n += len(dav1d_ref.decode(tus, n_threads=ncpu, keep=False))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": ncpu, "kind": "reference",
            "sample": f"libdav1d {dav1d_ref.version()} n_threads={ncpu}, files c5_00..c5_07 of the batch (8 x 16 frames 4K10) decoded back to back, preloaded in RAM, no MD5"}


# clip key (tools/make_streams.py) -> bench workload name (profiles/ncu_traffic.json is keyed by the latter)
WORKLOAD_NAMES = {"c2": "c2_intra_1080p8", "c1": "c1_1080p8", "c3": "c3_4k10_inter", "c4": "c4_4k10_grain"}


def _clip_workload(name):
    return (lambda *a: run_c2(*a, name=name)), (lambda: cpu_baseline_c2(name))


WORKLOADS = {"filmgrain_4k10": (run_filmgrain, cpu_baseline_filmgrain), "c2_intra_1080p8": (run_c2, cpu_baseline_c2),
             "c1_1080p8": _clip_workload("c1"), "c3_4k10_inter": _clip_workload("c3"), "c4_4k10_grain": _clip_workload("c4"),
             "c3_small": _clip_workload("c3_small"), "c5_batch_4k10": (run_c5, cpu_baseline_c5)}
DEFAULT_WORKLOAD = "c2_intra_1080p8"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()
    run, cpu_base = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_base()
        line = {"impl": "reference", "metric": "AV1 decode-verify frames/s", "value": cb["value"], "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": {"workload": args.workload},
                "ms_per_step": cb.get("ms_per_step"), "dtype": cb.get("dtype"),
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # libraries (NCCL prints its version banner) must not write to stdout: the contract is ONE JSON line there
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 2
    out = run(args, torch, dist, rank, world, local)
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_base()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

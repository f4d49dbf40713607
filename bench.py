#!/usr/bin/env python
"""bench.py -- AV1 decode-verify throughput on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one clip (or one batch of files).
  --gpus 1 (default): headline = c3_4k10_inter (BASELINE configs[2]: 3840x2160 10-bit, compound / OBMC / warped motion / loop
                      restoration); the same line carries `per_config`: c1, c2, c4 and the 32-file batch c5, each with value, e2e,
                      cpu_baseline and the dominant kernel's roofline fraction.
  --gpus N > 1:       c5_batch_4k10 (BASELINE configs[4]): the batch's GOP segments sharded longest-first over the N ranks (one
                      engine per GPU, no collective on the data path), strong scaling.
  --workload NAME     one of WORKLOADS (single-clip line without per_config).
  --impl reference    libdav1d (the decoder inside the reference's FFmpeg build) on the host cores, same workload / steps.
`value` is timed from the first work-list H2D enqueue to the last kernel; `value_hbm_resident` is the same pass with
the work-lists already in HBM.  bench.py refuses a clip whose tool histogram lacks the tools of the BASELINE config it stands for.
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
# clip workloads: BASELINE configs[0..3] as single clips, configs[4] as a batch of 32 files
# ------------------------------------------------------------------------------------------
CLIP_DESC = {
    "c1": ("c1_1080p8: BASELINE configs[0] -- 1920x1080 8-bit 4:2:0 Main profile, 60 frames, testsrc2-like source, libaom 3.13.1 cq 32, "
           "lag_in_frames 19 (hidden ARFs + show_existing_frame), kf_max_dist 30 (2 closed GOPs), 2 tile columns, all default tools"),
    "c2": ("c2_intra_1080p8: BASELINE configs[1] -- 1920x1080 8-bit 4:2:0, 60 frames, every frame KEY (libaom 3.13.1, cq 32, "
           "CDEF on, LR off, synthetic pan/zoom texture)"),
    "c3": ("c3_4k10_inter: BASELINE configs[2] -- 3840x2160 10-bit 4:2:0, 60 frames, fine texture under pan/zoom with independently moving, "
           "crossing and fading foreground objects, libaom 3.13.1 cpu-used 1 cq 32, compound (average / distance / wedge / difference-weighted), "
           "inter-intra, OBMC, local + global warped motion, loop restoration, 4x2 tiles, lag 19, two closed GOPs of 30"),
    "c4": ("c4_4k10_grain: BASELINE configs[3] -- 3840x2160 10-bit noise-heavy source with film grain synthesis (libaom film-grain-test 5), "
           "60 frames, 4x2 tiles, lag 19, kf_max_dist 30"),
    "c3_small": "c3_small: 960x544 10-bit inter clip (smoke-size version of c3)",
    "c2_small": "c2_small: 640x360 8-bit all-key clip (smoke-size version of c2)",
}
STEP_NOTE = "; step = one pass over the whole clip"
# The two 4K configs name no clip length (BASELINE configs[2], [3]); a file the daemon verifies holds hundreds of GOPs, and GOP segments are
# the unit of parallelism of this path, so the 4K clips are benchmarked as the encoded 60-frame sequence played CLIP_REPEAT
# times back to back (every repetition starts with its key frame and sequence header: 2 x CLIP_REPEAT closed GOPs).  Both arms -- this
# engine and libdav1d -- decode the same repeated stream.  configs[0] says 60 frames and stays at 60; c2 is 60 independent key frames.
CLIP_REPEAT = {"c3": int(os.environ.get("AV1R_BENCH_REPEAT", "4")), "c4": int(os.environ.get("AV1R_BENCH_REPEAT", "4"))}


def bench_clip(name):
    """Temporal units of a bench workload: the cached clip, repeated for the 4K configs (see CLIP_REPEAT)."""
    return get_clip(name) * CLIP_REPEAT.get(name, 1)

ENGINE_STREAMS = int(os.environ.get("AV1R_BENCH_STREAMS", "32"))
ENGINE_FRAMES_IN_FLIGHT = int(os.environ.get("AV1R_BENCH_FIF", "64"))
TIMING_NOTE = ("per frame one H2D copy of its work-lists, then the reconstruction kernels, frames pipelined over 32 streams; the K timed steps are "
               "enqueued back to back (av1r_clip_decode_passes: no drain between steps, the first key frames of step k+1 overlap the tail of step "
               "k as consecutive GOPs of a long file do) and bracketed by ONE pair of CUDA events from the first H2D enqueue of step 1 to the last "
               "kernel of step K, max over ranks, every step's digests checked against the first pass; value_fenced_steps is the same with every "
               "step fenced and drained on its own (rounds 1-2 definition); the sequential host symbol parse is reported separately as host_parse_ms")

# tools / stages a clip must really contain to stand for its BASELINE config (block counts from the host parser, frames per stage)
REQUIRED = {
    "c1": dict(tools=["inter_blocks", "compound_avg"], stages=["cdef_frames"]),
    "c2": dict(tools=[], stages=["cdef_frames", "deblock_frames"], forbid=["inter_blocks"]),
    "c3": dict(tools=["inter_blocks", "compound_avg", "compound_dist", "compound_wedge", "compound_diffwtd", "interintra", "obmc", "local_warp",
                      "global_warp"], stages=["lr_frames", "cdef_frames", "deblock_frames"]),
    "c4": dict(tools=["inter_blocks", "compound_avg"], stages=["grain_frames", "cdef_frames"]),
    "c5": dict(tools=["inter_blocks", "compound_avg", "compound_wedge", "compound_diffwtd", "interintra", "obmc", "local_warp", "global_warp"],
               stages=["lr_frames", "cdef_frames", "deblock_frames"]),
}


class ClipLacksTools(RuntimeError):
    pass


def check_clip_tools(key, info, hist):
    """Refuse to benchmark a clip that does not exercise what its config names (the 4K10 clip held no OBMC / masked compound
    and no loop restoration)."""
    req = REQUIRED.get(key)
    if not req:
        return
    missing = [t for t in req["tools"] if not hist.get(t)] + [s for s in req["stages"] if not int(getattr(info, s))]
    present = [t for t in req.get("forbid", []) if hist.get(t)]
    if missing or present:
        raise ClipLacksTools(f"clip {key} does not stand for its BASELINE config: missing {missing}, unexpected {present}; "
                             f"regenerate it with `python -m tools.make_streams {key}`")


def get_clip(name):
    from tools.make_streams import get_clip as gc
    return gc(name, verbose=True)


def stage_model_bytes(info):
    """Algorithmic bytes per stage summed over the clip (with the counts the parser collected)."""
    F = int(info.frame_bytes)
    A = int(info.coded_samples)
    bps = 1 if info.bit_depth == 8 else 2
    return {
        "itx": 4 * int(info.coef_tokens) + 32 * int(info.tx_blocks) + 2 * A,       # C (tokens) + records + residual write
        # the intra wavefront kernel is two kernels: the whole-frame build on frames without inter blocks (key frames: a dependency
        # chain over the whole picture) and the scattered-unit build on inter frames; samples written + residual read + records
        "intra_frame": bps * int(info.intra_frame_samples) + 2 * int(info.intra_frame_coded_samples) + 32 * int(info.intra_frame_tx_blocks),
        "intra": bps * (int(info.intra_samples) - int(info.intra_frame_samples)) + 2 * (A - int(info.intra_frame_coded_samples))
                 + 32 * (int(info.tx_blocks) - int(info.intra_frame_tx_blocks)),
        "inter": bps * (int(info.inter_ref_samples) + int(info.inter_samples)) + 40 * int(info.inter_blocks),   # Rbar*F_inter + F_inter
        "deblock": 2 * F * int(info.deblock_frames),
        "cdef": 2 * F * int(info.cdef_frames),
        "lr": int(2.0625 * F) * int(info.lr_frames),
        "grain": 2 * F * int(info.grain_frames),
        "digest": F * int(info.frames_shown),
    }


def dist_max(torch, dist, world, local, x):
    if world <= 1:
        return x
    t = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_sum(torch, dist, world, local, x):
    if world <= 1:
        return x
    t = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def measure_clip(key, tus, blob, args, torch, dist, rank, world, local, steps, warmup, desc, sample_clocks=True, e2e_reps=5, gate_key=None):
    """Device path + e2e + live roofline of one clip (or of this rank's share of a batch, given as one list of temporal units)."""
    import av1recon
    torch.cuda.set_device(local)
    dec = av1recon.Decoder(device=local, streams=ENGINE_STREAMS, frames_in_flight=ENGINE_FRAMES_IN_FLIGHT)
    clip = av1recon.Clip(dec, tus)
    info = clip.info
    hist = av1recon.tool_hist(info)
    if world == 1:
        check_clip_tools(gate_key or key, info, hist)
    nfr_local = int(info.frames_shown)
    ms0, cks0 = clip.decode()                       # first pass: allocations; its digests are the reference for every later pass
    for _ in range(warmup):
        clip.decode()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0 and sample_clocks:
        sampler.start()
    # the K timed steps: enqueued back to back inside one fenced region (the library fails the call if any step's digests differ)
    total_ms, cks = clip.decode_passes(steps)
    if cks != cks0:
        raise RuntimeError("replay produced different digests: non-deterministic reconstruction")
    torch.cuda.synchronize()
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    total_ms = dist_max(torch, dist, world, local, total_ms)
    nfr = int(round(dist_sum(torch, dist, world, local, nfr_local)))
    value = nfr * steps / (total_ms / 1e3)
    # the same steps, each fenced and drained on its own (what rounds 1-2 reported as `value`)
    fenced_ms = 0.0
    nfen = max(2, min(10, steps))
    for _ in range(nfen):
        ms, cks = clip.decode()
        fenced_ms += ms
        if cks != cks0:
            raise RuntimeError("replay produced different digests: non-deterministic reconstruction")
    fenced_ms = dist_max(torch, dist, world, local, fenced_ms)
    value_fenced = nfr * nfen / (fenced_ms / 1e3)
    # the same pass with the work-lists already resident in HBM (kernel-only figure)
    clip.set_resident(True)
    clip.decode()
    nres = max(2, min(10, steps))
    res_ms, cks = clip.decode_passes(nres)
    if cks != cks0:
        raise RuntimeError("resident replay produced different digests")
    res_ms = dist_max(torch, dist, world, local, res_ms)
    value_resident = nfr * nres / (res_ms / 1e3)
    clip.set_resident(False)
    # per-stage device time (serialised replay, CUDA events on the launching stream) -> live roofline of the dominant kernel
    prof = clip.profile()
    sb = stage_model_bytes(info)
    stages = {}
    for k, (ms, launches) in prof.items():
        if launches:
            ent = {"ms_per_step": ms, "launches": launches}
            if k in sb and ms > 0:
                ent["algorithmic_gbs"] = sb[k] / (ms / 1e3) / 1e9
            stages[k] = ent
    peak, peak_src = measured_peaks()
    dom = max((k for k in stages if k in sb), key=lambda k: stages[k]["ms_per_step"])
    dom_ms = stages[dom]["ms_per_step"] / stages[dom]["launches"]
    dom_bytes = sb[dom] / stages[dom]["launches"]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9
    for k, ent in stages.items():
        if "algorithmic_gbs" in ent:
            ent["frac_of_peak"] = ent["algorithmic_gbs"] / peak
    launches_per_step = sum(v["launches"] for k, v in stages.items() if k != "h2d")
    # e2e: the call a user makes -- av1r_ctx_verify_buffer on the HOST container bytes with an engine that stays open (the daemon
    # keeps one): demux, GOP-segment-parallel host symbol parse, H2D of the work-lists, kernels, D2H of the 24-byte plane digests.
    host_threads = max(1, (os.cpu_count() or 1) // world)
    vdec = av1recon.Decoder(device=local, streams=16, frames_in_flight=32, host_threads=host_threads)
    vdec.verify_buffer(blob)                        # warm-up (allocations, first touch)
    best, parse_ms = None, 0.0
    for _ in range(e2e_reps):                       # wall clock of a host-bound path on a shared box: best of e2e_reps
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rc, rep, digs = vdec.verify_buffer(blob)
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(f"av1r_ctx_verify_buffer failed: {rep.message}")
        if digs != cks0:
            raise RuntimeError("e2e digests differ from replay digests")
        dt = dist_max(torch, dist, world, local, dt)
        if best is None or dt < best:
            best, parse_ms = dt, rep.host_parse_ms
    vdec.close()
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        per = tj.get(WORKLOAD_NAMES.get(key, key), {})
        traffic = (per.get(dom) or per.get("intra" if dom == "intra_frame" else dom, {})).get("dram_bytes_per_launch")
    except Exception:
        traffic = None
    nf = int(info.frames_decoded)
    F = int(info.frame_bytes)
    bps = 1 if info.bit_depth == 8 else 2
    out = {
        "metric": "AV1 decode-verify frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8" if info.bit_depth == 8 else "u16", "data": "synthetic",
        "config": workload_config(WORKLOAD_NAMES.get(key, key), world),
        "engine": {"streams": ENGINE_STREAMS, "frames_in_flight": ENGINE_FRAMES_IN_FLIGHT, "frames_per_step_all_ranks": nfr, "timing": TIMING_NOTE},
        "value_hbm_resident": value_resident, "value_fenced_steps": value_fenced,
        "gpu_launches": launches_per_step * steps,
        "e2e": {"value": nfr / best, "unit": "frames/s", "h2d_bytes_per_step": int(info.worklist_bytes) + len(blob) * 0,
                "d2h_bytes_per_step": 24 * nfr_local, "host_input_bytes_per_step": len(blob),
                "host_parse_ms_per_step": parse_ms, "host_threads": host_threads,
                "note": "av1r_ctx_verify_buffer on the container bytes in host memory: demux, GOP-segment- and tile-parallel host symbol parse, "
                        "one H2D copy of the work-lists per frame from pinned staging, kernels, D2H of the plane digests; wall clock, best of "
                        f"{e2e_reps}; host_parse_ms is the summed sequential symbol-parse time (north_star: reported separately)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": dom, "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                     "stages": stages,
                     "pipeline_algorithmic_gbs": sum(sb[k] for k in stages if k in sb) / (total_ms / steps / 1e3) / 1e9},
        "host_parse_ms_per_frame": float(info.host_parse_ms) / max(1, nf),
        "clip": {"bytes": sum(len(t) for t in tus), "frames": nfr_local, "frames_decoded": nf, "frame_bytes": F,
                 "coded_sample_fraction": int(info.coded_samples) / max(1, nf * (F // bps)),
                 "coef_tokens_per_frame": int(info.coef_tokens) / max(1, nf), "tx_blocks_per_frame": int(info.tx_blocks) / max(1, nf),
                 "inter_sample_fraction": int(info.inter_samples) / max(1, nf * (F // bps)),
                 "mean_refs_per_inter_sample": int(info.inter_ref_samples) / max(1, int(info.inter_samples)),
                 "lr_frames": int(info.lr_frames), "cdef_frames": int(info.cdef_frames), "deblock_frames": int(info.deblock_frames),
                 "grain_frames": int(info.grain_frames), "tools": hist},
        "clocks": clocks,
    }
    clip.free()
    dec.close()
    return out


def run_clip(args, torch, dist, rank, world, local, name, steps=None, warmup=None, sample_clocks=True):
    """Single clip; with world > 1 every rank decodes its own replica (weak scaling, `replicasN`)."""
    from tools.make_streams import clip_path
    from tools.obuio import ivf_header
    from av1recon import shard
    tus = bench_clip(name)
    if CLIP_REPEAT.get(name, 1) > 1:
        hdr = ivf_header(clip_path(name))
        blob = shard.ivf_bytes(tus, hdr["w"], hdr["h"])
    else:
        blob = open(clip_path(name), "rb").read()
    out = measure_clip(name, tus, blob, args, torch, dist, rank, world, local, steps or args.steps, warmup or args.warmup,
                       CLIP_DESC[name], sample_clocks=sample_clocks)
    return out


def dav1d_pass(tus, n_threads):
    from oracle import dav1d_ref
    t0 = time.perf_counter()
    n = len(dav1d_ref.decode(tus, n_threads=n_threads, keep=False))
    return n, time.perf_counter() - t0


def cpu_baseline_clip(name, passes=3):
    """libdav1d 1.5.3 (the decoder inside the reference's FFmpeg build) on the host cores, same clip."""
    from oracle import dav1d_ref
    tus = bench_clip(name)
    ncpu = os.cpu_count() or 1
    dav1d_ref.decode(tus[:8], n_threads=ncpu, keep=False)
    best, n = None, 0
    for _ in range(passes):
        n, dt = dav1d_pass(tus, ncpu)
        best = dt if best is None else min(best, dt)
    k1 = min(len(tus), 10)
    n1, dt1 = dav1d_pass(tus[:k1], 1)
    return {"value": n / best, "unit": "frames/s", "cores": ncpu, "kind": "reference", "ms_per_step": best * 1e3,
            "single_thread_fps": n1 / dt1,
            "sample": f"libdav1d {dav1d_ref.version()} driven directly (no ffmpeg binary in the image; omits FFmpeg's demux), n_threads={ncpu}, whole "
                      f"{len(tus)}-TU clip {name} preloaded in RAM, best of {passes}, no MD5; single-thread figure on the first {k1} units"}


# ------------------------------------------------------------------------------------------
# workload: BASELINE configs[4] -- batch of 32 4K 10-bit files, GOP-segment-sharded across the GPUs (no collective)
# ------------------------------------------------------------------------------------------
C5_DESC = ("c5_batch_4k10: BASELINE configs[4] -- batch of 32 synthetic 3840x2160 10-bit files (source and tools as c3, seeds 100..131, libaom "
           "cpu-used 2, 16 frames each in two closed GOPs of 8 => 64 independent GOP segments), segments assigned to the ranks longest-first by "
           "coded bytes, one engine per GPU, no collective on the data path (strong scaling: the batch is fixed, the per-rank share shrinks with N)")
C5_FILES = int(os.environ.get("AV1R_C5_FILES", "32"))


def c5_items(nfiles=C5_FILES):
    """-> (items [(key, weight)], tus_of {key: [tu bytes]}, (w, h))"""
    import av1recon
    from av1recon import shard
    items, tus_of = [], {}
    for f in range(nfiles):
        tus = get_clip(f"c5_{f:02d}")
        for s_idx, (a, b) in enumerate(shard.split_segments(tus, av1recon.scan_headers(tus))):
            key = (f, s_idx)
            tus_of[key] = tus[a:b]
            items.append((key, sum(len(t) for t in tus[a:b])))
    return items, tus_of, (3840, 2160)


def run_c5(args, torch, dist, rank, world, local, steps=None, warmup=None, sample_clocks=True):
    import av1recon
    from av1recon import shard
    items, tus_of, (w, h) = c5_items()
    mine = shard.assign(items, world)[rank]
    my_tus = [t for k in mine for t in tus_of[k]]
    if world == 1:   # whole batch on this GPU: gate on the tool histogram of the batch
        pass
    blob = shard.ivf_bytes(my_tus, w, h)      # this rank's share as one container (the segments stay independent: each starts with a key frame)
    out = measure_clip("c5", my_tus, blob, args, torch, dist, rank, world, local, steps or args.steps, warmup or args.warmup,
                       C5_DESC, sample_clocks=sample_clocks, e2e_reps=3, gate_key="c5")
    out["scaling"] = "strong"
    out["engine"].update({"items": len(items), "items_rank0": len(mine)})
    return out


def cpu_baseline_c5(nfiles=8):
    """libdav1d on all host cores over a bounded sample of the batch (the first `nfiles` files, one after the other)."""
    from oracle import dav1d_ref
    ncpu = os.cpu_count() or 1
    files = [get_clip(f"c5_{f:02d}") for f in range(nfiles)]
    dav1d_ref.decode(files[0], n_threads=ncpu, keep=False)
    t0 = time.perf_counter()
    n = 0
    for tus in files:
        n += len(dav1d_ref.decode(tus, n_threads=ncpu, keep=False))
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": ncpu, "kind": "reference", "ms_per_step": dt * 1e3,
            "sample": f"libdav1d {dav1d_ref.version()} n_threads={ncpu}, files c5_00..c5_{nfiles - 1:02d} of the batch ({nfiles} x 16 frames 4K10) decoded "
                      "back to back, preloaded in RAM, no MD5"}


# clip key (tools/make_streams.py) -> bench workload name (profiles/ncu_traffic.json is keyed by the latter)
WORKLOAD_NAMES = {"c2": "c2_intra_1080p8", "c1": "c1_1080p8", "c3": "c3_4k10_inter", "c4": "c4_4k10_grain", "c5": "c5_batch_4k10",
                  "c3_small": "c3_small", "c2_small": "c2_small"}


def _clip_workload(name):
    return (lambda *a, **kw: run_clip(*a, name=name, **kw)), (lambda: cpu_baseline_clip(name)), (lambda: bench_clip(name))


def _c5_sample():
    return [t for f in range(8) for t in get_clip(f"c5_{f:02d}")]


WORKLOADS = {"filmgrain_4k10": (run_filmgrain, cpu_baseline_filmgrain, None),
             "c2_intra_1080p8": _clip_workload("c2"), "c1_1080p8": _clip_workload("c1"), "c3_4k10_inter": _clip_workload("c3"),
             "c4_4k10_grain": _clip_workload("c4"), "c3_small": _clip_workload("c3_small"), "c2_small": _clip_workload("c2_small"),
             "c5_batch_4k10": (run_c5, cpu_baseline_c5, _c5_sample)}
HEADLINE_N1 = "c3_4k10_inter"          # BASELINE metric's 4K10 half, the config north_star names as the target
HEADLINE_NGPU = "c5_batch_4k10"        # BASELINE configs[4]: the batch sharded over the GPUs
PER_CONFIG = ["c1_1080p8", "c2_intra_1080p8", "c4_4k10_grain", "c5_batch_4k10"]


FRAMES_PER_STEP = {"c1_1080p8": 60, "c2_intra_1080p8": 60, "c3_4k10_inter": 60 * CLIP_REPEAT["c3"], "c4_4k10_grain": 60 * CLIP_REPEAT["c4"], "c3_small": 20, "c2_small": 8,
                   "c5_batch_4k10": 16 * C5_FILES}
DESC_OF = {"c1_1080p8": "c1", "c2_intra_1080p8": "c2", "c3_4k10_inter": "c3", "c4_4k10_grain": "c4", "c3_small": "c3_small", "c2_small": "c2_small"}


def workload_config(workload, world):
    """The `config` object of the JSON line: identical for the b200 arm and the reference arm of the same command."""
    if workload == "c5_batch_4k10":
        return {"workload": C5_DESC + "; step = one pass over the whole batch (reference arm: a bounded sample, the first 8 of the 32 files)", "frames_per_step": FRAMES_PER_STEP[workload], "files": C5_FILES,
                "parallelism": f"gop-segment sharding over {world} GPU(s), longest-first by coded bytes, no collective",
                "l2": "per-step working set exceeds the 126 MB L2 (no flush needed)"}
    if workload in DESC_OF:
        rep = CLIP_REPEAT.get(DESC_OF[workload], 1)
        rep_note = (f"; benchmarked as the 60-frame sequence {rep} times back to back = {60 * rep} frames in {2 * rep} closed GOPs (both arms)" if rep > 1 else "")
        return {"workload": CLIP_DESC[DESC_OF[workload]] + rep_note + STEP_NOTE, "frames_per_step": FRAMES_PER_STEP[workload],
                "parallelism": f"replicas{world} (independent clips per GPU, no collective)" if world > 1 else "single GPU",
                "l2": "per-step working set exceeds the 126 MB L2 (no flush needed)"}
    return {"workload": workload}


def compact(line):
    """per_config entry: the figures the verdict asked for, without the long tables."""
    r = line["roofline"]
    return {"workload": line["config"]["workload"].split(" --")[0].split(":")[0], "config": line["config"], "value": line["value"], "value_hbm_resident": line.get("value_hbm_resident"), "value_fenced_steps": line.get("value_fenced_steps"),
            "unit": "frames/s", "steps": line["steps"], "ms_per_step": line["ms_per_step"], "dtype": line["dtype"], "scaling": line["scaling"],
            "e2e": {k: line["e2e"][k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "host_parse_ms_per_step", "host_threads")},
            "cpu_baseline": line.get("cpu_baseline"),
            "roofline": {"kernel": r["kernel"], "achieved": r["achieved"], "peak": r["peak"], "frac": r["frac"], "unit": "GB/s",
                         "pipeline_algorithmic_gbs": r["pipeline_algorithmic_gbs"],
                         "stages": {k: {kk: v[kk] for kk in ("ms_per_step", "launches", "algorithmic_gbs") if kk in v} for k, v in r["stages"].items()}},
            "gpu_launches": line["gpu_launches"], "host_parse_ms_per_frame": line["host_parse_ms_per_frame"],
            "clip": {k: line["clip"][k] for k in ("frames", "lr_frames", "cdef_frames", "deblock_frames", "grain_frames", "tools")}}


def reference_arm(args, workload):
    """The reference's CPU implementation of the path (libdav1d, all host threads) on this arm's workload: W warm-up steps, then K
    timed steps; a step = one pass over the clip (c5: over a bounded sample of 8 of the 32 files)."""
    from oracle import dav1d_ref
    sample = WORKLOADS[workload][2]
    if sample is None:
        cb = WORKLOADS[workload][1]()
        return cb, cb["value"], None
    tus = sample()
    ncpu = os.cpu_count() or 1
    for _ in range(args.warmup):
        dav1d_pass(tus, ncpu)
    frames, t = 0, 0.0
    for _ in range(args.steps):
        n, dt = dav1d_pass(tus, ncpu)
        frames += n
        t += dt
    what = "first 8 of the 32 files (128 frames) per step" if workload == "c5_batch_4k10" else f"whole {len(tus)}-TU clip per step"
    cb = {"value": frames / t, "unit": "frames/s", "cores": ncpu, "kind": "reference",
          "sample": f"libdav1d {dav1d_ref.version()} driven directly (no ffmpeg binary in the image; omits FFmpeg's demux), n_threads={ncpu}, {what}, "
                    f"{args.warmup} warm-up + {args.steps} timed steps, stream preloaded in RAM, no MD5"}
    return cb, frames / t, t * 1e3 / args.steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()
    workload = args.workload or (HEADLINE_N1 if max(world, args.gpus) == 1 else HEADLINE_NGPU)
    run, cpu_base, _ = WORKLOADS[workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, value, ms_per_step = reference_arm(args, workload)
        line = {"impl": "reference", "metric": "AV1 decode-verify frames/s", "value": value, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                "scaling": "strong" if workload == "c5_batch_4k10" else "weak", "vs_baseline": None, "data": "synthetic",
                "config": workload_config(workload, max(1, args.gpus)), "ms_per_step": ms_per_step,
                "dtype": "u8" if workload in ("c1_1080p8", "c2_intra_1080p8", "c2_small") else "u16",
                "cpu_baseline": cb,
                "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    if rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_base()
    # N = 1 default run: the other BASELINE configs ride along as compact entries (fewer steps each; same timing rules)
    if world == 1 and args.workload is None and not args.no_per_config:
        out["per_config"] = []
        for wname in PER_CONFIG:
            r2, cb2, _ = WORKLOADS[wname]
            k = max(3, min(args.steps, 5 if wname != "c5_batch_4k10" else 3))
            line = r2(args, torch, dist, rank, world, local, steps=k, warmup=3, sample_clocks=False)
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cb2()
            out["per_config"].append(compact(line))
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

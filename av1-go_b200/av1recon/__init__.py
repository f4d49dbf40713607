"""ctypes binding of libav1r.so (the C ABI in include/av1r.h, include/av1r_stages.h).

This is the Python stand-in for the cgo package `internal/av1recon` (av1-go_b200/go/): the
image has no Go toolchain, so tests and bench.py drive the same C entry points from here.
There is no CPU fallback: if the CUDA library is missing, importing raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libav1r.so")


class FilmGrainParams(C.Structure):
    _fields_ = [
        ("apply_grain", C.c_int), ("grain_seed", C.c_int), ("update_grain", C.c_int),
        ("num_y_points", C.c_int), ("point_y_value", C.c_int * 16), ("point_y_scaling", C.c_int * 16),
        ("chroma_scaling_from_luma", C.c_int),
        ("num_cb_points", C.c_int), ("point_cb_value", C.c_int * 16), ("point_cb_scaling", C.c_int * 16),
        ("num_cr_points", C.c_int), ("point_cr_value", C.c_int * 16), ("point_cr_scaling", C.c_int * 16),
        ("grain_scaling", C.c_int), ("ar_coeff_lag", C.c_int),
        ("ar_coeffs_y", C.c_int * 24), ("ar_coeffs_cb", C.c_int * 25), ("ar_coeffs_cr", C.c_int * 25),
        ("ar_coeff_shift", C.c_int), ("grain_scale_shift", C.c_int),
        ("cb_mult", C.c_int), ("cb_luma_mult", C.c_int), ("cb_offset", C.c_int),
        ("cr_mult", C.c_int), ("cr_luma_mult", C.c_int), ("cr_offset", C.c_int),
        ("overlap_flag", C.c_int), ("clip_to_restricted_range", C.c_int),
    ]


class FrameHeaderInfo(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("tu_index", C.c_int),
        ("frame_type", C.c_int), ("show_frame", C.c_int), ("showable_frame", C.c_int),
        ("show_existing_frame", C.c_int), ("frame_to_show_map_idx", C.c_int),
        ("width", C.c_int), ("height", C.c_int), ("upscaled_width", C.c_int), ("bit_depth", C.c_int),
        ("subsampling_x", C.c_int), ("subsampling_y", C.c_int), ("mono_chrome", C.c_int),
        ("matrix_coefficients", C.c_int),
        ("refresh_frame_flags", C.c_int), ("order_hint", C.c_int), ("primary_ref_frame", C.c_int),
        ("base_q_idx", C.c_int), ("tile_cols", C.c_int), ("tile_rows", C.c_int), ("use_128x128_superblock", C.c_int),
        ("lf_level", C.c_int * 4), ("cdef_enabled", C.c_int), ("cdef_bits", C.c_int), ("lr_type", C.c_int * 3),
        ("tx_mode", C.c_int), ("reduced_tx_set", C.c_int), ("header_bytes", C.c_int),
        ("film_grain", FilmGrainParams),
        ("error_resilient_mode", C.c_int), ("disable_cdf_update", C.c_int), ("disable_frame_end_update_cdf", C.c_int),
        ("enable_order_hint", C.c_int), ("coded_lossless", C.c_int),
        ("segmentation_enabled", C.c_int), ("segmentation_update_map", C.c_int), ("segmentation_temporal_update", C.c_int),
        ("delta_q_present", C.c_int), ("delta_lf_present", C.c_int),
    ]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int), ("streams", C.c_int),
                ("frames_in_flight", C.c_int), ("parity_md5", C.c_int), ("apply_grain", C.c_int),
                ("inloop_filters", C.c_int), ("keep_frames", C.c_int), ("host_threads", C.c_int)]


class FrameResult(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("pts", C.c_int64),
                ("w", C.c_int), ("h", C.c_int), ("bpc", C.c_int), ("layout", C.c_int),
                ("status", C.c_int), ("frame_type", C.c_int), ("shown_existing", C.c_int),
                ("md5", (C.c_uint8 * 16) * 3), ("checksum", C.c_uint64 * 3),
                ("host_parse_ms", C.c_float), ("device_ms", C.c_float), ("frame_handle", C.c_int64)]


class StreamInfo(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("is_av1", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("bit_depth", C.c_int), ("profile", C.c_int),
                ("subsampling_x", C.c_int), ("subsampling_y", C.c_int), ("mono_chrome", C.c_int),
                ("film_grain_present", C.c_int), ("temporal_units", C.c_int64), ("keyframes", C.c_int64)]


class Report(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("status", C.c_int), ("frames", C.c_int64),
                ("width", C.c_int), ("height", C.c_int), ("bit_depth", C.c_int),
                ("first_bad_frame", C.c_int64),
                ("host_parse_ms", C.c_double), ("device_ms", C.c_double), ("wall_ms", C.c_double),
                ("frames_per_sec", C.c_double), ("message", C.c_char * 512)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make` (or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        l.av1r_abi_version.restype = C.c_uint32
        l.av1r_scan_headers.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int,
                                        C.POINTER(FrameHeaderInfo), C.c_int, C.POINTER(C.c_int)]
        l.av1r_film_grain_scratch_bytes.restype = C.c_size_t
        l.av1r_stage_film_grain.argtypes = [C.POINTER(FilmGrainParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p]
        l.av1r_stage_plane_checksum.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.av1r_plane_checksum_host.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        l.av1r_plane_checksum_host.restype = C.c_uint64
        l.av1r_stage_last_error.restype = C.c_char_p
        l.av1r_probe_buffer.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StreamInfo)]
        l.av1r_probe_file.argtypes = [C.c_char_p, C.POINTER(StreamInfo)]
        _lib = l
    return _lib


def scan_headers(tus):
    """Header-only scan of a list of temporal units -> list of FrameHeaderInfo."""
    l = lib()
    n_tus = len(tus)
    arr = (C.c_char_p * n_tus)(*tus)
    lens = (C.c_size_t * n_tus)(*[len(t) for t in tus])
    cap = 4 * n_tus + 8
    out = (FrameHeaderInfo * cap)()
    n = C.c_int(0)
    rc = l.av1r_scan_headers(arr, lens, n_tus, out, cap, C.byref(n))
    if rc:
        raise RuntimeError(f"av1r_scan_headers -> {rc}")
    return [out[i] for i in range(n.value)]


def probe_buffer(data):
    info = StreamInfo()
    info.struct_size = C.sizeof(StreamInfo)
    rc = lib().av1r_probe_buffer(data, len(data), C.byref(info))
    if rc:
        raise RuntimeError(f"av1r_probe_buffer -> {rc}")
    return info


class Decoder:
    """Thin wrapper over av1r_open / av1r_submit_tu / av1r_collect (the calls the cgo package makes)."""

    def __init__(self, device=0, parity_md5=0, apply_grain=1, inloop_filters=7, keep_frames=0, streams=2, frames_in_flight=8, host_threads=0):
        l = lib()
        l.av1r_open.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        l.av1r_close.argtypes = [C.c_void_p]
        l.av1r_submit_tu.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int64]
        l.av1r_collect.argtypes = [C.c_void_p, C.POINTER(FrameResult), C.c_int, C.POINTER(C.c_int)]
        l.av1r_flush.argtypes = [C.c_void_p]
        l.av1r_last_error.argtypes = [C.c_void_p]
        l.av1r_last_error.restype = C.c_char_p
        l.av1r_copy_frame.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_size_t]
        cfg = Config()
        l.av1r_default_config(C.byref(cfg))
        cfg.device, cfg.parity_md5, cfg.apply_grain, cfg.inloop_filters = device, parity_md5, apply_grain, inloop_filters
        cfg.keep_frames, cfg.streams, cfg.frames_in_flight = keep_frames, streams, frames_in_flight
        cfg.host_threads = host_threads
        self.l = l
        self.ctx = C.c_void_p()
        rc = l.av1r_open(C.byref(cfg), C.byref(self.ctx))
        if rc:
            raise RuntimeError(f"av1r_open -> {rc}")
        self.results = []

    def error(self):
        return self.l.av1r_last_error(self.ctx).decode()

    def submit(self, tu, pts=0):
        rc = self.l.av1r_submit_tu(self.ctx, tu, len(tu), pts)
        if rc:
            raise RuntimeError(f"av1r_submit_tu -> {rc}: {self.error()}")
        self._collect()

    def _collect(self):
        buf = (FrameResult * 32)()
        while True:
            n = C.c_int(0)
            rc = self.l.av1r_collect(self.ctx, buf, 32, C.byref(n))
            if rc:
                raise RuntimeError(f"av1r_collect -> {rc}: {self.error()}")
            for i in range(n.value):
                r = FrameResult()
                C.memmove(C.byref(r), C.byref(buf[i]), C.sizeof(FrameResult))
                self.results.append(r)
            if n.value == 0:
                break

    def flush(self):
        rc = self.l.av1r_flush(self.ctx)
        if rc:
            raise RuntimeError(f"av1r_flush -> {rc}: {self.error()}")
        self._collect()

    def frame_planes(self, res):
        """keep_frames mode: copy the three planes of a collected frame to numpy arrays."""
        import numpy as np
        out = []
        bps = 1 if res.bpc == 8 else 2
        for p in range(3):
            w = res.w if p == 0 else (res.w + 1) // 2
            h = res.h if p == 0 else (res.h + 1) // 2
            cw = (w + 7) // 8 * 8 if p == 0 else ((res.w + 7) // 8 * 8) // 2
            chh = (h + 7) // 8 * 8 if p == 0 else ((res.h + 7) // 8 * 8) // 2
            a = np.zeros((chh, cw * bps), dtype=np.uint8)
            rc = self.l.av1r_copy_frame(self.ctx, res.frame_handle, p, a.ctypes.data, cw * bps)
            if rc:
                raise RuntimeError(f"av1r_copy_frame -> {rc}")
            a = a.view(np.uint8 if bps == 1 else np.dtype("<u2"))[:h, :w]
            out.append(np.ascontiguousarray(a))
        return out

    def verify_buffer(self, data, want_digests=True, max_frames=100000):
        """av1r_ctx_verify_buffer on this (persistent) engine -> (rc, Report, digests)."""
        self.l.av1r_ctx_verify_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(Report), C.POINTER(C.c_uint64), C.c_int64]
        rep = Report()
        dig = (C.c_uint64 * (3 * max_frames))() if want_digests else None
        rc = self.l.av1r_ctx_verify_buffer(self.ctx, data, len(data), C.byref(rep), dig, max_frames if want_digests else 0)
        digs = [tuple(dig[3 * i:3 * i + 3]) for i in range(int(rep.frames))] if want_digests else []
        return rc, rep, digs

    def close(self):
        if self.ctx:
            self.l.av1r_close(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ClipInfo(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_int), ("height", C.c_int), ("bit_depth", C.c_int),
                ("frames_decoded", C.c_int64), ("frames_shown", C.c_int64), ("host_parse_ms", C.c_double),
                ("worklist_bytes", C.c_uint64), ("frame_bytes", C.c_uint64), ("coded_samples", C.c_uint64),
                ("coef_tokens", C.c_uint64), ("tx_blocks", C.c_uint64), ("intra_samples", C.c_uint64),
                ("inter_samples", C.c_uint64), ("inter_ref_samples", C.c_uint64), ("lr_frames", C.c_uint64), ("cdef_frames", C.c_uint64),
                ("deblock_frames", C.c_uint64), ("grain_frames", C.c_uint64), ("inter_blocks", C.c_uint64), ("obmc_neighbours", C.c_uint64),
                ("tool_hist", C.c_uint64 * 24),
                ("intra_frame_samples", C.c_uint64), ("intra_frame_coded_samples", C.c_uint64), ("intra_frame_tx_blocks", C.c_uint64),
                ("intra_frames", C.c_uint64)]


TOOL_NAMES = ["inter_blocks", "compound_avg", "compound_dist", "compound_wedge", "compound_diffwtd", "interintra", "interintra_wedge",
              "obmc", "local_warp", "global_warp", "skip_mode", "dual_filter", "temporal_mv", "intra_in_inter", "sub8x8_chroma", "newmv",
              "vartx_split", "switchable_filter", "palette", "intrabc"]


STAGES = ["h2d", "itx", "intra", "inter", "deblock", "cdef", "lr", "grain", "digest", "superres", "intra_frame"]


class StageTimes(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("ms", C.c_float * len(STAGES)), ("launches", C.c_int * len(STAGES))]


class Clip:
    """Device-resident clip replay (av1r_clip_*): parse + upload once, time the device path."""

    def __init__(self, dec, tus):
        l = dec.l
        l.av1r_clip_load.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_void_p)]
        l.av1r_clip_info_get.argtypes = [C.c_void_p, C.POINTER(ClipInfo)]
        l.av1r_clip_decode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        l.av1r_clip_profile.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StageTimes)]
        l.av1r_clip_free.argtypes = [C.c_void_p]
        self.dec = dec
        n = len(tus)
        arr = (C.c_char_p * n)(*tus)
        lens = (C.c_size_t * n)(*[len(t) for t in tus])
        self.h = C.c_void_p()
        rc = l.av1r_clip_load(dec.ctx, arr, lens, n, C.byref(self.h))
        if rc:
            raise RuntimeError(f"av1r_clip_load -> {rc}: {dec.error()}")
        self.info = ClipInfo()
        l.av1r_clip_info_get(self.h, C.byref(self.info))

    def decode(self):
        """-> (device_ms, checksums[list of 3-tuples])"""
        cap = int(self.info.frames_shown) + 4
        cks = (C.c_uint64 * (3 * cap))()
        n = C.c_int(0)
        ms = C.c_float(0)
        rc = self.dec.l.av1r_clip_decode(self.dec.ctx, self.h, cks, cap, C.byref(n), C.byref(ms))
        if rc:
            raise RuntimeError(f"av1r_clip_decode -> {rc}: {self.dec.error()}")
        return ms.value, [tuple(cks[3 * i:3 * i + 3]) for i in range(n.value)]

    def decode_passes(self, passes):
        """`passes` replays back to back inside one fenced region -> (device_ms of all passes, checksums of the first pass); the
        library checks that every later pass reproduces them."""
        cap = int(self.info.frames_shown) + 4
        cks = (C.c_uint64 * (3 * cap))()
        n = C.c_int(0)
        ms = C.c_float(0)
        f = self.dec.l.av1r_clip_decode_passes
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        rc = f(self.dec.ctx, self.h, int(passes), cks, cap, C.byref(n), C.byref(ms))
        if rc:
            raise RuntimeError(f"av1r_clip_decode_passes -> {rc}: {self.dec.error()}")
        return ms.value, [tuple(cks[3 * i:3 * i + 3]) for i in range(n.value)]

    def profile(self):
        st = StageTimes()
        rc = self.dec.l.av1r_clip_profile(self.dec.ctx, self.h, C.byref(st))
        if rc:
            raise RuntimeError(f"av1r_clip_profile -> {rc}: {self.dec.error()}")
        return {STAGES[i]: (st.ms[i], st.launches[i]) for i in range(len(STAGES)) if STAGES[i] != "superres" or st.launches[i]}

    def set_resident(self, resident):
        """resident=True: work-lists uploaded once, replays read them from HBM (kernel-only figure); default False: one H2D per
        frame inside every replay."""
        self.dec.l.av1r_clip_set_resident.argtypes = [C.c_void_p, C.c_int]
        self.dec.l.av1r_clip_set_resident(self.h, 1 if resident else 0)

    def free(self):
        if self.h:
            self.dec.l.av1r_clip_free(self.h)
            self.h = C.c_void_p()


def parse_stats(data):
    """Host-only statistics of a container (tool histogram, frames per post-filter stage, coded samples ...) -> ClipInfo."""
    l = lib()
    l.av1r_parse_stats.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(ClipInfo)]
    ci = ClipInfo()
    rc = l.av1r_parse_stats(data, len(data), C.byref(ci))
    if rc:
        raise RuntimeError(f"av1r_parse_stats -> {rc}")
    return ci


def tool_hist(ci):
    return {k: int(v) for k, v in zip(TOOL_NAMES, ci.tool_hist) if v}


class Pool:
    """av1r_pool_*: one engine + host thread per device inside this process; a batch of files is cut into GOP segments and
    assigned to the devices longest-first (what the single-process daemon calls to use several GPUs)."""

    def __init__(self, devices, host_threads=0, streams=16, frames_in_flight=32, apply_grain=1, inloop_filters=7):
        l = lib()
        l.av1r_pool_open.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(Config), C.POINTER(C.c_void_p)]
        l.av1r_pool_close.argtypes = [C.c_void_p]
        l.av1r_pool_verify_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.POINTER(Report),
                                               C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_int64), C.POINTER(Report)]
        l.av1r_pool_verify_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(Report), C.POINTER(Report)]
        cfg = Config()
        l.av1r_default_config(C.byref(cfg))
        cfg.host_threads, cfg.streams, cfg.frames_in_flight = host_threads, streams, frames_in_flight
        cfg.apply_grain, cfg.inloop_filters = apply_grain, inloop_filters
        self.l = l
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = l.av1r_pool_open(devs, len(devices), C.byref(cfg), C.byref(self.h))
        if rc:
            raise RuntimeError(f"av1r_pool_open -> {rc}")
        self.devices = list(devices)

    def verify_buffers(self, blobs, max_frames=4096, want_digests=True):
        """-> (rc, total Report, [Report per file], [digests per file])"""
        n = len(blobs)
        arr = (C.c_char_p * n)(*blobs)
        lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        reps = (Report * n)()
        total = Report()
        if want_digests:
            bufs = [(C.c_uint64 * (3 * max_frames))() for _ in range(n)]
            dptr = (C.POINTER(C.c_uint64) * n)(*[C.cast(b, C.POINTER(C.c_uint64)) for b in bufs])
            caps = (C.c_int64 * n)(*([max_frames] * n))
        else:
            bufs, dptr, caps = [], None, None
        rc = self.l.av1r_pool_verify_buffers(self.h, arr, lens, n, reps, dptr, caps, C.byref(total))
        digs = [[tuple(b[3 * i:3 * i + 3]) for i in range(int(reps[f].frames))] for f, b in enumerate(bufs)]
        return rc, total, [reps[i] for i in range(n)], digs

    def verify_files(self, paths):
        n = len(paths)
        arr = (C.c_char_p * n)(*[p.encode() for p in paths])
        reps = (Report * n)()
        total = Report()
        rc = self.l.av1r_pool_verify_files(self.h, arr, n, reps, C.byref(total))
        return rc, total, [reps[i] for i in range(n)]

    def close(self):
        if self.h:
            self.l.av1r_pool_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def batch_assign(weights, n_devices):
    l = lib()
    l.av1r_batch_assign.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int, C.POINTER(C.c_int)]
    n = len(weights)
    out = (C.c_int * n)()
    rc = l.av1r_batch_assign((C.c_uint64 * n)(*weights), n, n_devices, out)
    if rc:
        raise RuntimeError(f"av1r_batch_assign -> {rc}")
    return list(out)


def verify_buffer(data, device=0, host_threads=0, apply_grain=1, inloop_filters=7, streams=16, frames_in_flight=32, want_digests=True, max_frames=100000):
    """av1r_verify_buffer: whole container in host memory -> (Report, [digest triples])."""
    l = lib()
    l.av1r_verify_buffer.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Config), C.POINTER(Report), C.POINTER(C.c_uint64), C.c_int64]
    cfg = Config()
    l.av1r_default_config(C.byref(cfg))
    cfg.device, cfg.host_threads, cfg.apply_grain, cfg.inloop_filters = device, host_threads, apply_grain, inloop_filters
    cfg.streams, cfg.frames_in_flight = streams, frames_in_flight
    rep = Report()
    dig = (C.c_uint64 * (3 * max_frames))() if want_digests else None
    rc = l.av1r_verify_buffer(data, len(data), C.byref(cfg), C.byref(rep), dig, max_frames if want_digests else 0)
    digs = [tuple(dig[3 * i:3 * i + 3]) for i in range(int(rep.frames))] if want_digests else []
    return rc, rep, digs


# ---- mirror of the Go surface the daemon would use (av1-go_b200/go/internal/av1recon/av1recon.go) --------------
# Same names, argument meaning and error behaviour as the cgo package: Open(device) -> Engine, Engine.VerifyFile(path) -> Report or
# VerifyError with the text that would land in job.Reason, Engine.Close(); ProbeFile / VerifyOutput follow
# /root/reference/internal/metadata/probe.go:125 and the VerifyOutput helper of INTEGRATION.md section 3.
class VerifyError(RuntimeError):
    def __init__(self, code, message, report=None):
        super().__init__(f"av1 verify failed (code {code}): {message}")
        self.code = code
        self.report = report


class Engine:
    def __init__(self, device=0):
        self._dec = Decoder(device=device, streams=16, frames_in_flight=32)

    def VerifyFile(self, path):
        """Decodes every frame of an AV1 file (Matroska / IVF / raw OBU) on the GPU -> Report; raises VerifyError on failure."""
        try:
            data = open(path, "rb").read()
        except OSError as e:
            raise VerifyError(-5, f"failed to read {path}: {e}")
        if not data:   # the broken-output case the verifier exists for (the Go binding must not index data[0] either)
            raise VerifyError(-74, f"{path} is empty")
        rc, rep, _ = self._dec.verify_buffer(data, want_digests=False)
        if rc != 0:
            raise VerifyError(rc, rep.message.decode(errors="replace"), rep)
        return rep

    def Close(self):
        self._dec.close()


def Open(device=0):
    """av1recon.Open: one engine = one GPU.  Raises RuntimeError when no CUDA device is present (the daemon then skips verification,
    like it tolerates a failed QSV self-test: /root/reference/cmd/av1d/main.go:41-52)."""
    return Engine(device)


def ProbeFile(path):
    """Width / Height / BitDepth / is-AV1 of a file from its sequence header (metadata.ProbeFile without the ffprobe child)."""
    info = StreamInfo()
    info.struct_size = C.sizeof(StreamInfo)
    l = lib()
    l.av1r_probe_file.argtypes = [C.c_char_p, C.POINTER(StreamInfo)]
    rc = l.av1r_probe_file(os.fsencode(path), C.byref(info))
    if rc:
        raise RuntimeError(f"av1r_probe_file({path}) -> {rc}")
    return info


def VerifyOutput(eng, output_path, src_width, src_height, is_webrip_like=False):
    """ffmpeg.VerifyOutput (av1-go_b200/go/internal/ffmpeg/verify.go): decode the transcoded file and check it against what the
    probe said about the source.  The non-WebRip filter chain only rounds odd dimensions up to even
    (/root/reference/internal/ffmpeg/transcode.go:105-112), so the size must match exactly; the WebRip chain first rescales by the
    sample aspect ratio (transcode.go:93-101) which the probe result does not carry, so there only even dimensions not narrower
    than the source are required."""
    rep = eng.VerifyFile(output_path)
    if rep.frames == 0:
        raise VerifyError(-22, "no frames decoded", rep)
    if is_webrip_like:
        if rep.width % 2 or rep.height % 2 or rep.width < src_width:
            raise VerifyError(-22, f"decoded size {rep.width}x{rep.height} is not a valid rescale of source {src_width}x{src_height}", rep)
    else:
        want_w, want_h = (src_width + 1) // 2 * 2, (src_height + 1) // 2 * 2
        if (rep.width, rep.height) != (want_w, want_h):
            raise VerifyError(-22, f"decoded size {rep.width}x{rep.height} does not match source {src_width}x{src_height}", rep)
    return rep

/* av1r -- C ABI of the B200-native AV1 decode-verify engine.
 *
 * Drop-in boundary for IONIQ6000/av1-go.  The reference has no FFI: its operator boundary is
 * "a blocking Go function in internal/ffmpeg (or internal/metadata) that talks to an external
 * engine and returns (value, error)":
 *   RunTranscode(ffmpegPath, args) (int, error)      /root/reference/internal/ffmpeg/transcode.go:194
 *   VerifyFFmpeg(ffmpegPath) error                   /root/reference/internal/ffmpeg/binary.go:218
 *   ProbeFile(ffmpegPath, filePath) (*ProbeResult, error)  /root/reference/internal/metadata/probe.go:125
 * The entry points below are what a cgo package `internal/av1recon` binds to give the daemon
 *   ffmpeg.VerifyOutput(outputPath, probe) (*VerifyReport, error)
 * in the empty slot of daemon.ProcessJob between /root/reference/internal/daemon/daemon.go:112
 * (transcode succeeded) and :115 (stat output / size gate).  See INTEGRATION.md.
 *
 * Conventions: C99, plain pointers and sizes, no callbacks.  Functions return 0 or a negative
 * errno-style code (AV1R_E*); av1r_last_error() gives text.  One av1r_ctx = one GPU; a ctx is
 * not thread-safe, distinct ctxs are independent (8 GPUs = 8 ctxs on 8 host threads).  The
 * library owns all device and pinned memory; the caller owns `data` and result arrays.  Every
 * struct starts with struct_size for forward compatibility.
 */
#ifndef AV1R_H
#define AV1R_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AV1R_ABI_VERSION 0x00010000u

#define AV1R_OK 0
#define AV1R_EAGAIN (-11)    /* nothing to collect yet */
#define AV1R_ENOMEM (-12)
#define AV1R_EINVAL (-22)
#define AV1R_EIO (-5)        /* CUDA fault / device unavailable */
#define AV1R_ENOSYS (-38)    /* bitstream tool not supported by this build */
#define AV1R_EBITSTREAM (-74) /* EBADMSG: corrupt / non-conformant bitstream */
#define AV1R_ENOENT (-2)

typedef struct av1r_ctx av1r_ctx;

typedef struct av1r_config {
    uint32_t struct_size;
    int device;            /* CUDA device ordinal */
    int streams;           /* CUDA streams used for frame pipelining (0 = default 2) */
    int frames_in_flight;  /* max frames queued on the device before submit blocks (0 = default 8) */
    int parity_md5;        /* 1: copy every shown frame back and MD5 its planes (parity mode, untimed) */
    int apply_grain;       /* 1 (default): film grain synthesis on output frames, as libdav1d does */
    int inloop_filters;    /* bit mask 1=deblock 2=CDEF 4=loop restoration (7 = all; same meaning as dav1d) */
    int keep_frames;       /* 1: keep shown frames on the device so av1r_copy_frame can read them (tests) */
    int host_threads;      /* av1r_verify_*: host threads parsing independent key-frame-delimited GOP segments
                              in parallel (0 = number of online cores, capped at 32) */
} av1r_config;

typedef struct av1r_frame_result {
    uint32_t struct_size;
    int64_t pts;
    int w, h, bpc, layout;       /* layout: 0 I400, 1 I420, 2 I422, 3 I444 */
    int status;                  /* 0 ok, else AV1R_E* */
    int frame_type;              /* 0 key 1 inter 2 intra-only 3 switch */
    int shown_existing;          /* produced by show_existing_frame */
    uint8_t md5[3][16];          /* valid when parity_md5 */
    uint64_t checksum[3];        /* device-side 64-bit plane checksum (always) */
    float host_parse_ms;         /* sequential symbol parse on the host (reported separately) */
    float device_ms;             /* CUDA-event time of this frame's reconstruction */
    int64_t frame_handle;        /* valid when keep_frames; pass to av1r_copy_frame */
} av1r_frame_result;

/* Geometry of the first video sequence found -- the native stand-in for the fields the daemon
 * reads from ffprobe (metadata.StreamInfo Width/Height/BitDepth/CodecName, probe.go:34-46). */
typedef struct av1r_stream_info {
    uint32_t struct_size;
    int is_av1;
    int width, height, bit_depth, profile;
    int subsampling_x, subsampling_y, mono_chrome;
    int film_grain_present;
    int64_t temporal_units;      /* number of temporal units in the container */
    int64_t keyframes;           /* shown key frames = independently decodable GOP segments */
} av1r_stream_info;

typedef struct av1r_report {
    uint32_t struct_size;
    int status;                  /* 0 = every frame decoded */
    int64_t frames;              /* shown frames decoded */
    int width, height, bit_depth;
    int64_t first_bad_frame;     /* -1 if none */
    double host_parse_ms, device_ms, wall_ms;
    double frames_per_sec;
    char message[512];           /* short reason, sized like job.Reason (transcode.go:295-297) */
} av1r_report;

uint32_t av1r_abi_version(void);
void av1r_default_config(av1r_config* cfg);

int av1r_open(const av1r_config* cfg, av1r_ctx** out);
void av1r_close(av1r_ctx* ctx);
/* Submit one temporal unit.  `data` is parsed (and whatever is needed copied) before return. */
int av1r_submit_tu(av1r_ctx* ctx, const uint8_t* data, size_t len, int64_t pts);
/* Collect finished frames in display order.  Returns 0 and *n >= 0. */
int av1r_collect(av1r_ctx* ctx, av1r_frame_result* out, int cap, int* n);
/* Wait for all queued device work. */
int av1r_flush(av1r_ctx* ctx);
const char* av1r_last_error(const av1r_ctx* ctx);

/* keep_frames mode: copy plane `plane` of a collected frame into dst (tightly described by
 * dst_stride bytes per row).  Samples are uint8 (8-bit) or little-endian uint16. */
int av1r_copy_frame(av1r_ctx* ctx, int64_t frame_handle, int plane, void* dst, size_t dst_stride);
int av1r_release_frame(av1r_ctx* ctx, int64_t frame_handle);

/* Whole-file convenience: demux (IVF / raw OBU / Matroska), decode every frame, report. */
int av1r_verify_file(const char* path, const av1r_config* cfg, av1r_report* out);
/* Same on a container already in host memory (IVF / raw OBU / Matroska bytes).  If `digests` is non-NULL it
 * receives 3 x uint64 plane digests per shown frame in display order (cap_frames entries). */
int av1r_verify_buffer(const uint8_t* data, size_t len, const av1r_config* cfg, av1r_report* out,
                       uint64_t* digests, int64_t cap_frames);
/* Same, on an engine that stays open between files (what the daemon's job loop wants: one av1r_open at start-up,
 * one call per job; /root/reference/cmd/av1d/main.go:312-349).  Uses cfg.host_threads / streams of the ctx. */
int av1r_ctx_verify_buffer(av1r_ctx* ctx, const uint8_t* data, size_t len, av1r_report* out, uint64_t* digests, int64_t cap_frames);
/* ---- several GPUs inside one process (the daemon is one Go process: /root/reference/cmd/av1d/main.go:311-349) -------------------
 * A pool holds one engine + one host thread per device.  A batch of files is cut into key-frame-delimited GOP segments (a segment
 * starts only at a temporal unit whose first frame is a shown KEY_FRAME), the segments are assigned to the devices longest-first by
 * coded bytes, and every device parses (its share of the host cores) and reconstructs its segments; there is no exchange between
 * devices.  reports[f] describes file f (frames, status, first_bad_frame, message); *total (optional) aggregates the batch:
 * frames, wall_ms, frames_per_sec, and first_bad_frame = index of the first failing file.  digests[f] (optional, may be NULL or hold
 * NULL entries) receives 3 x uint64 plane digests per shown frame of file f in display order, cap_frames[f] entries. */
typedef struct av1r_pool av1r_pool;
int av1r_pool_open(const int* devices, int n_devices, const av1r_config* cfg, av1r_pool** out);
void av1r_pool_close(av1r_pool* pool);
int av1r_pool_devices(const av1r_pool* pool);
int av1r_pool_verify_files(av1r_pool* pool, const char* const* paths, int n_files, av1r_report* reports, av1r_report* total);
int av1r_pool_verify_buffers(av1r_pool* pool, const uint8_t* const* data, const size_t* lens, int n_files, av1r_report* reports,
                             uint64_t* const* digests, const int64_t* cap_frames, av1r_report* total);
/* open + verify_files + close */
int av1r_verify_batch(const char* const* paths, int n_files, const int* devices, int n_devices, const av1r_config* cfg,
                      av1r_report* reports, av1r_report* total);
/* The assignment rule alone (host only, for tests / planning): assignment[i] = device index of item i. */
int av1r_batch_assign(const uint64_t* weights, int n_items, int n_devices, int* assignment);

/* Host half only (no device): demux + symbol parse of every frame, GOP segments on `host_threads` threads (0 = all cores), the
 * tiles of a frame on the process-wide worker pool when tile_threads != 0.  Fills frames, host_parse_ms (summed over frames),
 * wall_ms and frames_per_sec: the "sequential parse, timed and reported separately" of the hot path. */
int av1r_parse_buffer(const uint8_t* data, size_t len, int host_threads, int tile_threads, av1r_report* out);
/* Host-only check of the intra kernel's plan (record order, 64x64 unit table, neighbour dependencies) for every frame of a
 * container: 0 or AV1R_EINVAL with the first violated invariant in msg.  Test / diagnosis entry point; touches no GPU. */
int av1r_debug_k3_check(const uint8_t* data, size_t len, long long* frames, long long* units_total, char* msg, size_t cap);
/* Host-only self test of the intra-block-copy rule of that plan: a block vector into a unit that is decoded earlier (forward = 0)
 * is accepted (returns the unit count), one into a unit that comes later (forward != 0) is rejected with -2. */
int av1r_debug_k3_ibc_selftest(int forward);
int av1r_probe_file(const char* path, av1r_stream_info* out);
int av1r_probe_buffer(const uint8_t* data, size_t len, av1r_stream_info* out);

#ifdef __cplusplus
}
#endif
#endif /* AV1R_H */

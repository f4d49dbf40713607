/* av1r stage-level and measurement entry points.
 *
 * Callable in isolation on caller-owned device planes: K8 film grain (av1r_stage_film_grain) and the K9 plane digest
 * (av1r_stage_plane_checksum) -- the two stages whose whole input is a frame.  K1-K7 consume the parser's work-lists, which only
 * exist inside the engine; they are isolated for parity the way libdav1d isolates them: av1r_config.inloop_filters (bit mask
 * 1 = deblock, 2 = CDEF, 4 = loop restoration) and av1r_config.apply_grain switch stages off so that the output after each stage
 * can be compared with the oracle run with the same switches (tests/test_decode_intra.py, tests/test_decode_inter.py), and
 * av1r_clip_profile times every stage separately (CUDA events between the stages of a serialised replay).
 * Plain pointers and sizes, no torch types.  All image pointers are DEVICE pointers; `stream` is a cudaStream_t passed as void*
 * (NULL = default stream).  Pitches are in bytes.  Samples are uint8 when bpc == 8, else uint16.
 */
#ifndef AV1R_STAGES_H
#define AV1R_STAGES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* film_grain_params() of the frame header (AV1 spec 5.9.30), already de-biased:
 * ar_coeffs_* are signed (coded value - 128), grain_scaling is 8..11, ar_coeff_shift 6..9. */
typedef struct av1r_film_grain_params {
    int apply_grain, grain_seed, update_grain;
    int num_y_points, point_y_value[16], point_y_scaling[16];
    int chroma_scaling_from_luma;
    int num_cb_points, point_cb_value[16], point_cb_scaling[16];
    int num_cr_points, point_cr_value[16], point_cr_scaling[16];
    int grain_scaling;
    int ar_coeff_lag;
    int ar_coeffs_y[24], ar_coeffs_cb[25], ar_coeffs_cr[25];
    int ar_coeff_shift;
    int grain_scale_shift;
    int cb_mult, cb_luma_mult, cb_offset, cr_mult, cr_luma_mult, cr_offset;
    int overlap_flag, clip_to_restricted_range;
} av1r_film_grain_params;

typedef struct av1r_frame_header_info {
    uint32_t struct_size;
    int tu_index;
    int frame_type, show_frame, showable_frame, show_existing_frame, frame_to_show_map_idx;
    int width, height, upscaled_width, bit_depth, subsampling_x, subsampling_y, mono_chrome;
    int matrix_coefficients;
    int refresh_frame_flags, order_hint, primary_ref_frame;
    int base_q_idx, tile_cols, tile_rows, use_128x128_superblock;
    int lf_level[4], cdef_enabled, cdef_bits, lr_type[3], tx_mode, reduced_tx_set;
    int header_bytes;
    av1r_film_grain_params film_grain;
    /* syntax switches of the frame / sequence header (tests assert that the golden streams really use them) */
    int error_resilient_mode, disable_cdf_update, disable_frame_end_update_cdf, enable_order_hint, coded_lossless;
    int segmentation_enabled, segmentation_update_map, segmentation_temporal_update, delta_q_present, delta_lf_present;
} av1r_frame_header_info;

/* Header-only scan of a list of temporal units (no pixel work, no GPU): the native stand-in for
 * the ffprobe child of /root/reference/internal/metadata/probe.go:145-153. */
int av1r_scan_headers(const uint8_t* const* tus, const size_t* lens, int n_tus,
                      av1r_frame_header_info* out, int cap, int* n);

/* K8 film grain synthesis: src planes (grain-free reconstruction) -> dst planes (display copy).
 * `scratch` is a device buffer of at least av1r_film_grain_scratch_bytes() bytes per call in
 * flight.  Launches 2 kernels on `stream`; returns without synchronising. */
size_t av1r_film_grain_scratch_bytes(void);
int av1r_stage_film_grain(const av1r_film_grain_params* p, int bpc, int w, int h,
                          int subsampling_x, int subsampling_y, int mono_chrome, int mc_identity,
                          const void* const src[3], const size_t src_pitch[3],
                          void* const dst[3], const size_t dst_pitch[3],
                          void* scratch, void* stream);

/* 64-bit plane checksum (sum over rows of a position-salted multiplicative hash): the
 * device-side digest returned per frame when parity_md5 is off.  out = device uint64_t*. */
int av1r_stage_plane_checksum(const void* src, size_t pitch, int w, int h, int bpc,
                              uint64_t* out_dev, void* stream);
/* host restatement of the same digest, for tests */
uint64_t av1r_plane_checksum_host(const void* src, size_t pitch, int w, int h, int bpc);

const char* av1r_stage_last_error(void);

/* ---- clip replay (measurement): parse once, then time the device path ----
 * av1r_clip_load runs the host parse of every temporal unit (GOP segments in parallel) and keeps the work-lists in pinned
 * host memory (untimed).  av1r_clip_decode replays the reconstruction of the whole clip: per frame one H2D copy of its
 * work-lists, then the kernels -- the region `value` in bench.py times (CUDA events, all streams fenced by the
 * start/stop events).  av1r_clip_profile replays serially on one stream with an event between the
 * stages and reports per-stage device time and launch counts (the live roofline input). */
struct av1r_ctx;
typedef struct av1r_clip av1r_clip;

typedef struct av1r_clip_info {
    uint32_t struct_size;
    int width, height, bit_depth;
    int64_t frames_decoded, frames_shown;
    double host_parse_ms;          /* sequential symbol parse, reported separately (north_star) */
    uint64_t worklist_bytes;       /* H2D bytes of all work-lists */
    uint64_t frame_bytes;          /* F: bytes of one frame (all planes, visible size) */
    uint64_t coded_samples;        /* A summed over frames */
    uint64_t coef_tokens;          /* non-zero coefficients summed over frames (4 B each = C) */
    uint64_t tx_blocks;            /* transform-block records summed over frames (32 B each) */
    uint64_t intra_samples;        /* samples reconstructed by the intra wavefront kernel */
    uint64_t inter_samples;        /* samples predicted by K2 (all planes), summed over frames */
    uint64_t inter_ref_samples;    /* reference samples those predictions are formed from (2x for compound): Rbar * inter_samples */
    uint64_t lr_frames, cdef_frames, deblock_frames, grain_frames;   /* frames on which each post-filter stage runs */
    uint64_t inter_blocks, obmc_neighbours;
    uint64_t tool_hist[24];        /* block counts per coding tool, order of TOOL_* in csrc/frame_state.h */
    /* the share of intra_samples / coded_samples / tx_blocks that belongs to frames without inter blocks (key and intra-only
     * frames): those run the whole-frame wavefront build of the intra kernel, a different kernel than the scattered units of inter frames */
    uint64_t intra_frame_samples, intra_frame_coded_samples, intra_frame_tx_blocks, intra_frames;
} av1r_clip_info;

enum { AV1R_ST_H2D = 0, AV1R_ST_ITX, AV1R_ST_INTRA, AV1R_ST_INTER, AV1R_ST_DEBLOCK, AV1R_ST_CDEF, AV1R_ST_LR, AV1R_ST_GRAIN,
       AV1R_ST_DIGEST, AV1R_ST_SUPERRES,
       AV1R_ST_INTRA_FRAME,   /* the intra kernel on frames without inter blocks (whole-frame wavefront); AV1R_ST_INTRA = inter frames' units */
       AV1R_ST_COUNT };
typedef struct av1r_stage_times {
    uint32_t struct_size;
    float ms[AV1R_ST_COUNT];       /* summed over the clip */
    int launches[AV1R_ST_COUNT];   /* kernel launches (copies for H2D) */
} av1r_stage_times;

int av1r_clip_load(struct av1r_ctx* ctx, const uint8_t* const* tus, const size_t* lens, int n_tus, av1r_clip** out);
int av1r_clip_info_get(const av1r_clip* clip, av1r_clip_info* out);
/* checksums: caller array of cap_frames*3 uint64 (display order); *n_frames = frames produced. */
int av1r_clip_decode(struct av1r_ctx* ctx, av1r_clip* clip, uint64_t* checksums, int cap_frames, int* n_frames, float* device_ms);
/* `passes` replays enqueued back to back between one pair of fencing events (no drain between passes: the first frames of pass
 * k + 1 overlap the last frames of pass k, as consecutive GOPs of a long file do in the streaming path); *device_ms covers all of
 * them.  checksums / *n_frames describe the first pass; a later pass whose digests differ makes the call fail with AV1R_EIO. */
int av1r_clip_decode_passes(struct av1r_ctx* ctx, av1r_clip* clip, int passes, uint64_t* checksums, int cap_frames, int* n_frames,
                            float* device_ms);
int av1r_clip_profile(struct av1r_ctx* ctx, av1r_clip* clip, av1r_stage_times* out);
void av1r_clip_free(av1r_clip* clip);
/* Default (0): every replay copies each frame's work-lists host -> device inside the timed pass ("from the first
 * work-list H2D enqueue").  1: upload once, replays read the lists from HBM (the kernel-only figure). */
int av1r_clip_set_resident(av1r_clip* clip, int resident);
/* Host-only statistics of a container (IVF / raw OBU / Matroska bytes): the parser-side fields of av1r_clip_info (tool histogram,
 * coded samples, frames per post-filter stage ...) without touching a GPU.  bench.py uses it to refuse a clip that lacks the tools
 * of the BASELINE config it stands for. */
int av1r_parse_stats(const uint8_t* data, size_t len, av1r_clip_info* out);

#ifdef __cplusplus
}
#endif
#endif

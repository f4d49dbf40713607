/* av1r stage-level entry points: one per reconstruction stage of SURVEY.md section 8(a2).
 *
 * These exist so that every CUDA kernel family can be parity-tested and timed in isolation
 * against the oracle (libdav1d with apply_grain / inloop_filters toggled), exactly the way
 * the whole path is: plain pointers and sizes, no torch types.  All image pointers are DEVICE
 * pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream).  Pitches are in
 * bytes.  Samples are uint8 when bpc == 8, else uint16.
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

#ifdef __cplusplus
}
#endif
#endif

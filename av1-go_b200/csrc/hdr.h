// Sequence / frame header state for the AV1 host parser (spec sections 5.5, 5.9).
#pragma once
#include <cstdint>
#include <cstring>

namespace av1r {

enum { KEY_FRAME = 0, INTER_FRAME = 1, INTRA_ONLY_FRAME = 2, SWITCH_FRAME = 3 };
enum { OBU_SEQUENCE_HEADER = 1, OBU_TEMPORAL_DELIMITER = 2, OBU_FRAME_HEADER = 3, OBU_TILE_GROUP = 4,
       OBU_METADATA = 5, OBU_FRAME = 6, OBU_REDUNDANT_FRAME_HEADER = 7, OBU_TILE_LIST = 8, OBU_PADDING = 15 };
enum { NUM_REF_FRAMES = 8, REFS_PER_FRAME = 7, PRIMARY_REF_NONE = 7, MAX_SEGMENTS = 8, SEG_LVL_MAX = 8,
       SEG_LVL_REF_FRAME = 5, SEG_LVL_SKIP = 6, SEG_LVL_GLOBALMV = 7 };
enum { INTRA_FRAME = 0, LAST_FRAME = 1, LAST2_FRAME = 2, LAST3_FRAME = 3, GOLDEN_FRAME = 4, BWDREF_FRAME = 5,
       ALTREF2_FRAME = 6, ALTREF_FRAME = 7 };
enum { RESTORE_NONE = 0, RESTORE_WIENER = 1, RESTORE_SGRPROJ = 2, RESTORE_SWITCHABLE = 3 };
enum { ONLY_4X4 = 0, TX_MODE_LARGEST = 1, TX_MODE_SELECT = 2 };
enum { GM_IDENTITY = 0, GM_TRANSLATION = 1, GM_ROTZOOM = 2, GM_AFFINE = 3 };
enum { INTERP_EIGHTTAP = 0, INTERP_SMOOTH = 1, INTERP_SHARP = 2, INTERP_BILINEAR = 3, INTERP_SWITCHABLE = 4 };
enum { MAX_TILE_COLS = 64, MAX_TILE_ROWS = 64 };

struct SeqHdr {
    int profile, still_picture, reduced_still_picture_header;
    int timing_info_present, equal_picture_interval, decoder_model_info_present;
    int buffer_delay_length_minus_1, buffer_removal_time_length_minus_1, frame_presentation_time_length_minus_1;
    int initial_display_delay_present, operating_points_cnt_minus_1;
    int operating_point_idc[32], decoder_model_present_for_this_op[32];
    int frame_width_bits, frame_height_bits, max_frame_width, max_frame_height;
    int frame_id_numbers_present, delta_frame_id_length_minus_2, additional_frame_id_length_minus_1;
    int use_128x128_superblock, enable_filter_intra, enable_intra_edge_filter;
    int enable_interintra_compound, enable_masked_compound, enable_warped_motion, enable_dual_filter;
    int enable_order_hint, enable_jnt_comp, enable_ref_frame_mvs;
    int seq_force_screen_content_tools, seq_force_integer_mv, order_hint_bits;
    int enable_superres, enable_cdef, enable_restoration;
    int bit_depth, mono_chrome, num_planes, color_primaries, transfer_characteristics, matrix_coefficients;
    int color_range, subsampling_x, subsampling_y, chroma_sample_position, separate_uv_delta_q;
    int film_grain_params_present;
    bool valid;
};

struct FilmGrainParams {
    int apply_grain, grain_seed, update_grain;
    int num_y_points, point_y_value[16], point_y_scaling[16];
    int chroma_scaling_from_luma;
    int num_cb_points, point_cb_value[16], point_cb_scaling[16];
    int num_cr_points, point_cr_value[16], point_cr_scaling[16];
    int grain_scaling;  // scaling_shift 8..11
    int ar_coeff_lag;
    int ar_coeffs_y[24], ar_coeffs_cb[25], ar_coeffs_cr[25];  // signed (minus 128 applied)
    int ar_coeff_shift;  // 6..9
    int grain_scale_shift;
    int cb_mult, cb_luma_mult, cb_offset, cr_mult, cr_luma_mult, cr_offset;
    int overlap_flag, clip_to_restricted_range;
};

struct SegmentationParams {
    int enabled, update_map, temporal_update, update_data;
    int feature_enabled[MAX_SEGMENTS][SEG_LVL_MAX];
    int feature_data[MAX_SEGMENTS][SEG_LVL_MAX];
    int seg_id_pre_skip, last_active_seg_id;
};

struct LoopFilterParams {
    int level[4], sharpness, delta_enabled, delta_update;
    int ref_deltas[8], mode_deltas[2];
};

struct FrameHdr {
    int show_existing_frame, frame_to_show_map_idx;
    int frame_type, frame_is_intra, show_frame, showable_frame, error_resilient_mode;
    int disable_cdf_update, allow_screen_content_tools, force_integer_mv;
    int current_frame_id, frame_size_override_flag, order_hint, primary_ref_frame;
    int refresh_frame_flags;
    int ref_order_hint[NUM_REF_FRAMES];
    int frame_width, frame_height, upscaled_width, render_width, render_height;
    int use_superres, superres_denom;
    int mi_cols, mi_rows;
    int allow_intrabc;
    int frame_refs_short_signaling;
    int ref_frame_idx[REFS_PER_FRAME];
    int allow_high_precision_mv, is_filter_switchable, interpolation_filter, is_motion_mode_switchable;
    int use_ref_frame_mvs;
    int order_hints[8];           // OrderHints[refFrame]
    int ref_frame_sign_bias[8];
    int disable_frame_end_update_cdf;
    // tile info
    int tile_cols, tile_rows, tile_cols_log2, tile_rows_log2;
    int mi_col_starts[MAX_TILE_COLS + 1], mi_row_starts[MAX_TILE_ROWS + 1];
    int context_update_tile_id, tile_size_bytes;
    // quant
    int base_q_idx, delta_q_y_dc, delta_q_u_dc, delta_q_u_ac, delta_q_v_dc, delta_q_v_ac;
    int using_qmatrix, qm_y, qm_u, qm_v;
    SegmentationParams seg;
    int delta_q_present, delta_q_res, delta_lf_present, delta_lf_res, delta_lf_multi;
    int coded_lossless, all_lossless;
    int lossless_array[MAX_SEGMENTS], seg_qm_level[3][MAX_SEGMENTS];
    int qidx_seg[MAX_SEGMENTS];
    LoopFilterParams lf;
    int cdef_damping, cdef_bits, cdef_y_pri[8], cdef_y_sec[8], cdef_uv_pri[8], cdef_uv_sec[8];
    int enable_cdef_frame;   // cdef not bypassed for this frame
    int lr_type[3], uses_lr, lr_unit_shift, lr_uv_shift, lr_size[3];
    int tx_mode, reference_select, skip_mode_allowed, skip_mode_present, skip_mode_frame[2];
    int allow_warped_motion, reduced_tx_set;
    int gm_type[8];
    int32_t gm_params[8][6];
    FilmGrainParams fg;
    // bookkeeping
    int temporal_id, spatial_id;
    size_t header_bytes;   // bytes consumed by the uncompressed header (byte aligned, OBU_FRAME case)
};

}  // namespace av1r

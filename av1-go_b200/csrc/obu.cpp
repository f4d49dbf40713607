// OBU / sequence header / frame header parsing, written from the AV1 bitstream syntax
// (spec 5.3-5.11).  This is the "probe" half of the graft: it yields the geometry the
// reference gets from ffprobe (/root/reference/internal/metadata/probe.go:125-204) and the
// per-frame parameters every reconstruction kernel consumes.
#include "obu.h"

#include <algorithm>
#include <cstdio>

namespace av1r {

static const int kDefaultRefDeltas[8] = {1, 0, 0, 0, -1, 0, -1, -1};
static const int kSegFeatureBits[8] = {8, 6, 6, 6, 6, 3, 0, 0};
static const int kSegFeatureSigned[8] = {1, 1, 1, 1, 1, 0, 0, 0};
static const int kSegFeatureMax[8] = {255, 63, 63, 63, 63, 7, 0, 0};

static void default_gm(int32_t gm[8][6]) {
    for (int r = 0; r < 8; r++)
        for (int i = 0; i < 6; i++) gm[r][i] = (i % 3 == 2) ? (1 << 16) : 0;
}

HeaderParser::HeaderParser() {
    memset(&seq, 0, sizeof(seq));
    for (auto& r : refs) {
        default_gm(r.saved_gm_params);
        memcpy(r.lf_ref_deltas, kDefaultRefDeltas, sizeof(kDefaultRefDeltas));
        r.lf_mode_deltas[0] = r.lf_mode_deltas[1] = 0;
        memset(&r.seg, 0, sizeof(r.seg));
        memset(&r.fg, 0, sizeof(r.fg));
    }
    default_gm(prev_gm_params);
}

bool HeaderParser::split_obus(const uint8_t* data, size_t len, std::vector<ObuUnit>& out) {
    size_t pos = 0;
    while (pos < len) {
        BitReader br(data + pos, len - pos);
        if (br.f(1)) return fail("obu forbidden bit set");
        ObuUnit u;
        u.type = br.f(4);
        int ext = br.f(1);
        int has_size = br.f(1);
        br.f(1);
        u.temporal_id = u.spatial_id = 0;
        if (ext) {
            u.temporal_id = br.f(3);
            u.spatial_id = br.f(2);
            br.f(3);
        }
        size_t sz;
        if (has_size) {
            sz = (size_t)br.leb128();
        } else {
            sz = len - pos - 1 - ext;
        }
        size_t hdr = br.byte_pos();
        if (br.err || pos + hdr + sz > len) return fail("obu size exceeds temporal unit");
        u.data = data + pos + hdr;
        u.size = sz;
        out.push_back(u);
        pos += hdr + sz;
    }
    return true;
}

bool HeaderParser::parse_sequence_header(const uint8_t* d, size_t n) {
    BitReader br(d, n);
    SeqHdr s;
    memset(&s, 0, sizeof(s));
    s.profile = br.f(3);
    s.still_picture = br.f(1);
    s.reduced_still_picture_header = br.f(1);
    if (s.reduced_still_picture_header) {
        br.f(5);  // seq_level_idx[0]
    } else {
        s.timing_info_present = br.f(1);
        if (s.timing_info_present) {
            br.f(32);
            br.f(32);
            s.equal_picture_interval = br.f(1);
            if (s.equal_picture_interval) br.uvlc();
            s.decoder_model_info_present = br.f(1);
            if (s.decoder_model_info_present) {
                s.buffer_delay_length_minus_1 = br.f(5);
                br.f(32);
                s.buffer_removal_time_length_minus_1 = br.f(5);
                s.frame_presentation_time_length_minus_1 = br.f(5);
            }
        }
        s.initial_display_delay_present = br.f(1);
        s.operating_points_cnt_minus_1 = br.f(5);
        for (int i = 0; i <= s.operating_points_cnt_minus_1; i++) {
            s.operating_point_idc[i] = br.f(12);
            int lvl = br.f(5);
            if (lvl > 7) br.f(1);
            if (s.decoder_model_info_present) {
                s.decoder_model_present_for_this_op[i] = br.f(1);
                if (s.decoder_model_present_for_this_op[i]) {
                    int nb = s.buffer_delay_length_minus_1 + 1;
                    br.f(nb);
                    br.f(nb);
                    br.f(1);
                }
            }
            if (s.initial_display_delay_present) {
                if (br.f(1)) br.f(4);
            }
        }
    }
    s.frame_width_bits = br.f(4) + 1;
    s.frame_height_bits = br.f(4) + 1;
    s.max_frame_width = br.f(s.frame_width_bits) + 1;
    s.max_frame_height = br.f(s.frame_height_bits) + 1;
    if (!s.reduced_still_picture_header) s.frame_id_numbers_present = br.f(1);
    if (s.frame_id_numbers_present) {
        s.delta_frame_id_length_minus_2 = br.f(4);
        s.additional_frame_id_length_minus_1 = br.f(3);
    }
    s.use_128x128_superblock = br.f(1);
    s.enable_filter_intra = br.f(1);
    s.enable_intra_edge_filter = br.f(1);
    if (s.reduced_still_picture_header) {
        s.seq_force_screen_content_tools = 2;
        s.seq_force_integer_mv = 2;
    } else {
        s.enable_interintra_compound = br.f(1);
        s.enable_masked_compound = br.f(1);
        s.enable_warped_motion = br.f(1);
        s.enable_dual_filter = br.f(1);
        s.enable_order_hint = br.f(1);
        if (s.enable_order_hint) {
            s.enable_jnt_comp = br.f(1);
            s.enable_ref_frame_mvs = br.f(1);
        }
        if (br.f(1)) s.seq_force_screen_content_tools = 2;
        else s.seq_force_screen_content_tools = br.f(1);
        if (s.seq_force_screen_content_tools > 0) {
            if (br.f(1)) s.seq_force_integer_mv = 2;
            else s.seq_force_integer_mv = br.f(1);
        } else {
            s.seq_force_integer_mv = 2;
        }
        if (s.enable_order_hint) s.order_hint_bits = br.f(3) + 1;
    }
    s.enable_superres = br.f(1);
    s.enable_cdef = br.f(1);
    s.enable_restoration = br.f(1);
    // color_config
    int high_bitdepth = br.f(1);
    if (s.profile == 2 && high_bitdepth) s.bit_depth = br.f(1) ? 12 : 10;
    else s.bit_depth = high_bitdepth ? 10 : 8;
    s.mono_chrome = (s.profile == 1) ? 0 : br.f(1);
    s.num_planes = s.mono_chrome ? 1 : 3;
    if (br.f(1)) {
        s.color_primaries = br.f(8);
        s.transfer_characteristics = br.f(8);
        s.matrix_coefficients = br.f(8);
    } else {
        s.color_primaries = s.transfer_characteristics = s.matrix_coefficients = 2;
    }
    if (s.mono_chrome) {
        s.color_range = br.f(1);
        s.subsampling_x = s.subsampling_y = 1;
        s.separate_uv_delta_q = 0;
    } else if (s.color_primaries == 1 && s.transfer_characteristics == 13 && s.matrix_coefficients == 0) {
        s.color_range = 1;
        s.subsampling_x = s.subsampling_y = 0;
        s.separate_uv_delta_q = br.f(1);
    } else {
        s.color_range = br.f(1);
        if (s.profile == 0) s.subsampling_x = s.subsampling_y = 1;
        else if (s.profile == 1) s.subsampling_x = s.subsampling_y = 0;
        else if (s.bit_depth == 12) {
            s.subsampling_x = br.f(1);
            s.subsampling_y = s.subsampling_x ? br.f(1) : 0;
        } else {
            s.subsampling_x = 1;
            s.subsampling_y = 0;
        }
        if (s.subsampling_x && s.subsampling_y) s.chroma_sample_position = br.f(2);
        s.separate_uv_delta_q = br.f(1);
    }
    s.film_grain_params_present = br.f(1);
    if (br.err) return fail("truncated sequence header");
    s.valid = true;
    seq = s;
    return true;
}

void HeaderParser::setup_past_independence(FrameHdr& fh) {
    memset(fh.seg.feature_data, 0, sizeof(fh.seg.feature_data));
    memset(fh.seg.feature_enabled, 0, sizeof(fh.seg.feature_enabled));
    memcpy(fh.lf.ref_deltas, kDefaultRefDeltas, sizeof(kDefaultRefDeltas));
    fh.lf.mode_deltas[0] = fh.lf.mode_deltas[1] = 0;
    default_gm(prev_gm_params);
}

void HeaderParser::load_previous(FrameHdr& fh) {
    const RefHdrState& r = refs[fh.ref_frame_idx[fh.primary_ref_frame]];
    memcpy(prev_gm_params, r.saved_gm_params, sizeof(prev_gm_params));
    memcpy(fh.lf.ref_deltas, r.lf_ref_deltas, sizeof(r.lf_ref_deltas));
    memcpy(fh.lf.mode_deltas, r.lf_mode_deltas, sizeof(r.lf_mode_deltas));
    memcpy(fh.seg.feature_data, r.seg.feature_data, sizeof(r.seg.feature_data));
    memcpy(fh.seg.feature_enabled, r.seg.feature_enabled, sizeof(r.seg.feature_enabled));
}

void HeaderParser::superres_params(BitReader& br, FrameHdr& fh) {
    fh.use_superres = seq.enable_superres ? br.f(1) : 0;
    if (fh.use_superres) fh.superres_denom = br.f(3) + 9;
    else fh.superres_denom = 8;
    fh.upscaled_width = fh.frame_width;
    fh.frame_width = (fh.upscaled_width * 8 + (fh.superres_denom / 2)) / fh.superres_denom;
}

void HeaderParser::compute_image_size(FrameHdr& fh) {
    fh.mi_cols = 2 * ((fh.frame_width + 7) >> 3);
    fh.mi_rows = 2 * ((fh.frame_height + 7) >> 3);
}

bool HeaderParser::frame_size(BitReader& br, FrameHdr& fh) {
    if (fh.frame_size_override_flag) {
        fh.frame_width = br.f(seq.frame_width_bits) + 1;
        fh.frame_height = br.f(seq.frame_height_bits) + 1;
    } else {
        fh.frame_width = seq.max_frame_width;
        fh.frame_height = seq.max_frame_height;
    }
    superres_params(br, fh);
    compute_image_size(fh);
    return true;
}

void HeaderParser::render_size(BitReader& br, FrameHdr& fh) {
    if (br.f(1)) {
        fh.render_width = br.f(16) + 1;
        fh.render_height = br.f(16) + 1;
    } else {
        fh.render_width = fh.upscaled_width;
        fh.render_height = fh.frame_height;
    }
}

bool HeaderParser::frame_size_with_refs(BitReader& br, FrameHdr& fh) {
    for (int i = 0; i < REFS_PER_FRAME; i++) {
        if (br.f(1)) {
            const RefHdrState& r = refs[fh.ref_frame_idx[i]];
            fh.upscaled_width = r.upscaled_width;
            fh.frame_width = fh.upscaled_width;
            fh.frame_height = r.frame_height;
            fh.render_width = r.render_width;
            fh.render_height = r.render_height;
            superres_params(br, fh);
            compute_image_size(fh);
            return true;
        }
    }
    frame_size(br, fh);
    render_size(br, fh);
    return true;
}

void HeaderParser::set_frame_refs(FrameHdr& fh, int last_frame_idx, int gold_frame_idx) {
    int* idx = fh.ref_frame_idx;
    for (int i = 0; i < REFS_PER_FRAME; i++) idx[i] = -1;
    idx[0] = last_frame_idx;
    idx[GOLDEN_FRAME - LAST_FRAME] = gold_frame_idx;
    int used[NUM_REF_FRAMES] = {0};
    used[last_frame_idx] = 1;
    used[gold_frame_idx] = 1;
    int cur = 1 << (seq.order_hint_bits - 1);
    int shifted[NUM_REF_FRAMES];
    for (int i = 0; i < NUM_REF_FRAMES; i++) shifted[i] = cur + get_relative_dist(refs[i].order_hint, fh.order_hint);
    auto latest_backward = [&]() {
        int ref = -1, best = 0;
        for (int i = 0; i < NUM_REF_FRAMES; i++)
            if (!used[i] && shifted[i] >= cur && (ref < 0 || shifted[i] >= best)) { ref = i; best = shifted[i]; }
        return ref;
    };
    auto earliest_backward = [&]() {
        int ref = -1, best = 0;
        for (int i = 0; i < NUM_REF_FRAMES; i++)
            if (!used[i] && shifted[i] >= cur && (ref < 0 || shifted[i] < best)) { ref = i; best = shifted[i]; }
        return ref;
    };
    auto latest_forward = [&]() {
        int ref = -1, best = 0;
        for (int i = 0; i < NUM_REF_FRAMES; i++)
            if (!used[i] && shifted[i] < cur && (ref < 0 || shifted[i] >= best)) { ref = i; best = shifted[i]; }
        return ref;
    };
    int ref = latest_backward();
    if (ref >= 0) { idx[ALTREF_FRAME - LAST_FRAME] = ref; used[ref] = 1; }
    ref = earliest_backward();
    if (ref >= 0) { idx[BWDREF_FRAME - LAST_FRAME] = ref; used[ref] = 1; }
    ref = earliest_backward();
    if (ref >= 0) { idx[ALTREF2_FRAME - LAST_FRAME] = ref; used[ref] = 1; }
    static const int list[5] = {LAST2_FRAME, LAST3_FRAME, BWDREF_FRAME, ALTREF2_FRAME, ALTREF_FRAME};
    for (int i = 0; i < 5; i++) {
        int rf = list[i];
        if (idx[rf - LAST_FRAME] < 0) {
            ref = latest_forward();
            if (ref >= 0) { idx[rf - LAST_FRAME] = ref; used[ref] = 1; }
        }
    }
    ref = -1;
    int best = 0;
    for (int i = 0; i < NUM_REF_FRAMES; i++)
        if (ref < 0 || shifted[i] < best) { ref = i; best = shifted[i]; }
    for (int i = 0; i < REFS_PER_FRAME; i++)
        if (idx[i] < 0) idx[i] = ref;
}

static int tile_log2(int blk, int target) {
    int k = 0;
    while ((blk << k) < target) k++;
    return k;
}

bool HeaderParser::tile_info(BitReader& br, FrameHdr& fh) {
    int sb128 = seq.use_128x128_superblock;
    int sbCols = sb128 ? (fh.mi_cols + 31) >> 5 : (fh.mi_cols + 15) >> 4;
    int sbRows = sb128 ? (fh.mi_rows + 31) >> 5 : (fh.mi_rows + 15) >> 4;
    int sbShift = sb128 ? 5 : 4;
    int sbSize = sbShift + 2;
    int maxTileWidthSb = 4096 >> sbSize;
    int maxTileAreaSb = (4096 * 2304) >> (2 * sbSize);
    int minLog2TileCols = tile_log2(maxTileWidthSb, sbCols);
    int maxLog2TileCols = tile_log2(1, std::min(sbCols, (int)MAX_TILE_COLS));
    int maxLog2TileRows = tile_log2(1, std::min(sbRows, (int)MAX_TILE_ROWS));
    int minLog2Tiles = std::max(minLog2TileCols, tile_log2(maxTileAreaSb, sbRows * sbCols));
    int uniform = br.f(1);
    if (uniform) {
        fh.tile_cols_log2 = minLog2TileCols;
        while (fh.tile_cols_log2 < maxLog2TileCols) {
            if (br.f(1)) fh.tile_cols_log2++;
            else break;
        }
        int tileWidthSb = (sbCols + (1 << fh.tile_cols_log2) - 1) >> fh.tile_cols_log2;
        int i = 0;
        for (int startSb = 0; startSb < sbCols; startSb += tileWidthSb) {
            if (i >= MAX_TILE_COLS) return fail("too many tile columns");
            fh.mi_col_starts[i++] = startSb << sbShift;
        }
        fh.mi_col_starts[i] = fh.mi_cols;
        fh.tile_cols = i;
        int minLog2TileRows = std::max(minLog2Tiles - fh.tile_cols_log2, 0);
        fh.tile_rows_log2 = minLog2TileRows;
        while (fh.tile_rows_log2 < maxLog2TileRows) {
            if (br.f(1)) fh.tile_rows_log2++;
            else break;
        }
        int tileHeightSb = (sbRows + (1 << fh.tile_rows_log2) - 1) >> fh.tile_rows_log2;
        i = 0;
        for (int startSb = 0; startSb < sbRows; startSb += tileHeightSb) {
            if (i >= MAX_TILE_ROWS) return fail("too many tile rows");
            fh.mi_row_starts[i++] = startSb << sbShift;
        }
        fh.mi_row_starts[i] = fh.mi_rows;
        fh.tile_rows = i;
    } else {
        int widestTileSb = 0, startSb = 0, i = 0;
        for (; startSb < sbCols; i++) {
            if (i >= MAX_TILE_COLS) return fail("too many tile columns");
            fh.mi_col_starts[i] = startSb << sbShift;
            int maxWidth = std::min(sbCols - startSb, maxTileWidthSb);
            int sizeSb = br.ns(maxWidth) + 1;
            widestTileSb = std::max(sizeSb, widestTileSb);
            startSb += sizeSb;
        }
        fh.mi_col_starts[i] = fh.mi_cols;
        fh.tile_cols = i;
        fh.tile_cols_log2 = tile_log2(1, fh.tile_cols);
        if (minLog2Tiles > 0) maxTileAreaSb = (sbRows * sbCols) >> (minLog2Tiles + 1);
        else maxTileAreaSb = sbRows * sbCols;
        int maxTileHeightSb = std::max(maxTileAreaSb / widestTileSb, 1);
        startSb = 0;
        for (i = 0; startSb < sbRows; i++) {
            if (i >= MAX_TILE_ROWS) return fail("too many tile rows");
            fh.mi_row_starts[i] = startSb << sbShift;
            int maxHeight = std::min(sbRows - startSb, maxTileHeightSb);
            int sizeSb = br.ns(maxHeight) + 1;
            startSb += sizeSb;
        }
        fh.mi_row_starts[i] = fh.mi_rows;
        fh.tile_rows = i;
        fh.tile_rows_log2 = tile_log2(1, fh.tile_rows);
    }
    if (fh.tile_cols_log2 > 0 || fh.tile_rows_log2 > 0) {
        fh.context_update_tile_id = br.f(fh.tile_rows_log2 + fh.tile_cols_log2);
        fh.tile_size_bytes = br.f(2) + 1;
    } else {
        fh.context_update_tile_id = 0;
        fh.tile_size_bytes = 4;
    }
    return true;
}

int HeaderParser::read_delta_q(BitReader& br) {
    if (br.f(1)) return br.su(7);
    return 0;
}

void HeaderParser::quantization_params(BitReader& br, FrameHdr& fh) {
    fh.base_q_idx = br.f(8);
    fh.delta_q_y_dc = read_delta_q(br);
    if (seq.num_planes > 1) {
        int diff_uv_delta = seq.separate_uv_delta_q ? br.f(1) : 0;
        fh.delta_q_u_dc = read_delta_q(br);
        fh.delta_q_u_ac = read_delta_q(br);
        if (diff_uv_delta) {
            fh.delta_q_v_dc = read_delta_q(br);
            fh.delta_q_v_ac = read_delta_q(br);
        } else {
            fh.delta_q_v_dc = fh.delta_q_u_dc;
            fh.delta_q_v_ac = fh.delta_q_u_ac;
        }
    } else {
        fh.delta_q_u_dc = fh.delta_q_u_ac = fh.delta_q_v_dc = fh.delta_q_v_ac = 0;
    }
    fh.using_qmatrix = br.f(1);
    if (fh.using_qmatrix) {
        fh.qm_y = br.f(4);
        fh.qm_u = br.f(4);
        fh.qm_v = seq.separate_uv_delta_q ? br.f(4) : fh.qm_u;
    } else {
        fh.qm_y = fh.qm_u = fh.qm_v = 15;
    }
}

void HeaderParser::segmentation_params(BitReader& br, FrameHdr& fh) {
    SegmentationParams& s = fh.seg;
    s.enabled = br.f(1);
    s.update_map = s.temporal_update = s.update_data = 0;
    if (s.enabled) {
        if (fh.primary_ref_frame == PRIMARY_REF_NONE) {
            s.update_map = 1;
            s.temporal_update = 0;
            s.update_data = 1;
        } else {
            s.update_map = br.f(1);
            if (s.update_map) s.temporal_update = br.f(1);
            s.update_data = br.f(1);
        }
        if (s.update_data) {
            for (int i = 0; i < MAX_SEGMENTS; i++)
                for (int j = 0; j < SEG_LVL_MAX; j++) {
                    int en = br.f(1);
                    int clipped = 0;
                    s.feature_enabled[i][j] = en;
                    if (en) {
                        int bits = kSegFeatureBits[j], lim = kSegFeatureMax[j];
                        if (kSegFeatureSigned[j]) {
                            int v = br.su(1 + bits);
                            clipped = std::max(-lim, std::min(lim, v));
                        } else {
                            int v = bits ? (int)br.f(bits) : 0;
                            clipped = std::max(0, std::min(lim, v));
                        }
                    }
                    s.feature_data[i][j] = clipped;
                }
        }
    } else {
        memset(s.feature_enabled, 0, sizeof(s.feature_enabled));
        memset(s.feature_data, 0, sizeof(s.feature_data));
    }
    s.seg_id_pre_skip = 0;
    s.last_active_seg_id = 0;
    for (int i = 0; i < MAX_SEGMENTS; i++)
        for (int j = 0; j < SEG_LVL_MAX; j++)
            if (s.feature_enabled[i][j]) {
                s.last_active_seg_id = i;
                if (j >= SEG_LVL_REF_FRAME) s.seg_id_pre_skip = 1;
            }
}

int get_qidx(const FrameHdr& fh, int ignore_deltas, int segment_id, int current_q_index) {
    if (fh.seg.enabled && fh.seg.feature_enabled[segment_id][0]) {
        int data = fh.seg.feature_data[segment_id][0];
        int q = fh.base_q_idx + data;
        if (!ignore_deltas && fh.delta_q_present) q = current_q_index + data;
        return std::max(0, std::min(255, q));
    }
    if (!ignore_deltas && fh.delta_q_present) return current_q_index;
    return fh.base_q_idx;
}

void HeaderParser::loop_filter_params(BitReader& br, FrameHdr& fh) {
    LoopFilterParams& lf = fh.lf;
    lf.level[0] = lf.level[1] = lf.level[2] = lf.level[3] = 0;
    lf.sharpness = 0;
    lf.delta_enabled = lf.delta_update = 0;
    if (fh.coded_lossless || fh.allow_intrabc) {
        memcpy(lf.ref_deltas, kDefaultRefDeltas, sizeof(kDefaultRefDeltas));
        lf.mode_deltas[0] = lf.mode_deltas[1] = 0;
        return;
    }
    lf.level[0] = br.f(6);
    lf.level[1] = br.f(6);
    if (seq.num_planes > 1 && (lf.level[0] || lf.level[1])) {
        lf.level[2] = br.f(6);
        lf.level[3] = br.f(6);
    }
    lf.sharpness = br.f(3);
    lf.delta_enabled = br.f(1);
    if (lf.delta_enabled) {
        lf.delta_update = br.f(1);
        if (lf.delta_update) {
            for (int i = 0; i < 8; i++)
                if (br.f(1)) lf.ref_deltas[i] = br.su(7);
            for (int i = 0; i < 2; i++)
                if (br.f(1)) lf.mode_deltas[i] = br.su(7);
        }
    }
}

void HeaderParser::cdef_params(BitReader& br, FrameHdr& fh) {
    fh.enable_cdef_frame = 0;
    fh.cdef_bits = 0;
    fh.cdef_damping = 3;
    memset(fh.cdef_y_pri, 0, sizeof(fh.cdef_y_pri));
    memset(fh.cdef_y_sec, 0, sizeof(fh.cdef_y_sec));
    memset(fh.cdef_uv_pri, 0, sizeof(fh.cdef_uv_pri));
    memset(fh.cdef_uv_sec, 0, sizeof(fh.cdef_uv_sec));
    if (fh.coded_lossless || fh.allow_intrabc || !seq.enable_cdef) return;
    fh.enable_cdef_frame = 1;
    fh.cdef_damping = br.f(2) + 3;
    fh.cdef_bits = br.f(2);
    for (int i = 0; i < (1 << fh.cdef_bits); i++) {
        fh.cdef_y_pri[i] = br.f(4);
        fh.cdef_y_sec[i] = br.f(2);
        if (fh.cdef_y_sec[i] == 3) fh.cdef_y_sec[i]++;
        if (seq.num_planes > 1) {
            fh.cdef_uv_pri[i] = br.f(4);
            fh.cdef_uv_sec[i] = br.f(2);
            if (fh.cdef_uv_sec[i] == 3) fh.cdef_uv_sec[i]++;
        }
    }
}

void HeaderParser::lr_params(BitReader& br, FrameHdr& fh) {
    fh.lr_type[0] = fh.lr_type[1] = fh.lr_type[2] = RESTORE_NONE;
    fh.uses_lr = 0;
    fh.lr_unit_shift = fh.lr_uv_shift = 0;
    fh.lr_size[0] = fh.lr_size[1] = fh.lr_size[2] = 64;
    if (fh.all_lossless || fh.allow_intrabc || !seq.enable_restoration) return;
    static const int remap[4] = {RESTORE_NONE, RESTORE_SWITCHABLE, RESTORE_WIENER, RESTORE_SGRPROJ};
    int usesChromaLr = 0;
    for (int i = 0; i < seq.num_planes; i++) {
        fh.lr_type[i] = remap[br.f(2)];
        if (fh.lr_type[i] != RESTORE_NONE) {
            fh.uses_lr = 1;
            if (i > 0) usesChromaLr = 1;
        }
    }
    if (fh.uses_lr) {
        if (seq.use_128x128_superblock) {
            fh.lr_unit_shift = br.f(1) + 1;
        } else {
            fh.lr_unit_shift = br.f(1);
            if (fh.lr_unit_shift) fh.lr_unit_shift += br.f(1);
        }
        if (seq.subsampling_x && seq.subsampling_y && usesChromaLr) fh.lr_uv_shift = br.f(1);
        fh.lr_size[0] = 256 >> (2 - fh.lr_unit_shift);
        fh.lr_size[1] = fh.lr_size[2] = fh.lr_size[0] >> fh.lr_uv_shift;
    }
}

void HeaderParser::skip_mode_params(BitReader& br, FrameHdr& fh) {
    fh.skip_mode_allowed = 0;
    fh.skip_mode_frame[0] = fh.skip_mode_frame[1] = 0;
    if (!(fh.frame_is_intra || !fh.reference_select || !seq.enable_order_hint)) {
        int forwardIdx = -1, backwardIdx = -1, forwardHint = 0, backwardHint = 0;
        for (int i = 0; i < REFS_PER_FRAME; i++) {
            int refHint = refs[fh.ref_frame_idx[i]].order_hint;
            if (get_relative_dist(refHint, fh.order_hint) < 0) {
                if (forwardIdx < 0 || get_relative_dist(refHint, forwardHint) > 0) { forwardIdx = i; forwardHint = refHint; }
            } else if (get_relative_dist(refHint, fh.order_hint) > 0) {
                if (backwardIdx < 0 || get_relative_dist(refHint, backwardHint) < 0) { backwardIdx = i; backwardHint = refHint; }
            }
        }
        if (forwardIdx < 0) {
            fh.skip_mode_allowed = 0;
        } else if (backwardIdx >= 0) {
            fh.skip_mode_allowed = 1;
            fh.skip_mode_frame[0] = LAST_FRAME + std::min(forwardIdx, backwardIdx);
            fh.skip_mode_frame[1] = LAST_FRAME + std::max(forwardIdx, backwardIdx);
        } else {
            int secondForwardIdx = -1, secondForwardHint = 0;
            for (int i = 0; i < REFS_PER_FRAME; i++) {
                int refHint = refs[fh.ref_frame_idx[i]].order_hint;
                if (get_relative_dist(refHint, forwardHint) < 0) {
                    if (secondForwardIdx < 0 || get_relative_dist(refHint, secondForwardHint) > 0) {
                        secondForwardIdx = i;
                        secondForwardHint = refHint;
                    }
                }
            }
            if (secondForwardIdx >= 0) {
                fh.skip_mode_allowed = 1;
                fh.skip_mode_frame[0] = LAST_FRAME + std::min(forwardIdx, secondForwardIdx);
                fh.skip_mode_frame[1] = LAST_FRAME + std::max(forwardIdx, secondForwardIdx);
            }
        }
    }
    fh.skip_mode_present = fh.skip_mode_allowed ? br.f(1) : 0;
}

static int inverse_recenter(int r, int v) {
    if (v > 2 * r) return v;
    if (v & 1) return r - ((v + 1) >> 1);
    return r + (v >> 1);
}

static int decode_subexp(BitReader& br, int numSyms) {
    int i = 0, mk = 0, k = 3;
    while (true) {
        int b2 = i ? k + i - 1 : k;
        int a = 1 << b2;
        if (numSyms <= mk + 3 * a) {
            return (int)br.ns(numSyms - mk) + mk;
        }
        if (br.f(1)) {
            i++;
            mk += a;
        } else {
            return (int)br.f(b2) + mk;
        }
        if (br.err) return 0;
    }
}

static int decode_unsigned_subexp_with_ref(BitReader& br, int mx, int r) {
    int v = decode_subexp(br, mx);
    if ((r << 1) <= mx) return inverse_recenter(r, v);
    return mx - 1 - inverse_recenter(mx - 1 - r, v);
}

static int decode_signed_subexp_with_ref(BitReader& br, int low, int high, int r) {
    return decode_unsigned_subexp_with_ref(br, high - low, r - low) + low;
}

void HeaderParser::read_global_param(BitReader& br, FrameHdr& fh, int type, int ref, int idx) {
    int absBits = 12, precBits = 15;
    if (idx < 2) {
        if (type == GM_TRANSLATION) {
            absBits = 9 - !fh.allow_high_precision_mv;
            precBits = 3 - !fh.allow_high_precision_mv;
        } else {
            absBits = 12;
            precBits = 6;
        }
    }
    int precDiff = 16 - precBits;
    int round = (idx % 3) == 2 ? (1 << 16) : 0;
    int sub = (idx % 3) == 2 ? (1 << precBits) : 0;
    int mx = 1 << absBits;
    int r = (prev_gm_params[ref][idx] >> precDiff) - sub;
    fh.gm_params[ref][idx] = (decode_signed_subexp_with_ref(br, -mx, mx + 1, r) << precDiff) + round;
}

void HeaderParser::global_motion_params(BitReader& br, FrameHdr& fh) {
    for (int ref = LAST_FRAME; ref <= ALTREF_FRAME; ref++) {
        fh.gm_type[ref] = GM_IDENTITY;
        for (int i = 0; i < 6; i++) fh.gm_params[ref][i] = (i % 3 == 2) ? (1 << 16) : 0;
    }
    fh.gm_type[0] = GM_IDENTITY;
    for (int i = 0; i < 6; i++) fh.gm_params[0][i] = (i % 3 == 2) ? (1 << 16) : 0;
    if (fh.frame_is_intra) return;
    for (int ref = LAST_FRAME; ref <= ALTREF_FRAME; ref++) {
        int type;
        if (br.f(1)) {
            if (br.f(1)) type = GM_ROTZOOM;
            else type = br.f(1) ? GM_TRANSLATION : GM_AFFINE;
        } else {
            type = GM_IDENTITY;
        }
        fh.gm_type[ref] = type;
        if (type >= GM_ROTZOOM) {
            read_global_param(br, fh, type, ref, 2);
            read_global_param(br, fh, type, ref, 3);
            if (type == GM_AFFINE) {
                read_global_param(br, fh, type, ref, 4);
                read_global_param(br, fh, type, ref, 5);
            } else {
                fh.gm_params[ref][4] = -fh.gm_params[ref][3];
                fh.gm_params[ref][5] = fh.gm_params[ref][2];
            }
        }
        if (type >= GM_TRANSLATION) {
            read_global_param(br, fh, type, ref, 0);
            read_global_param(br, fh, type, ref, 1);
        }
    }
}

void HeaderParser::film_grain_params(BitReader& br, FrameHdr& fh) {
    FilmGrainParams& g = fh.fg;
    if (!seq.film_grain_params_present || (!fh.show_frame && !fh.showable_frame)) {
        memset(&g, 0, sizeof(g));
        return;
    }
    int apply = br.f(1);
    if (!apply) {
        memset(&g, 0, sizeof(g));
        return;
    }
    int seed = br.f(16);
    int update = (fh.frame_type == INTER_FRAME) ? br.f(1) : 1;
    if (!update) {
        int idx = br.f(3);
        g = refs[idx].fg;
        g.apply_grain = 1;
        g.grain_seed = seed;
        g.update_grain = 0;
        return;
    }
    memset(&g, 0, sizeof(g));
    g.apply_grain = 1;
    g.grain_seed = seed;
    g.update_grain = 1;
    g.num_y_points = br.f(4);
    for (int i = 0; i < g.num_y_points; i++) {
        g.point_y_value[i] = br.f(8);
        g.point_y_scaling[i] = br.f(8);
    }
    g.chroma_scaling_from_luma = seq.mono_chrome ? 0 : br.f(1);
    if (seq.mono_chrome || g.chroma_scaling_from_luma ||
        (seq.subsampling_x == 1 && seq.subsampling_y == 1 && g.num_y_points == 0)) {
        g.num_cb_points = g.num_cr_points = 0;
    } else {
        g.num_cb_points = br.f(4);
        for (int i = 0; i < g.num_cb_points; i++) {
            g.point_cb_value[i] = br.f(8);
            g.point_cb_scaling[i] = br.f(8);
        }
        g.num_cr_points = br.f(4);
        for (int i = 0; i < g.num_cr_points; i++) {
            g.point_cr_value[i] = br.f(8);
            g.point_cr_scaling[i] = br.f(8);
        }
    }
    g.grain_scaling = br.f(2) + 8;
    g.ar_coeff_lag = br.f(2);
    int numPosLuma = 2 * g.ar_coeff_lag * (g.ar_coeff_lag + 1);
    int numPosChroma = numPosLuma;
    if (g.num_y_points) {
        numPosChroma = numPosLuma + 1;
        for (int i = 0; i < numPosLuma; i++) g.ar_coeffs_y[i] = (int)br.f(8) - 128;
    }
    if (g.chroma_scaling_from_luma || g.num_cb_points)
        for (int i = 0; i < numPosChroma; i++) g.ar_coeffs_cb[i] = (int)br.f(8) - 128;
    if (g.chroma_scaling_from_luma || g.num_cr_points)
        for (int i = 0; i < numPosChroma; i++) g.ar_coeffs_cr[i] = (int)br.f(8) - 128;
    g.ar_coeff_shift = br.f(2) + 6;
    g.grain_scale_shift = br.f(2);
    if (g.num_cb_points) {
        g.cb_mult = br.f(8);
        g.cb_luma_mult = br.f(8);
        g.cb_offset = br.f(9);
    }
    if (g.num_cr_points) {
        g.cr_mult = br.f(8);
        g.cr_luma_mult = br.f(8);
        g.cr_offset = br.f(9);
    }
    g.overlap_flag = br.f(1);
    g.clip_to_restricted_range = br.f(1);
}

bool HeaderParser::parse_frame_header(BitReader& br, FrameHdr& fh, int temporal_id, int spatial_id) {
    if (!seq.valid) return fail("frame header before sequence header");
    memset(&fh, 0, sizeof(fh));
    fh.temporal_id = temporal_id;
    fh.spatial_id = spatial_id;
    const int allFrames = (1 << NUM_REF_FRAMES) - 1;
    int idLen = 0;
    if (seq.frame_id_numbers_present) idLen = seq.additional_frame_id_length_minus_1 + seq.delta_frame_id_length_minus_2 + 3;
    if (seq.reduced_still_picture_header) {
        fh.frame_type = KEY_FRAME;
        fh.frame_is_intra = 1;
        fh.show_frame = 1;
    } else {
        fh.show_existing_frame = br.f(1);
        if (fh.show_existing_frame) {
            fh.frame_to_show_map_idx = br.f(3);
            if (seq.decoder_model_info_present && !seq.equal_picture_interval)
                br.f(seq.frame_presentation_time_length_minus_1 + 1);
            fh.refresh_frame_flags = 0;
            if (seq.frame_id_numbers_present) br.f(idLen);
            const RefHdrState& r = refs[fh.frame_to_show_map_idx];
            if (!r.valid) return fail("show_existing_frame of an empty slot");
            fh.frame_type = r.frame_type;
            if (fh.frame_type == KEY_FRAME) fh.refresh_frame_flags = allFrames;
            if (seq.film_grain_params_present) fh.fg = r.fg;
            fh.frame_width = r.frame_width;
            fh.frame_height = r.frame_height;
            fh.upscaled_width = r.upscaled_width;
            fh.render_width = r.render_width;
            fh.render_height = r.render_height;
            fh.mi_cols = r.mi_cols;
            fh.mi_rows = r.mi_rows;
            fh.order_hint = r.order_hint;
            fh.show_frame = 1;
            return !br.err || fail("truncated frame header");
        }
        fh.frame_type = br.f(2);
        fh.frame_is_intra = (fh.frame_type == INTRA_ONLY_FRAME || fh.frame_type == KEY_FRAME);
        fh.show_frame = br.f(1);
        if (fh.show_frame && seq.decoder_model_info_present && !seq.equal_picture_interval)
            br.f(seq.frame_presentation_time_length_minus_1 + 1);
        if (fh.show_frame) fh.showable_frame = fh.frame_type != KEY_FRAME;
        else fh.showable_frame = br.f(1);
        if (fh.frame_type == SWITCH_FRAME || (fh.frame_type == KEY_FRAME && fh.show_frame)) fh.error_resilient_mode = 1;
        else fh.error_resilient_mode = br.f(1);
    }
    if (fh.frame_type == KEY_FRAME && fh.show_frame) {
        for (int i = 0; i < NUM_REF_FRAMES; i++) {
            refs[i].valid = 0;
            refs[i].order_hint = 0;
        }
        for (int i = 0; i < 8; i++) fh.order_hints[i] = 0;
    }
    fh.disable_cdf_update = br.f(1);
    if (seq.seq_force_screen_content_tools == 2) fh.allow_screen_content_tools = br.f(1);
    else fh.allow_screen_content_tools = seq.seq_force_screen_content_tools;
    if (fh.allow_screen_content_tools) {
        if (seq.seq_force_integer_mv == 2) fh.force_integer_mv = br.f(1);
        else fh.force_integer_mv = seq.seq_force_integer_mv;
    } else {
        fh.force_integer_mv = 0;
    }
    if (fh.frame_is_intra) fh.force_integer_mv = 1;
    if (seq.frame_id_numbers_present) {
        fh.current_frame_id = br.f(idLen);
        // mark_ref_frames: invalidate refs too far in id space
        int diffLen = seq.delta_frame_id_length_minus_2 + 2;
        for (int i = 0; i < NUM_REF_FRAMES; i++) {
            if (fh.frame_type == KEY_FRAME && fh.show_frame) {
                refs[i].valid = 0;
            } else if (fh.current_frame_id > (1 << diffLen)) {
                if (refs[i].frame_id > fh.current_frame_id || refs[i].frame_id < (fh.current_frame_id - (1 << diffLen)))
                    refs[i].valid = 0;
            } else {
                if (refs[i].frame_id > fh.current_frame_id &&
                    refs[i].frame_id < ((1 << idLen) + fh.current_frame_id - (1 << diffLen)))
                    refs[i].valid = 0;
            }
        }
    }
    if (fh.frame_type == SWITCH_FRAME) fh.frame_size_override_flag = 1;
    else if (seq.reduced_still_picture_header) fh.frame_size_override_flag = 0;
    else fh.frame_size_override_flag = br.f(1);
    fh.order_hint = seq.order_hint_bits ? br.f(seq.order_hint_bits) : 0;
    if (fh.frame_is_intra || fh.error_resilient_mode) fh.primary_ref_frame = PRIMARY_REF_NONE;
    else fh.primary_ref_frame = br.f(3);
    if (seq.decoder_model_info_present) {
        int present = br.f(1);
        if (present) {
            for (int op = 0; op <= seq.operating_points_cnt_minus_1; op++) {
                if (seq.decoder_model_present_for_this_op[op]) {
                    int idc = seq.operating_point_idc[op];
                    int inT = (idc >> temporal_id) & 1;
                    int inS = (idc >> (spatial_id + 8)) & 1;
                    if (idc == 0 || (inT && inS)) br.f(seq.buffer_removal_time_length_minus_1 + 1);
                }
            }
        }
    }
    fh.allow_high_precision_mv = 0;
    fh.use_ref_frame_mvs = 0;
    fh.allow_intrabc = 0;
    if (fh.frame_type == SWITCH_FRAME || (fh.frame_type == KEY_FRAME && fh.show_frame)) fh.refresh_frame_flags = allFrames;
    else fh.refresh_frame_flags = br.f(8);
    if (!fh.frame_is_intra || fh.refresh_frame_flags != allFrames) {
        if (fh.error_resilient_mode && seq.enable_order_hint) {
            for (int i = 0; i < NUM_REF_FRAMES; i++) {
                fh.ref_order_hint[i] = br.f(seq.order_hint_bits);
                if (fh.ref_order_hint[i] != refs[i].order_hint || !refs[i].valid) {
                    // missing reference: spec 7.20 sets up a grey frame; we only track the hint
                    refs[i].order_hint = fh.ref_order_hint[i];
                }
            }
        }
    }
    if (fh.frame_is_intra) {
        frame_size(br, fh);
        render_size(br, fh);
        if (fh.allow_screen_content_tools && fh.upscaled_width == fh.frame_width) fh.allow_intrabc = br.f(1);
    } else {
        if (!seq.enable_order_hint) {
            fh.frame_refs_short_signaling = 0;
        } else {
            fh.frame_refs_short_signaling = br.f(1);
            if (fh.frame_refs_short_signaling) {
                int last_idx = br.f(3);
                int gold_idx = br.f(3);
                set_frame_refs(fh, last_idx, gold_idx);
            }
        }
        for (int i = 0; i < REFS_PER_FRAME; i++) {
            if (!fh.frame_refs_short_signaling) fh.ref_frame_idx[i] = br.f(3);
            if (seq.frame_id_numbers_present) br.f(seq.delta_frame_id_length_minus_2 + 2);
        }
        if (fh.frame_size_override_flag && !fh.error_resilient_mode) {
            frame_size_with_refs(br, fh);
        } else {
            frame_size(br, fh);
            render_size(br, fh);
        }
        if (fh.force_integer_mv) fh.allow_high_precision_mv = 0;
        else fh.allow_high_precision_mv = br.f(1);
        fh.is_filter_switchable = br.f(1);
        fh.interpolation_filter = fh.is_filter_switchable ? INTERP_SWITCHABLE : (int)br.f(2);
        fh.is_motion_mode_switchable = br.f(1);
        if (fh.error_resilient_mode || !seq.enable_ref_frame_mvs) fh.use_ref_frame_mvs = 0;
        else fh.use_ref_frame_mvs = br.f(1);
        for (int i = 0; i < REFS_PER_FRAME; i++) {
            int refFrame = LAST_FRAME + i;
            int hint = refs[fh.ref_frame_idx[i]].order_hint;
            fh.order_hints[refFrame] = hint;
            if (!seq.enable_order_hint) fh.ref_frame_sign_bias[refFrame] = 0;
            else fh.ref_frame_sign_bias[refFrame] = get_relative_dist(hint, fh.order_hint) > 0;
        }
    }
    if (seq.reduced_still_picture_header || fh.disable_cdf_update) fh.disable_frame_end_update_cdf = 1;
    else fh.disable_frame_end_update_cdf = br.f(1);
    if (fh.primary_ref_frame == PRIMARY_REF_NONE) setup_past_independence(fh);
    else load_previous(fh);
    if (!tile_info(br, fh)) return false;
    quantization_params(br, fh);
    segmentation_params(br, fh);
    // delta_q / delta_lf params
    fh.delta_q_res = 0;
    fh.delta_q_present = 0;
    if (fh.base_q_idx > 0) fh.delta_q_present = br.f(1);
    if (fh.delta_q_present) fh.delta_q_res = br.f(2);
    fh.delta_lf_present = fh.delta_lf_res = fh.delta_lf_multi = 0;
    if (fh.delta_q_present) {
        if (!fh.allow_intrabc) fh.delta_lf_present = br.f(1);
        if (fh.delta_lf_present) {
            fh.delta_lf_res = br.f(2);
            fh.delta_lf_multi = br.f(1);
        }
    }
    fh.coded_lossless = 1;
    for (int s = 0; s < MAX_SEGMENTS; s++) {
        int qidx = get_qidx(fh, 1, s, 0);
        fh.qidx_seg[s] = qidx;
        fh.lossless_array[s] = qidx == 0 && fh.delta_q_y_dc == 0 && fh.delta_q_u_ac == 0 && fh.delta_q_u_dc == 0 &&
                               fh.delta_q_v_ac == 0 && fh.delta_q_v_dc == 0;
        if (!fh.lossless_array[s]) fh.coded_lossless = 0;
        if (fh.using_qmatrix) {
            if (fh.lossless_array[s]) {
                fh.seg_qm_level[0][s] = fh.seg_qm_level[1][s] = fh.seg_qm_level[2][s] = 15;
            } else {
                fh.seg_qm_level[0][s] = fh.qm_y;
                fh.seg_qm_level[1][s] = fh.qm_u;
                fh.seg_qm_level[2][s] = fh.qm_v;
            }
        } else {
            fh.seg_qm_level[0][s] = fh.seg_qm_level[1][s] = fh.seg_qm_level[2][s] = 15;
        }
    }
    fh.all_lossless = fh.coded_lossless && (fh.frame_width == fh.upscaled_width);
    loop_filter_params(br, fh);
    cdef_params(br, fh);
    lr_params(br, fh);
    if (fh.coded_lossless) fh.tx_mode = ONLY_4X4;
    else fh.tx_mode = br.f(1) ? TX_MODE_SELECT : TX_MODE_LARGEST;
    fh.reference_select = fh.frame_is_intra ? 0 : br.f(1);
    skip_mode_params(br, fh);
    if (fh.frame_is_intra || fh.error_resilient_mode || !seq.enable_warped_motion) fh.allow_warped_motion = 0;
    else fh.allow_warped_motion = br.f(1);
    fh.reduced_tx_set = br.f(1);
    global_motion_params(br, fh);
    film_grain_params(br, fh);
    if (br.err) return fail("truncated frame header");
    return true;
}

bool HeaderParser::parse_tile_group_header(BitReader& br, const FrameHdr& fh, TileGroupInfo& tg) {
    int numTiles = fh.tile_cols * fh.tile_rows;
    int present = 0;
    if (numTiles > 1) present = br.f(1);
    if (numTiles == 1 || !present) {
        tg.tg_start = 0;
        tg.tg_end = numTiles - 1;
    } else {
        int bits = fh.tile_cols_log2 + fh.tile_rows_log2;
        tg.tg_start = br.f(bits);
        tg.tg_end = br.f(bits);
    }
    br.byte_align();
    tg.data_offset = br.byte_pos();
    if (br.err) return fail("truncated tile group header");
    if (tg.tg_end < tg.tg_start || tg.tg_end >= numTiles) return fail("bad tile group range");
    return true;
}

void HeaderParser::reference_update(const FrameHdr& fh) {
    for (int i = 0; i < NUM_REF_FRAMES; i++) {
        if (!((fh.refresh_frame_flags >> i) & 1)) continue;
        RefHdrState& r = refs[i];
        r.valid = 1;
        r.frame_id = fh.current_frame_id;
        r.upscaled_width = fh.upscaled_width;
        r.frame_width = fh.frame_width;
        r.frame_height = fh.frame_height;
        r.render_width = fh.render_width;
        r.render_height = fh.render_height;
        r.mi_cols = fh.mi_cols;
        r.mi_rows = fh.mi_rows;
        r.frame_type = fh.frame_type;
        r.order_hint = fh.order_hint;
        r.bit_depth = seq.bit_depth;
        r.subsampling_x = seq.subsampling_x;
        r.subsampling_y = seq.subsampling_y;
        r.showable_frame = fh.showable_frame;
        for (int j = 0; j < 8; j++) r.saved_order_hints[j] = fh.order_hints[j];
        memcpy(r.saved_gm_params, fh.gm_params, sizeof(fh.gm_params));
        memcpy(r.lf_ref_deltas, fh.lf.ref_deltas, sizeof(r.lf_ref_deltas));
        memcpy(r.lf_mode_deltas, fh.lf.mode_deltas, sizeof(r.lf_mode_deltas));
        r.seg = fh.seg;
        r.fg = fh.fg;
    }
}

}  // namespace av1r

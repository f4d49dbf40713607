// Per-tile symbol parse (host, sequential): partition tree, mode info, transform sizes/types,
// coefficients -> device work-lists.  Spec 5.11 (syntax) + 8.3 (CDF selection).
#pragma once
#include <string>

#include "frame_state.h"
#include "msac.h"
#include "obu.h"

namespace av1r {

struct RefFrameInfo;   // inter prediction state of reference slots (defined in inter.h)

class TileDecoder {
public:
    TileDecoder(const SeqHdr& seq, const HeaderParser& hp, FrameWork& fw, TileOut& to, const CdfCtx& init_cdf);
    // returns 0, AV1R_EBITSTREAM or AV1R_ENOSYS (err holds the reason)
    int decode_tile(const uint8_t* data, size_t sz, int tile_row, int tile_col);
    CdfCtx cdf;
    uint16_t cdf_tail_pad_[16] = {0};   // the 16-lane symbol decoder (msac.h) loads / stores 32 bytes from the start of a CDF
    std::string err;

private:
    const SeqHdr& seq;
    const HeaderParser& hp;
    FrameWork& fw;
    TileOut& to;
    const FrameHdr& fh;
    Msac ms;
    int fail_code = 0;
    // tile geometry
    int mi_row_start = 0, mi_row_end = 0, mi_col_start = 0, mi_col_end = 0;
    int current_q_index = 0;
    int read_deltas = 0;
    int delta_lf[4] = {0, 0, 0, 0};
    // coefficient contexts (absolute 4x4 indices per plane)
    std::vector<uint8_t> above_level[3], above_dc[3], left_level[3], left_dc[3];
    std::vector<uint8_t> above_seg_pred, left_seg_pred;
    // loop restoration references
    int ref_sgr_xqd[3][2];
    int ref_lr_wiener[3][2][3];
    // per-superblock "block decoded" flags: [plane][y+1][x+1], y,x in -1..32
    uint8_t block_decoded[3][35][35];
    // current block
    BlockInfo* b = nullptr;
    int mi_row = 0, mi_col = 0, bw4 = 0, bh4 = 0;
    int avail_u = 0, avail_l = 0, avail_u_chroma = 0, avail_l_chroma = 0;
    int max_luma_w = 0, max_luma_h = 0;
    SbRange cur_sb;
    int cur_unit = -1;
    uint8_t levels[(32 + 4) * (32 + 5)];   // zero-padded level map of the transform block being parsed: min(level, 15)
    uint8_t levels3[(32 + 4) * (32 + 5)];  // same, min(level, 3)
    int ftype_cache[2] = {-1, -1};

    bool fail(int code, const char* msg) { if (!fail_code) { fail_code = code; err = msg; } return false; }
    bool is_inside(int r, int c) const { return c >= mi_col_start && c < mi_col_end && r >= mi_row_start && r < mi_row_end; }
    BlockInfo* blk(int r, int c) const { return fw.mi[(size_t)r * fw.mi_cols + c]; }

    void clear_block_decoded_flags(int r, int c, int sb4);
    void read_lr(int r, int c, int bsize);
    void read_lr_unit(int plane, int unit_row, int unit_col);
    int decode_subexp_bool(int num_syms, int k);
    int decode_signed_subexp_with_ref_bool(int low, int high, int k, int r);
    bool decode_partition(int r, int c, int bsize);
    bool decode_block(int r, int c, int bsize);
    void intra_frame_mode_info();
    void intra_segment_id();
    void read_segment_id();
    void read_skip();
    void read_cdef();
    void read_delta_qindex();
    void read_delta_lf();
    void intra_angle_info_y();
    void intra_angle_info_uv();
    void read_cfl_alphas();
    void filter_intra_mode_info();
    void read_block_tx_size();
    void read_tx_size(int allow_select);
    void reset_block_context();
    void residual();
    void transform_block(int plane, int base_x, int base_y, int txsz, int x, int y);
    int coeffs(int plane, int start_x, int start_y, int txsz, TxRec& rec);
    int get_tx_set(int txsz) const;
    void read_transform_type(int x4, int y4, int txsz);
    int compute_tx_type(int plane, int txsz, int block_x, int block_y) const;
    int filter_type(int plane) const;
    void push_record(const TxRec& rec, int ux, int uy);
    void palette_mode_info();
    void palette_tokens();
    int get_palette_cache(int plane, uint16_t* cache) const;
    uint32_t pal_entry[3] = {0, 0, 0};
    void intra_mode_tail();   // uv mode, palette, filter-intra: shared by intra frames and intra blocks of inter frames
    // inter (tile_inter.cpp)
    void inter_frame_mode_info();
    void read_var_tx_size(int row, int col, int txsz, int depth);
    void transform_tree(int start_x, int start_y, int w, int h);
    void inter_segment_id(int pre_skip);
    void read_skip_mode();
    void read_is_inter();
    void intra_block_mode_info();
    void inter_block_mode_info();
    void read_ref_frames();
    int count_refs(int frame_type) const;
    int seg_feature_active(int f) const { return fh.seg.enabled && fh.seg.feature_enabled[b->segment_id][f]; }
    void assign_mv(int is_compound);
    void assign_dv();
    void mark_skipped_inter_block();
    void read_mv(int list);
    int read_mv_component(int comp);
    void read_interintra_mode(int is_compound);
    void read_motion_mode(int is_compound);
    void read_compound_type(int is_compound);
    int has_overlappable_candidates() const;
    void find_warp_samples();
    void add_warp_sample(int delta_row, int delta_col);
    void warp_estimation();
    // motion vector prediction (spec 7.10.2)
    void find_mv_stack(int is_compound);
    void setup_global_mv(int list);
    void lower_mv_precision(Mv& mv) const;
    void scan_row(int delta_row, int is_compound);
    void scan_col(int delta_col, int is_compound);
    void scan_point(int delta_row, int delta_col, int is_compound);
    void add_ref_mv_candidate(int r, int c, int is_compound, int weight);
    void temporal_scan(int is_compound);
    void add_tpl_ref_mv(int delta_row, int delta_col, int is_compound);
    void sort_stack(int start, int end);
    void extra_search(int is_compound);
    void add_extra_mv_candidate(int r, int c, int is_compound);
    void context_and_clamping(int is_compound, int num_new);
    void emit_inter_block();
    void emit_interintra_records();

    // neighbour context of the current block (spec 5.11.7: AboveRefFrame, LeftIntra ...)
    int above_ref[2] = {0, -1}, left_ref[2] = {0, -1};
    int above_intra = 1, left_intra = 1, above_single = 1, left_single = 1;
    // mv stack
    int num_mv_found = 0, new_mv_count = 0, found_match = 0, close_matches = 0, total_matches = 0;
    int zero_mv_ctx = 0, new_mv_ctx = 0, ref_mv_ctx = 0, ref_mv_idx = 0;
    Mv ref_stack[12][2];
    int weight_stack[12];
    Mv global_mvs[2];
    int ref_id_count[2], ref_diff_count[2];
    Mv ref_id_mvs[2][2], ref_diff_mvs[2][2];
    // warp samples
    int num_samples = 0, num_samples_scanned = 0;
    int cand_list[8][4];
    std::vector<uint8_t> above_seg_pred_ctx_unused;
};

// spec 7.11.3.6: shear parameters of a warp model; returns warpValid
int setup_shear(const int32_t* mat, int16_t out[4]);

}  // namespace av1r

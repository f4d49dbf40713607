// Inter-frame mode info (spec 5.11.7 - 5.11.27), motion vector prediction (7.10.2), warp sample
// search and local warp estimation (7.10.4, 7.11.3.8), variable transform trees (5.11.17) and the
// emission of the K2 work-list (InterBlk / ObmcNb / WarpRec).  Host, sequential, not the optimised path.
#include <algorithm>
#include <cstdlib>

#include "../../include/av1r.h"
#include "tile.h"

namespace av1r {

#include "tables/tables_inter.inc"

namespace {
enum { MV_JOINT_ZERO = 0, MV_JOINT_HNZVZ = 1, MV_JOINT_HZVNZ = 2, MV_JOINT_HNZVNZ = 3 };
enum { REF_CAT_LEVEL = 640, MAX_REF_MV_STACK_SIZE = 8, MV_BORDER = 128, MAX_FRAME_DISTANCE = 31 };
enum { WARPEDMODEL_PREC_BITS = 16, WARP_PARAM_REDUCE_BITS = 6, DIV_LUT_BITS = 8, DIV_LUT_PREC_BITS = 14, LS_MV_MAX = 256 };
const int kDivMult[32] = {0,    16384, 8192, 5461, 4096, 3276, 2730, 2340, 2048, 1820, 1638, 1489, 1365, 1260, 1170, 1092,
                          1024, 963,   910,  862,  819,  780,  744,  712,  682,  655,  630,  606,  585,  564,  546,  528};
const uint8_t kWedgeBits[BLOCK_SIZES_ALL] = {0, 0, 0, 4, 4, 4, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 0, 0};
const uint8_t kCompoundModeCtxMap[3][5] = {{0, 1, 1, 1, 1}, {1, 2, 3, 4, 4}, {4, 4, 5, 6, 7}};

inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
inline int64_t round2s64(int64_t x, int n) {
    if (n == 0) return x;
    return x >= 0 ? (x + ((int64_t)1 << (n - 1))) >> n : -((-x + ((int64_t)1 << (n - 1))) >> n);
}
inline int floor_log2(uint32_t x) { return 31 - __builtin_clz(x); }
inline bool mv_eq(const Mv& a, const Mv& b) { return a.row == b.row && a.col == b.col; }
inline bool has_newmv(int m) { return m == NEWMV || m == NEW_NEWMV || m == NEAR_NEWMV || m == NEW_NEARMV || m == NEAREST_NEWMV || m == NEW_NEARESTMV; }
inline bool has_nearmv(int m) { return m == NEARMV || m == NEAR_NEARMV || m == NEAR_NEWMV || m == NEW_NEARMV; }
inline int find_tx_size(int w, int h) {
    for (int t = 0; t < TX_SIZES_ALL; t++)
        if (kTxW[t] == w && kTxH[t] == h) return t;
    return TX_4X4;
}

void resolve_divisor(int64_t d, int& shift, int& factor) {
    const uint64_t a = (uint64_t)(d < 0 ? -d : d);
    const int n = 63 - __builtin_clzll(a);
    const uint64_t e = a - ((uint64_t)1 << n);
    int f;
    if (n > DIV_LUT_BITS) f = (int)((e + ((uint64_t)1 << (n - DIV_LUT_BITS - 1))) >> (n - DIV_LUT_BITS));
    else f = (int)(e << (DIV_LUT_BITS - n));
    shift = n + DIV_LUT_PREC_BITS;
    factor = d < 0 ? -(int)av1t_div_lut[f] : (int)av1t_div_lut[f];
}
}  // namespace

int setup_shear(const int32_t* mat, int16_t out[4]) {
    if (mat[2] <= 0) return 0;
    const int alpha0 = clip3(-32768, 32767, mat[2] - (1 << WARPEDMODEL_PREC_BITS));
    const int beta0 = clip3(-32768, 32767, mat[3]);
    int shift, factor;
    resolve_divisor(mat[2], shift, factor);
    const int64_t v = (int64_t)mat[4] * (1 << WARPEDMODEL_PREC_BITS);
    const int gamma0 = clip3(-32768, 32767, (int)round2s64(v * factor, shift));
    const int64_t w = (int64_t)mat[3] * mat[4];
    const int delta0 = clip3(-32768, 32767, mat[5] - (int)round2s64(w * factor, shift) - (1 << WARPEDMODEL_PREC_BITS));
    auto red = [](int x) { return (int)(round2s64(x, WARP_PARAM_REDUCE_BITS) * (1 << WARP_PARAM_REDUCE_BITS)); };
    const int alpha = red(alpha0), beta = red(beta0), gamma = red(gamma0), delta = red(delta0);
    out[0] = (int16_t)alpha;
    out[1] = (int16_t)beta;
    out[2] = (int16_t)gamma;
    out[3] = (int16_t)delta;
    if (4 * std::abs(alpha) + 7 * std::abs(beta) >= (1 << WARPEDMODEL_PREC_BITS)) return 0;
    if (4 * std::abs(gamma) + 4 * std::abs(delta) >= (1 << WARPEDMODEL_PREC_BITS)) return 0;
    return 1;
}

// ---------------------------------------------------------------- mode info
void TileDecoder::inter_frame_mode_info() {
    b->use_intrabc = 0;
    above_ref[0] = left_ref[0] = INTRA_FRAME;
    above_ref[1] = left_ref[1] = -1;
    if (avail_u) {
        const BlockInfo* a = blk(mi_row - 1, mi_col);
        above_ref[0] = a->ref_frame[0];
        above_ref[1] = a->ref_frame[1];
    }
    if (avail_l) {
        const BlockInfo* l = blk(mi_row, mi_col - 1);
        left_ref[0] = l->ref_frame[0];
        left_ref[1] = l->ref_frame[1];
    }
    above_intra = above_ref[0] <= INTRA_FRAME;
    left_intra = left_ref[0] <= INTRA_FRAME;
    above_single = above_ref[1] <= INTRA_FRAME;
    left_single = left_ref[1] <= INTRA_FRAME;
    b->skip = 0;
    inter_segment_id(1);
    read_skip_mode();
    if (b->skip_mode) b->skip = 1;
    else read_skip();
    if (!fh.seg.seg_id_pre_skip) inter_segment_id(0);
    b->lossless = (uint8_t)fh.lossless_array[b->segment_id];
    read_cdef();
    read_delta_qindex();
    read_delta_lf();
    read_deltas = 0;
    read_is_inter();
    if (b->is_inter) inter_block_mode_info();
    else intra_block_mode_info();
}

void TileDecoder::inter_segment_id(int pre_skip) {
    if (!fh.seg.enabled) {
        b->segment_id = 0;
        return;
    }
    int predicted = 0;
    if (!fw.prev_seg_ids.empty()) {
        const int xmis = std::min(fw.mi_cols - mi_col, bw4), ymis = std::min(fw.mi_rows - mi_row, bh4);
        predicted = 7;
        for (int y = 0; y < ymis; y++)
            for (int x = 0; x < xmis; x++) predicted = std::min(predicted, (int)fw.prev_seg_ids[(size_t)(mi_row + y) * fw.mi_cols + mi_col + x]);
    }
    if (!fh.seg.update_map) {
        b->segment_id = (uint8_t)predicted;
        return;
    }
    if (pre_skip && !fh.seg.seg_id_pre_skip) {
        b->segment_id = 0;
        return;
    }
    auto set_ctx = [&](int v) {
        for (int i = 0; i < bw4; i++) above_seg_pred[mi_col + i] = (uint8_t)v;
        for (int i = 0; i < bh4; i++) left_seg_pred[mi_row + i] = (uint8_t)v;
    };
    if (!pre_skip && b->skip) {
        set_ctx(0);
        read_segment_id();
        return;
    }
    if (fh.seg.temporal_update) {
        const int ctx = left_seg_pred[mi_row] + above_seg_pred[mi_col];
        const int sp = ms.symbol(cdf.seg_pred[ctx], 2);
        if (sp) b->segment_id = (uint8_t)predicted;
        else read_segment_id();
        set_ctx(sp);
    } else {
        read_segment_id();
    }
}

void TileDecoder::read_skip_mode() {
    b->skip_mode = 0;
    if (seg_feature_active(SEG_LVL_SKIP) || seg_feature_active(SEG_LVL_REF_FRAME) || seg_feature_active(SEG_LVL_GLOBALMV) || !fh.skip_mode_present ||
        kBlockW[b->bsize] < 8 || kBlockH[b->bsize] < 8)
        return;
    int ctx = 0;
    if (avail_u) ctx += blk(mi_row - 1, mi_col)->skip_mode;
    if (avail_l) ctx += blk(mi_row, mi_col - 1)->skip_mode;
    b->skip_mode = (uint8_t)ms.symbol(cdf.skip_mode[ctx], 2);
}

void TileDecoder::read_is_inter() {
    if (b->skip_mode) {
        b->is_inter = 1;
    } else if (seg_feature_active(SEG_LVL_REF_FRAME)) {
        b->is_inter = fh.seg.feature_data[b->segment_id][SEG_LVL_REF_FRAME] != INTRA_FRAME;
    } else if (seg_feature_active(SEG_LVL_GLOBALMV)) {
        b->is_inter = 1;
    } else {
        int ctx;
        if (avail_u && avail_l) ctx = (left_intra && above_intra) ? 3 : (left_intra || above_intra);
        else if (avail_u || avail_l) ctx = 2 * (avail_u ? above_intra : left_intra);
        else ctx = 0;
        b->is_inter = (uint8_t)ms.symbol(cdf.intra_inter[ctx], 2);
    }
}

void TileDecoder::intra_block_mode_info() {
    to.tool_hist[TOOL_INTRA_IN_INTER]++;
    b->ref_frame[0] = INTRA_FRAME;
    b->ref_frame[1] = -1;
    b->y_mode = (uint8_t)ms.symbol(cdf.y_mode[kSizeGroup[b->bsize]], 13);
    intra_angle_info_y();
    intra_mode_tail();
}

int TileDecoder::count_refs(int frame_type) const {
    int c = 0;
    if (avail_u) c += (above_ref[0] == frame_type) + (above_ref[1] == frame_type);
    if (avail_l) c += (left_ref[0] == frame_type) + (left_ref[1] == frame_type);
    return c;
}

static inline int ref_count_ctx(int c0, int c1) { return c0 < c1 ? 0 : (c0 == c1 ? 1 : 2); }
static inline int check_backward(int r) { return r >= BWDREF_FRAME && r <= ALTREF_FRAME; }
static inline int is_samedir_ref_pair(int r0, int r1) {
    if (r0 <= INTRA_FRAME || r1 <= INTRA_FRAME) return 0;
    return (r0 >= BWDREF_FRAME) == (r1 >= BWDREF_FRAME);
}

void TileDecoder::read_ref_frames() {
    if (b->skip_mode) {
        b->ref_frame[0] = (int8_t)fh.skip_mode_frame[0];
        b->ref_frame[1] = (int8_t)fh.skip_mode_frame[1];
        return;
    }
    if (seg_feature_active(SEG_LVL_REF_FRAME)) {
        b->ref_frame[0] = (int8_t)fh.seg.feature_data[b->segment_id][SEG_LVL_REF_FRAME];
        b->ref_frame[1] = -1;
        return;
    }
    if (seg_feature_active(SEG_LVL_SKIP) || seg_feature_active(SEG_LVL_GLOBALMV)) {
        b->ref_frame[0] = LAST_FRAME;
        b->ref_frame[1] = -1;
        return;
    }
    const int cl = count_refs(LAST_FRAME), cl2 = count_refs(LAST2_FRAME), cl3 = count_refs(LAST3_FRAME), cg = count_refs(GOLDEN_FRAME);
    const int cb = count_refs(BWDREF_FRAME), ca2 = count_refs(ALTREF2_FRAME), ca = count_refs(ALTREF_FRAME);
    const int ctx_p1 = ref_count_ctx(cl + cl2 + cl3 + cg, cb + ca2 + ca);
    const int ctx_p2 = ref_count_ctx(cb + ca2, ca);
    const int ctx_p3 = ref_count_ctx(cl + cl2, cl3 + cg);
    const int ctx_p4 = ref_count_ctx(cl, cl2);
    const int ctx_p5 = ref_count_ctx(cl3, cg);
    const int ctx_p6 = ref_count_ctx(cb, ca2);
    int comp_mode = 0;
    if (fh.reference_select && std::min(bw4, bh4) >= 2) {
        int ctx;
        if (avail_u && avail_l) {
            if (above_single && left_single) ctx = check_backward(above_ref[0]) ^ check_backward(left_ref[0]);
            else if (above_single) ctx = 2 + (check_backward(above_ref[0]) || above_intra);
            else if (left_single) ctx = 2 + (check_backward(left_ref[0]) || left_intra);
            else ctx = 4;
        } else if (avail_u) {
            ctx = above_single ? check_backward(above_ref[0]) : 3;
        } else if (avail_l) {
            ctx = left_single ? check_backward(left_ref[0]) : 3;
        } else {
            ctx = 1;
        }
        comp_mode = ms.symbol(cdf.comp_inter[ctx], 2);
    }
    if (comp_mode) {
        int ctx;
        {
            const int above0 = above_ref[0], above1 = above_ref[1], left0 = left_ref[0], left1 = left_ref[1];
            const int above_comp_inter = avail_u && !above_intra && !above_single;
            const int left_comp_inter = avail_l && !left_intra && !left_single;
            const int above_uni = above_comp_inter && is_samedir_ref_pair(above0, above1);
            const int left_uni = left_comp_inter && is_samedir_ref_pair(left0, left1);
            if (avail_u && !above_intra && avail_l && !left_intra) {
                const int samedir = is_samedir_ref_pair(above0, left0);
                if (!above_comp_inter && !left_comp_inter) ctx = 1 + 2 * samedir;
                else if (!above_comp_inter) ctx = !left_uni ? 1 : 3 + samedir;
                else if (!left_comp_inter) ctx = !above_uni ? 1 : 3 + samedir;
                else if (!above_uni && !left_uni) ctx = 0;
                else if (!above_uni || !left_uni) ctx = 2;
                else ctx = 3 + ((above0 == BWDREF_FRAME) == (left0 == BWDREF_FRAME));
            } else if (avail_u && avail_l) {
                if (above_comp_inter) ctx = 1 + 2 * above_uni;
                else if (left_comp_inter) ctx = 1 + 2 * left_uni;
                else ctx = 2;
            } else if (above_comp_inter) {
                ctx = 4 * above_uni;
            } else if (left_comp_inter) {
                ctx = 4 * left_uni;
            } else {
                ctx = 2;
            }
        }
        const int comp_ref_type = ms.symbol(cdf.comp_ref_type[ctx], 2);
        if (comp_ref_type == 0) {   // UNIDIR_COMP_REFERENCE
            if (ms.symbol(cdf.uni_comp_ref[ctx_p1][0], 2)) {
                b->ref_frame[0] = BWDREF_FRAME;
                b->ref_frame[1] = ALTREF_FRAME;
            } else {
                const int ctx1 = ref_count_ctx(cl2, cl3 + cg);
                if (ms.symbol(cdf.uni_comp_ref[ctx1][1], 2)) {
                    b->ref_frame[0] = LAST_FRAME;
                    b->ref_frame[1] = ms.symbol(cdf.uni_comp_ref[ctx_p5][2], 2) ? GOLDEN_FRAME : LAST3_FRAME;
                } else {
                    b->ref_frame[0] = LAST_FRAME;
                    b->ref_frame[1] = LAST2_FRAME;
                }
            }
        } else {
            if (ms.symbol(cdf.comp_ref[ctx_p3][0], 2) == 0) b->ref_frame[0] = ms.symbol(cdf.comp_ref[ctx_p4][1], 2) ? LAST2_FRAME : LAST_FRAME;
            else b->ref_frame[0] = ms.symbol(cdf.comp_ref[ctx_p5][2], 2) ? GOLDEN_FRAME : LAST3_FRAME;
            if (ms.symbol(cdf.comp_bwdref[ctx_p2][0], 2) == 0) b->ref_frame[1] = ms.symbol(cdf.comp_bwdref[ctx_p6][1], 2) ? ALTREF2_FRAME : BWDREF_FRAME;
            else b->ref_frame[1] = ALTREF_FRAME;
        }
    } else {
        if (ms.symbol(cdf.single_ref[ctx_p1][0], 2)) {
            if (ms.symbol(cdf.single_ref[ctx_p2][1], 2) == 0) b->ref_frame[0] = ms.symbol(cdf.single_ref[ctx_p6][5], 2) ? ALTREF2_FRAME : BWDREF_FRAME;
            else b->ref_frame[0] = ALTREF_FRAME;
        } else {
            if (ms.symbol(cdf.single_ref[ctx_p3][2], 2)) b->ref_frame[0] = ms.symbol(cdf.single_ref[ctx_p5][4], 2) ? GOLDEN_FRAME : LAST3_FRAME;
            else b->ref_frame[0] = ms.symbol(cdf.single_ref[ctx_p4][3], 2) ? LAST2_FRAME : LAST_FRAME;
        }
        b->ref_frame[1] = -1;
    }
}

void TileDecoder::inter_block_mode_info() {
    b->pal_size[0] = b->pal_size[1] = 0;
    b->uv_mode = DC_PRED;
    read_ref_frames();
    const int is_compound = b->ref_frame[1] > INTRA_FRAME;
    find_mv_stack(is_compound);
    if (b->skip_mode) {
        b->y_mode = NEAREST_NEARESTMV;
    } else if (seg_feature_active(SEG_LVL_SKIP) || seg_feature_active(SEG_LVL_GLOBALMV)) {
        b->y_mode = GLOBALMV;
    } else if (is_compound) {
        const int ctx = kCompoundModeCtxMap[ref_mv_ctx >> 1][std::min(new_mv_ctx, 4)];
        b->y_mode = (uint8_t)(NEAREST_NEARESTMV + ms.symbol(cdf.inter_compound_mode[ctx], 8));
    } else {
        if (ms.symbol(cdf.newmv[new_mv_ctx], 2) == 0) b->y_mode = NEWMV;
        else if (ms.symbol(cdf.zeromv[zero_mv_ctx], 2) == 0) b->y_mode = GLOBALMV;
        else b->y_mode = ms.symbol(cdf.refmv[ref_mv_ctx], 2) ? NEARMV : NEARESTMV;
    }
    auto drl_ctx = [&](int idx) {
        const int w0 = weight_stack[idx], w1 = weight_stack[idx + 1];
        if (w0 >= REF_CAT_LEVEL) return w1 < REF_CAT_LEVEL ? 1 : 0;
        return 2;   // (a lower-weight entry never precedes a higher one after sorting)
    };
    ref_mv_idx = 0;
    if (b->y_mode == NEWMV || b->y_mode == NEW_NEWMV) {
        for (int idx = 0; idx < 2; idx++)
            if (num_mv_found > idx + 1) {
                if (ms.symbol(cdf.drl[drl_ctx(idx)], 2) == 0) {
                    ref_mv_idx = idx;
                    break;
                }
                ref_mv_idx = idx + 1;
            }
    } else if (has_nearmv(b->y_mode)) {
        ref_mv_idx = 1;
        for (int idx = 1; idx < 3; idx++)
            if (num_mv_found > idx + 1) {
                if (ms.symbol(cdf.drl[drl_ctx(idx)], 2) == 0) {
                    ref_mv_idx = idx;
                    break;
                }
                ref_mv_idx = idx + 1;
            }
    }
    assign_mv(is_compound);
    read_interintra_mode(is_compound);
    read_motion_mode(is_compound);
    read_compound_type(is_compound);
    if (fh.interpolation_filter == INTERP_SWITCHABLE) {
        const int ndir = seq.enable_dual_filter ? 2 : 1;
        for (int dir = 0; dir < ndir; dir++) {
            int needs;
            const int large = std::min(kBlockW[b->bsize], kBlockH[b->bsize]) >= 8;
            if (b->skip_mode || b->motion_mode == WARPED_CAUSAL) needs = 0;
            else if (large && b->y_mode == GLOBALMV) needs = fh.gm_type[b->ref_frame[0]] == GM_TRANSLATION;
            else if (large && b->y_mode == GLOBAL_GLOBALMV) needs = fh.gm_type[b->ref_frame[1]] == GM_TRANSLATION;
            else needs = 1;
            if (needs) {
                int ctx = ((dir & 1) * 2 + (b->ref_frame[1] > INTRA_FRAME)) * 4;
                int left_type = 3, above_type = 3;
                if (avail_l) {
                    const BlockInfo* l = blk(mi_row, mi_col - 1);
                    if (l->ref_frame[0] == b->ref_frame[0] || l->ref_frame[1] == b->ref_frame[0]) left_type = l->interp_filter[dir];
                }
                if (avail_u) {
                    const BlockInfo* a = blk(mi_row - 1, mi_col);
                    if (a->ref_frame[0] == b->ref_frame[0] || a->ref_frame[1] == b->ref_frame[0]) above_type = a->interp_filter[dir];
                }
                if (left_type == above_type) ctx += left_type;
                else if (left_type == 3) ctx += above_type;
                else if (above_type == 3) ctx += left_type;
                else ctx += 3;
                b->interp_filter[dir] = (uint8_t)ms.symbol(cdf.switchable_interp[ctx], 3);
            } else {
                b->interp_filter[dir] = INTERP_EIGHTTAP;
            }
        }
        if (!seq.enable_dual_filter) b->interp_filter[1] = b->interp_filter[0];
    } else {
        b->interp_filter[0] = b->interp_filter[1] = (uint8_t)fh.interpolation_filter;
    }
}

static int get_mode(int y_mode, int list) {
    if (list == 0) {
        if (y_mode < NEAREST_NEARESTMV) return y_mode;
        if (y_mode == NEW_NEWMV || y_mode == NEW_NEARESTMV || y_mode == NEW_NEARMV) return NEWMV;
        if (y_mode == NEAREST_NEARESTMV || y_mode == NEAREST_NEWMV) return NEARESTMV;
        if (y_mode == NEAR_NEARMV || y_mode == NEAR_NEWMV) return NEARMV;
        return GLOBALMV;
    }
    if (y_mode == NEW_NEWMV || y_mode == NEAREST_NEWMV || y_mode == NEAR_NEWMV) return NEWMV;
    if (y_mode == NEAREST_NEARESTMV || y_mode == NEW_NEARESTMV) return NEARESTMV;
    if (y_mode == NEAR_NEARMV || y_mode == NEW_NEARMV) return NEARMV;
    return GLOBALMV;
}

void TileDecoder::assign_mv(int is_compound) {
    for (int i = 0; i < 1 + is_compound; i++) {
        const int comp_mode = get_mode(b->y_mode, i);
        if (comp_mode == GLOBALMV) {
            b->mv[i] = global_mvs[i];
        } else {
            int pos = comp_mode == NEARESTMV ? 0 : ref_mv_idx;
            if (comp_mode == NEWMV && num_mv_found <= 1) pos = 0;
            b->mv[i] = ref_stack[pos][i];
        }
        if (comp_mode == NEWMV) read_mv(i);
    }
    if (!is_compound) b->mv[1] = Mv{0, 0};
}

// Block vector of an intra-block-copy block (spec 5.11.26 assign_mv with use_intrabc, MvCtx = MV_CONTEXTS_INTRABC): predicted from
// the first non-zero entry of the stack, else from a default one superblock up (or a superblock + 256 samples to the left on the
// tile's first superblock row); the difference is coded like a NEWMV at integer precision (force_integer_mv is 1 in intra frames).
void TileDecoder::assign_dv() {
    Mv pred = ref_stack[0][0];
    if (pred.row == 0 && pred.col == 0) pred = ref_stack[1][0];
    if (pred.row == 0 && pred.col == 0) {
        const int sb4 = seq.use_128x128_superblock ? 32 : 16;
        if (mi_row - sb4 < mi_row_start) {
            pred.row = 0;
            pred.col = (int16_t)(-(sb4 * 4 + 256) * 8);
        } else {
            pred.row = (int16_t)(-(sb4 * 4 * 8));
            pred.col = 0;
        }
    }
    int diff[2] = {0, 0};
    const int joint = ms.symbol(cdf.dv_joints, 4);
    auto comp = [&](int c) {
        const int sign = ms.symbol(c ? cdf.dv_c1_sign : cdf.dv_c0_sign, 2);
        const int mv_class = ms.symbol(c ? cdf.dv_c1_classes : cdf.dv_c0_classes, 11);
        int mag;
        if (mv_class == 0) {
            const int class0_bit = ms.symbol(c ? cdf.dv_c1_class0 : cdf.dv_c0_class0, 2);
            mag = ((class0_bit << 3) | (3 << 1) | 1) + 1;
        } else {
            int d = 0;
            for (int i = 0; i < mv_class; i++) d |= ms.symbol(c ? cdf.dv_c1_bits[i] : cdf.dv_c0_bits[i], 2) << i;
            mag = (2 << (mv_class + 2)) + ((d << 3) | (3 << 1) | 1) + 1;
        }
        return sign ? -mag : mag;
    };
    if (joint == MV_JOINT_HZVNZ || joint == MV_JOINT_HNZVNZ) diff[0] = comp(0);
    if (joint == MV_JOINT_HNZVZ || joint == MV_JOINT_HNZVNZ) diff[1] = comp(1);
    b->mv[0].row = (int16_t)(pred.row + diff[0]);
    b->mv[0].col = (int16_t)(pred.col + diff[1]);
    b->mv[1] = Mv{0, 0};
    // conformance (spec 6.10.25 / 7.11.3.2): whole-sample vector, source rectangle inside the tile.  That the source has been
    // reconstructed already is checked where the wavefront kernel's plan is built (k3_plan.h): a vector into the future must be
    // an error, not a wait that never ends.
    const int sx0 = mi_col * 4 + (b->mv[0].col >> 3), sy0 = mi_row * 4 + (b->mv[0].row >> 3);
    if ((b->mv[0].row & 7) || (b->mv[0].col & 7) || sx0 < mi_col_start * 4 || sy0 < mi_row_start * 4 ||
        sx0 + bw4 * 4 > mi_col_end * 4 || sy0 + bh4 * 4 > mi_row_end * 4)
        fail(AV1R_EBITSTREAM, "intra block copy vector points outside the tile");
}

int TileDecoder::read_mv_component(int comp) {
    uint16_t* sign_c = comp ? cdf.mv_c1_sign : cdf.mv_c0_sign;
    uint16_t* classes_c = comp ? cdf.mv_c1_classes : cdf.mv_c0_classes;
    const int sign = ms.symbol(sign_c, 2);
    const int mv_class = ms.symbol(classes_c, 11);
    int mag;
    if (mv_class == 0) {
        const int class0_bit = ms.symbol(comp ? cdf.mv_c1_class0 : cdf.mv_c0_class0, 2);
        int fr = 3, hp = 1;
        if (!fh.force_integer_mv) fr = ms.symbol(comp ? cdf.mv_c1_class0_fp[class0_bit] : cdf.mv_c0_class0_fp[class0_bit], 4);
        if (fh.allow_high_precision_mv) hp = ms.symbol(comp ? cdf.mv_c1_class0_hp : cdf.mv_c0_class0_hp, 2);
        mag = ((class0_bit << 3) | (fr << 1) | hp) + 1;
    } else {
        int d = 0;
        for (int i = 0; i < mv_class; i++) d |= ms.symbol(comp ? cdf.mv_c1_bits[i] : cdf.mv_c0_bits[i], 2) << i;
        mag = 2 << (mv_class + 2);
        int fr = 3, hp = 1;
        if (!fh.force_integer_mv) fr = ms.symbol(comp ? cdf.mv_c1_fp : cdf.mv_c0_fp, 4);
        if (fh.allow_high_precision_mv) hp = ms.symbol(comp ? cdf.mv_c1_hp : cdf.mv_c0_hp, 2);
        mag += ((d << 3) | (fr << 1) | hp) + 1;
    }
    return sign ? -mag : mag;
}

void TileDecoder::read_mv(int list) {
    int diff[2] = {0, 0};
    const int joint = ms.symbol(cdf.mv_joints, 4);
    if (joint == MV_JOINT_HZVNZ || joint == MV_JOINT_HNZVNZ) diff[0] = read_mv_component(0);
    if (joint == MV_JOINT_HNZVZ || joint == MV_JOINT_HNZVNZ) diff[1] = read_mv_component(1);
    b->mv[list].row = (int16_t)(b->mv[list].row + diff[0]);
    b->mv[list].col = (int16_t)(b->mv[list].col + diff[1]);
}

void TileDecoder::read_interintra_mode(int is_compound) {
    b->interintra = 0;
    if (!b->skip_mode && seq.enable_interintra_compound && !is_compound && b->bsize >= BLOCK_8X8 && b->bsize <= BLOCK_32X32) {
        const int g = kSizeGroup[b->bsize];   // cdf arrays keep libaom's layout: index = size group, entry 0 unused
        b->interintra = (uint8_t)ms.symbol(cdf.interintra[g], 2);
        if (b->interintra) {
            b->interintra_mode = (uint8_t)ms.symbol(cdf.interintra_mode[g], 4);
            b->ref_frame[1] = INTRA_FRAME;
            b->angle_y = b->angle_uv = 0;
            b->use_filter_intra = 0;
            b->wedge_interintra = (uint8_t)ms.symbol(cdf.wedge_interintra[b->bsize], 2);
            if (b->wedge_interintra) {
                b->wedge_index = (uint8_t)ms.symbol(cdf.wedge_idx[b->bsize], 16);
                b->wedge_sign = 0;
            }
        }
    }
}

int TileDecoder::has_overlappable_candidates() const {
    if (avail_u)
        for (int x4 = mi_col; x4 < std::min(fw.mi_cols, mi_col + bw4); x4 += 2) {
            const int x5 = std::min(x4 | 1, fw.mi_cols - 1);
            if (blk(mi_row - 1, x5)->ref_frame[0] > INTRA_FRAME) return 1;
        }
    if (avail_l)
        for (int y4 = mi_row; y4 < std::min(fw.mi_rows, mi_row + bh4); y4 += 2) {
            const int y5 = std::min(y4 | 1, fw.mi_rows - 1);
            if (blk(y5, mi_col - 1)->ref_frame[0] > INTRA_FRAME) return 1;
        }
    return 0;
}

void TileDecoder::add_warp_sample(int delta_row, int delta_col) {
    if (num_samples_scanned >= 8) return;
    const int mv_row = mi_row + delta_row, mv_col = mi_col + delta_col;
    if (!is_inside(mv_row, mv_col)) return;
    const BlockInfo* n = blk(mv_row, mv_col);
    if (!n) return;   // not yet decoded
    if (n->ref_frame[0] != b->ref_frame[0] || n->ref_frame[1] != -1) return;
    const int cw4 = kBlockW4[n->bsize], ch4 = kBlockH4[n->bsize];
    const int cand_row = mv_row & ~(ch4 - 1), cand_col = mv_col & ~(cw4 - 1);
    const int mid_y = cand_row * 4 + ch4 * 2 - 1, mid_x = cand_col * 4 + cw4 * 2 - 1;
    const int threshold = clip3(16, 112, std::max((int)kBlockW[b->bsize], (int)kBlockH[b->bsize]));
    const int diff = std::abs(n->mv[0].row - b->mv[0].row) + std::abs(n->mv[0].col - b->mv[0].col);
    const int valid = diff <= threshold;
    num_samples_scanned++;
    if (!valid && num_samples_scanned > 1) return;
    cand_list[num_samples][0] = mid_y * 8;
    cand_list[num_samples][1] = mid_x * 8;
    cand_list[num_samples][2] = mid_y * 8 + n->mv[0].row;
    cand_list[num_samples][3] = mid_x * 8 + n->mv[0].col;
    if (valid) num_samples++;
}

void TileDecoder::find_warp_samples() {
    num_samples = num_samples_scanned = 0;
    int do_top_left = 1, do_top_right = 1;
    if (avail_u) {
        const int src_w = kBlockW4[blk(mi_row - 1, mi_col)->bsize];
        if (bw4 <= src_w) {
            const int src_col = -(mi_col & (src_w - 1));
            if (src_col < 0) do_top_left = 0;
            if (src_col + src_w > bw4) do_top_right = 0;
            add_warp_sample(-1, 0);
        } else {
            int step;
            for (int i = 0; i < std::min(bw4, fw.mi_cols - mi_col); i += step) {
                step = std::min(bw4, (int)kBlockW4[blk(mi_row - 1, mi_col + i)->bsize]);
                add_warp_sample(-1, i);
            }
        }
    }
    if (avail_l) {
        const int src_h = kBlockH4[blk(mi_row, mi_col - 1)->bsize];
        if (bh4 <= src_h) {
            const int src_row = -(mi_row & (src_h - 1));
            if (src_row < 0) do_top_left = 0;
            add_warp_sample(0, -1);
        } else {
            int step;
            for (int i = 0; i < std::min(bh4, fw.mi_rows - mi_row); i += step) {
                step = std::min(bh4, (int)kBlockH4[blk(mi_row + i, mi_col - 1)->bsize]);
                add_warp_sample(i, -1);
            }
        }
    }
    if (do_top_left) add_warp_sample(-1, -1);
    if (do_top_right && std::max(bw4, bh4) <= 16) add_warp_sample(-1, bw4);
    if (num_samples == 0 && num_samples_scanned > 0) num_samples = 1;
}

void TileDecoder::read_motion_mode(int is_compound) {
    b->motion_mode = SIMPLE_TRANSLATION;
    if (b->skip_mode) return;
    if (!fh.is_motion_mode_switchable) return;
    if (std::min(kBlockW[b->bsize], kBlockH[b->bsize]) < 8) return;
    if (!fh.force_integer_mv && (b->y_mode == GLOBALMV || b->y_mode == GLOBAL_GLOBALMV) && fh.gm_type[b->ref_frame[0]] > GM_TRANSLATION) return;
    if (is_compound || b->ref_frame[1] == INTRA_FRAME || !has_overlappable_candidates()) return;
    find_warp_samples();
    if (fh.force_integer_mv || num_samples == 0 || !fh.allow_warped_motion || fw.ref_scaled[b->ref_frame[0]]) {
        b->motion_mode = ms.symbol(cdf.obmc[b->bsize], 2) ? OBMC_CAUSAL : SIMPLE_TRANSLATION;
    } else {
        b->motion_mode = (uint8_t)ms.symbol(cdf.motion_mode[b->bsize], 3);
    }
}

void TileDecoder::read_compound_type(int is_compound) {
    b->comp_group_idx = 0;
    b->compound_idx = 1;
    if (b->skip_mode) {
        b->compound_type = COMPOUND_AVERAGE;
        return;
    }
    if (is_compound) {
        const int n = kWedgeBits[b->bsize];
        if (seq.enable_masked_compound) {
            int ctx = 0;
            if (avail_u) {
                if (!above_single) ctx += blk(mi_row - 1, mi_col)->comp_group_idx;
                else if (above_ref[0] == ALTREF_FRAME) ctx += 3;
            }
            if (avail_l) {
                if (!left_single) ctx += blk(mi_row, mi_col - 1)->comp_group_idx;
                else if (left_ref[0] == ALTREF_FRAME) ctx += 3;
            }
            b->comp_group_idx = (uint8_t)ms.symbol(cdf.comp_group_idx[std::min(5, ctx)], 2);
        }
        if (b->comp_group_idx == 0) {
            if (seq.enable_jnt_comp) {
                const int fwd = std::abs(hp.get_relative_dist(fh.order_hints[b->ref_frame[0]], fh.order_hint));
                const int bck = std::abs(hp.get_relative_dist(fh.order_hints[b->ref_frame[1]], fh.order_hint));
                int ctx = fwd == bck ? 3 : 0;
                if (avail_u) {
                    if (!above_single) ctx += blk(mi_row - 1, mi_col)->compound_idx;
                    else if (above_ref[0] == ALTREF_FRAME) ctx++;
                }
                if (avail_l) {
                    if (!left_single) ctx += blk(mi_row, mi_col - 1)->compound_idx;
                    else if (left_ref[0] == ALTREF_FRAME) ctx++;
                }
                b->compound_idx = (uint8_t)ms.symbol(cdf.compound_index[ctx], 2);
                b->compound_type = b->compound_idx ? COMPOUND_AVERAGE : COMPOUND_DISTANCE;
            } else {
                b->compound_type = COMPOUND_AVERAGE;
            }
        } else {
            if (n == 0) b->compound_type = COMPOUND_DIFFWTD;
            else b->compound_type = (uint8_t)ms.symbol(cdf.compound_type[b->bsize], 2);
        }
        if (b->compound_type == COMPOUND_WEDGE) {
            b->wedge_index = (uint8_t)ms.symbol(cdf.wedge_idx[b->bsize], 16);
            b->wedge_sign = (uint8_t)ms.literal(1);
        } else if (b->compound_type == COMPOUND_DIFFWTD) {
            b->mask_type = (uint8_t)ms.literal(1);
        }
    } else {
        if (b->interintra) b->compound_type = b->wedge_interintra ? COMPOUND_WEDGE : COMPOUND_INTRA;
        else b->compound_type = COMPOUND_AVERAGE;
    }
}

// ---------------------------------------------------------------- motion vector prediction (7.10.2)
void TileDecoder::lower_mv_precision(Mv& mv) const {
    if (fh.allow_high_precision_mv) return;
    int16_t* c[2] = {&mv.row, &mv.col};
    for (int i = 0; i < 2; i++) {
        int v = *c[i];
        if (fh.force_integer_mv) {
            const int a = std::abs(v), a_int = (a + 3) >> 3;
            if (a > 0) v = v > 0 ? (a_int << 3) : -(a_int << 3);
        } else if (v & 1) {
            v += v > 0 ? -1 : 1;
        }
        *c[i] = (int16_t)v;
    }
}

void TileDecoder::setup_global_mv(int list) {
    const int ref = b->ref_frame[list];
    Mv mv{0, 0};
    const int typ = ref > INTRA_FRAME ? fh.gm_type[ref] : GM_IDENTITY;
    if (ref <= INTRA_FRAME || typ == GM_IDENTITY) {
        mv = Mv{0, 0};
    } else if (typ == GM_TRANSLATION) {
        mv.row = (int16_t)(fh.gm_params[ref][0] >> (WARPEDMODEL_PREC_BITS - 3));
        mv.col = (int16_t)(fh.gm_params[ref][1] >> (WARPEDMODEL_PREC_BITS - 3));
    } else {
        const int32_t* gm = fh.gm_params[ref];
        const int x = mi_col * 4 + kBlockW[b->bsize] / 2 - 1, y = mi_row * 4 + kBlockH[b->bsize] / 2 - 1;
        const int64_t xc = (int64_t)(gm[2] - (1 << WARPEDMODEL_PREC_BITS)) * x + (int64_t)gm[3] * y + gm[0];
        const int64_t yc = (int64_t)gm[4] * x + (int64_t)(gm[5] - (1 << WARPEDMODEL_PREC_BITS)) * y + gm[1];
        if (fh.allow_high_precision_mv) {
            mv.row = (int16_t)round2s64(yc, WARPEDMODEL_PREC_BITS - 3);
            mv.col = (int16_t)round2s64(xc, WARPEDMODEL_PREC_BITS - 3);
        } else {
            mv.row = (int16_t)(round2s64(yc, WARPEDMODEL_PREC_BITS - 2) * 2);
            mv.col = (int16_t)(round2s64(xc, WARPEDMODEL_PREC_BITS - 2) * 2);
        }
    }
    lower_mv_precision(mv);
    global_mvs[list] = mv;
}

void TileDecoder::add_ref_mv_candidate(int r, int c, int is_compound, int weight) {
    const BlockInfo* n = blk(r, c);
    if (!n || !n->is_inter) return;
    const int large = std::min(kBlockW[n->bsize], kBlockH[n->bsize]) >= 8;
    if (!is_compound) {
        for (int cl = 0; cl < 2; cl++) {
            if (n->ref_frame[cl] != b->ref_frame[0]) continue;
            Mv cand;
            if ((n->y_mode == GLOBALMV || n->y_mode == GLOBAL_GLOBALMV) && fh.gm_type[b->ref_frame[0]] > GM_TRANSLATION && large) cand = global_mvs[0];
            else cand = n->mv[cl];
            lower_mv_precision(cand);
            found_match = 1;
            if (has_newmv(n->y_mode)) new_mv_count++;
            int idx;
            for (idx = 0; idx < num_mv_found; idx++)
                if (mv_eq(cand, ref_stack[idx][0])) break;
            if (idx < num_mv_found) {
                weight_stack[idx] += weight;
            } else if (num_mv_found < MAX_REF_MV_STACK_SIZE) {
                ref_stack[num_mv_found][0] = cand;
                weight_stack[num_mv_found] = weight;
                num_mv_found++;
            }
        }
    } else if (n->ref_frame[0] == b->ref_frame[0] && n->ref_frame[1] == b->ref_frame[1]) {
        Mv cand[2] = {n->mv[0], n->mv[1]};
        for (int i = 0; i < 2; i++) {
            if (n->y_mode == GLOBAL_GLOBALMV && fh.gm_type[b->ref_frame[i]] > GM_TRANSLATION && large) cand[i] = global_mvs[i];
            lower_mv_precision(cand[i]);
        }
        found_match = 1;
        if (has_newmv(n->y_mode)) new_mv_count++;
        int idx;
        for (idx = 0; idx < num_mv_found; idx++)
            if (mv_eq(cand[0], ref_stack[idx][0]) && mv_eq(cand[1], ref_stack[idx][1])) break;
        if (idx < num_mv_found) {
            weight_stack[idx] += weight;
        } else if (num_mv_found < MAX_REF_MV_STACK_SIZE) {
            ref_stack[num_mv_found][0] = cand[0];
            ref_stack[num_mv_found][1] = cand[1];
            weight_stack[num_mv_found] = weight;
            num_mv_found++;
        }
    }
}

void TileDecoder::scan_row(int delta_row, int is_compound) {
    int delta_col = 0;
    const int end4 = std::min(std::min(bw4, fw.mi_cols - mi_col), 16);
    if (std::abs(delta_row) > 1) {
        delta_row += mi_row & 1;
        delta_col = 1 - (mi_col & 1);
    }
    const int use_step16 = bw4 >= 16;
    for (int i = 0; i < end4;) {
        const int mv_row = mi_row + delta_row, mv_col = mi_col + delta_col + i;
        if (!is_inside(mv_row, mv_col) || !blk(mv_row, mv_col)) break;
        int len = std::min(bw4, (int)kBlockW4[blk(mv_row, mv_col)->bsize]);
        if (std::abs(delta_row) > 1) len = std::max(2, len);
        if (use_step16) len = std::max(4, len);
        add_ref_mv_candidate(mv_row, mv_col, is_compound, len * 2);
        i += len;
    }
}

void TileDecoder::scan_col(int delta_col, int is_compound) {
    int delta_row = 0;
    const int end4 = std::min(std::min(bh4, fw.mi_rows - mi_row), 16);
    if (std::abs(delta_col) > 1) {
        delta_row = 1 - (mi_row & 1);
        delta_col += mi_col & 1;
    }
    const int use_step16 = bh4 >= 16;
    for (int i = 0; i < end4;) {
        const int mv_row = mi_row + delta_row + i, mv_col = mi_col + delta_col;
        if (!is_inside(mv_row, mv_col) || !blk(mv_row, mv_col)) break;
        int len = std::min(bh4, (int)kBlockH4[blk(mv_row, mv_col)->bsize]);
        if (std::abs(delta_col) > 1) len = std::max(2, len);
        if (use_step16) len = std::max(4, len);
        add_ref_mv_candidate(mv_row, mv_col, is_compound, len * 2);
        i += len;
    }
}

void TileDecoder::scan_point(int delta_row, int delta_col, int is_compound) {
    const int mv_row = mi_row + delta_row, mv_col = mi_col + delta_col;
    if (!is_inside(mv_row, mv_col) || mv_row >= fw.mi_rows || mv_col >= fw.mi_cols) return;
    if (!blk(mv_row, mv_col)) return;   // not yet decoded in this frame
    add_ref_mv_candidate(mv_row, mv_col, is_compound, 4);
}

static Mv mv_projection(Mv mv, int numerator, int denominator) {
    const int cd = std::min((int)MAX_FRAME_DISTANCE, denominator);
    const int cn = clip3(-MAX_FRAME_DISTANCE, MAX_FRAME_DISTANCE, numerator);
    const int factor = kDivMult[cd];
    Mv o;
    o.row = (int16_t)clip3(-(1 << 14) + 1, (1 << 14) - 1, (int)round2s64((int64_t)mv.row * cn * factor, 14));
    o.col = (int16_t)clip3(-(1 << 14) + 1, (1 << 14) - 1, (int)round2s64((int64_t)mv.col * cn * factor, 14));
    return o;
}

void TileDecoder::add_tpl_ref_mv(int delta_row, int delta_col, int is_compound) {
    const int mv_row = (mi_row + delta_row) | 1, mv_col = (mi_col + delta_col) | 1;
    if (!is_inside(mv_row, mv_col)) return;
    const int x8 = mv_col >> 1, y8 = mv_row >> 1;
    if (delta_row == 0 && delta_col == 0) zero_mv_ctx = 1;
    const MfMv& m = fw.mfmv[(size_t)y8 * (fw.mi_cols >> 1) + x8];
    if (!m.ref_offset) return;
    to.tool_hist[TOOL_TEMPORAL_MV]++;
    Mv cand[2];
    for (int i = 0; i < 1 + is_compound; i++) {
        const int cur_offset = hp.get_relative_dist(fh.order_hint, fh.order_hints[b->ref_frame[i]]);
        cand[i] = mv_projection(m.mv, cur_offset, m.ref_offset);
        lower_mv_precision(cand[i]);
    }
    if (!is_compound) {
        if (delta_row == 0 && delta_col == 0)
            zero_mv_ctx = (std::abs(cand[0].row - global_mvs[0].row) >= 16 || std::abs(cand[0].col - global_mvs[0].col) >= 16) ? 1 : 0;
        int idx;
        for (idx = 0; idx < num_mv_found; idx++)
            if (mv_eq(cand[0], ref_stack[idx][0])) break;
        if (idx < num_mv_found) {
            weight_stack[idx] += 2;
        } else if (num_mv_found < MAX_REF_MV_STACK_SIZE) {
            ref_stack[num_mv_found][0] = cand[0];
            weight_stack[num_mv_found] = 2;
            num_mv_found++;
        }
    } else {
        if (delta_row == 0 && delta_col == 0)
            zero_mv_ctx = (std::abs(cand[0].row - global_mvs[0].row) >= 16 || std::abs(cand[0].col - global_mvs[0].col) >= 16 ||
                           std::abs(cand[1].row - global_mvs[1].row) >= 16 || std::abs(cand[1].col - global_mvs[1].col) >= 16)
                              ? 1
                              : 0;
        int idx;
        for (idx = 0; idx < num_mv_found; idx++)
            if (mv_eq(cand[0], ref_stack[idx][0]) && mv_eq(cand[1], ref_stack[idx][1])) break;
        if (idx < num_mv_found) {
            weight_stack[idx] += 2;
        } else if (num_mv_found < MAX_REF_MV_STACK_SIZE) {
            ref_stack[num_mv_found][0] = cand[0];
            ref_stack[num_mv_found][1] = cand[1];
            weight_stack[num_mv_found] = 2;
            num_mv_found++;
        }
    }
}

void TileDecoder::temporal_scan(int is_compound) {
    const int step_w4 = bw4 >= 16 ? 4 : 2, step_h4 = bh4 >= 16 ? 4 : 2;
    for (int dr = 0; dr < std::min(bh4, 16); dr += step_h4)
        for (int dc = 0; dc < std::min(bw4, 16); dc += step_w4) add_tpl_ref_mv(dr, dc, is_compound);
    const int allow_ext = bh4 >= 2 && bh4 < 16 && bw4 >= 2 && bw4 < 16;
    if (allow_ext) {
        const int pos[3][2] = {{bh4, -2}, {bh4, bw4}, {bh4 - 2, bw4}};
        for (int i = 0; i < 3; i++) {
            const int dr = pos[i][0], dc = pos[i][1];
            const int row = (mi_row & 15) + dr, col = (mi_col & 15) + dc;
            if (row >= 0 && row < 16 && col >= 0 && col < 16) add_tpl_ref_mv(dr, dc, is_compound);
        }
    }
}

void TileDecoder::sort_stack(int start, int end) {
    while (end > start) {
        int new_end = start;
        for (int idx = start + 1; idx < end; idx++)
            if (weight_stack[idx - 1] < weight_stack[idx]) {
                std::swap(weight_stack[idx - 1], weight_stack[idx]);
                std::swap(ref_stack[idx - 1][0], ref_stack[idx][0]);
                std::swap(ref_stack[idx - 1][1], ref_stack[idx][1]);
                new_end = idx;
            }
        end = new_end;
    }
}

void TileDecoder::add_extra_mv_candidate(int r, int c, int is_compound) {
    const BlockInfo* n = blk(r, c);
    if (is_compound) {
        for (int cl = 0; cl < 2; cl++) {
            const int cand_ref = n->ref_frame[cl];
            if (cand_ref <= INTRA_FRAME) continue;
            for (int list = 0; list < 2; list++) {
                Mv cand = n->mv[cl];
                if (cand_ref == b->ref_frame[list] && ref_id_count[list] < 2) {
                    ref_id_mvs[list][ref_id_count[list]++] = cand;
                } else if (ref_diff_count[list] < 2) {
                    if (fh.ref_frame_sign_bias[cand_ref] != fh.ref_frame_sign_bias[b->ref_frame[list]]) {
                        cand.row = (int16_t)-cand.row;
                        cand.col = (int16_t)-cand.col;
                    }
                    ref_diff_mvs[list][ref_diff_count[list]++] = cand;
                }
            }
        }
    } else {
        for (int cl = 0; cl < 2; cl++) {
            const int cand_ref = n->ref_frame[cl];
            if (cand_ref <= INTRA_FRAME) continue;
            Mv cand = n->mv[cl];
            if (fh.ref_frame_sign_bias[cand_ref] != fh.ref_frame_sign_bias[b->ref_frame[0]]) {
                cand.row = (int16_t)-cand.row;
                cand.col = (int16_t)-cand.col;
            }
            int idx;
            for (idx = 0; idx < num_mv_found; idx++)
                if (mv_eq(cand, ref_stack[idx][0])) break;
            if (idx == num_mv_found) {
                ref_stack[idx][0] = cand;
                weight_stack[idx] = 2;
                num_mv_found++;
            }
        }
    }
}

void TileDecoder::extra_search(int is_compound) {
    for (int l = 0; l < 2; l++) ref_id_count[l] = ref_diff_count[l] = 0;
    int w4 = std::min(16, bw4), h4 = std::min(16, bh4);
    w4 = std::min(w4, fw.mi_cols - mi_col);
    h4 = std::min(h4, fw.mi_rows - mi_row);
    const int num4 = std::min(w4, h4);
    for (int pass = 0; pass < 2; pass++) {
        int idx = 0;
        while (idx < num4 && num_mv_found < 2) {
            const int mv_row = pass == 0 ? mi_row - 1 : mi_row + idx;
            const int mv_col = pass == 0 ? mi_col + idx : mi_col - 1;
            if (!is_inside(mv_row, mv_col) || !blk(mv_row, mv_col)) break;
            add_extra_mv_candidate(mv_row, mv_col, is_compound);
            idx += pass == 0 ? kBlockW4[blk(mv_row, mv_col)->bsize] : kBlockH4[blk(mv_row, mv_col)->bsize];
        }
    }
    if (is_compound) {
        Mv combined[2][2];
        for (int list = 0; list < 2; list++) {
            int cc = 0;
            for (int idx = 0; idx < ref_id_count[list]; idx++) combined[cc++][list] = ref_id_mvs[list][idx];
            for (int idx = 0; idx < ref_diff_count[list] && cc < 2; idx++) combined[cc++][list] = ref_diff_mvs[list][idx];
            while (cc < 2) combined[cc++][list] = global_mvs[list];
        }
        if (num_mv_found == 1) {
            const int same = mv_eq(combined[0][0], ref_stack[0][0]) && mv_eq(combined[0][1], ref_stack[0][1]);
            ref_stack[1][0] = combined[same ? 1 : 0][0];
            ref_stack[1][1] = combined[same ? 1 : 0][1];
            weight_stack[1] = 2;
            num_mv_found = 2;
        } else {
            for (int idx = 0; idx < 2; idx++) {
                ref_stack[idx][0] = combined[idx][0];
                ref_stack[idx][1] = combined[idx][1];
                weight_stack[idx] = 2;
            }
            num_mv_found = 2;
        }
    } else {
        for (int idx = num_mv_found; idx < 2; idx++) ref_stack[idx][0] = global_mvs[0];
    }
}

void TileDecoder::context_and_clamping(int is_compound, int num_new) {
    const int bw = kBlockW[b->bsize], bh = kBlockH[b->bsize];
    const int nl = is_compound ? 2 : 1;
    const int to_top = -(mi_row * 4 * 8), to_bottom = (fw.mi_rows - bh4 - mi_row) * 4 * 8;
    const int to_left = -(mi_col * 4 * 8), to_right = (fw.mi_cols - bw4 - mi_col) * 4 * 8;
    for (int idx = 0; idx < num_mv_found; idx++)
        for (int l = 0; l < nl; l++) {
            Mv& m = ref_stack[idx][l];
            m.row = (int16_t)clip3(to_top - (MV_BORDER + bh * 8), to_bottom + (MV_BORDER + bh * 8), m.row);
            m.col = (int16_t)clip3(to_left - (MV_BORDER + bw * 8), to_right + (MV_BORDER + bw * 8), m.col);
        }
    if (close_matches == 0) {
        new_mv_ctx = std::min(total_matches, 1);
        ref_mv_ctx = total_matches;
    } else if (close_matches == 1) {
        new_mv_ctx = 3 - std::min(num_new, 1);
        ref_mv_ctx = 2 + total_matches;
    } else {
        new_mv_ctx = 5 - std::min(num_new, 1);
        ref_mv_ctx = 5;
    }
}

void TileDecoder::find_mv_stack(int is_compound) {
    num_mv_found = 0;
    new_mv_count = 0;
    for (int i = 0; i < 12; i++) {
        weight_stack[i] = 0;
        ref_stack[i][0] = ref_stack[i][1] = Mv{0, 0};
    }
    global_mvs[1] = Mv{0, 0};
    setup_global_mv(0);
    if (is_compound) setup_global_mv(1);
    found_match = 0;
    scan_row(-1, is_compound);
    int found_above = found_match;
    found_match = 0;
    scan_col(-1, is_compound);
    int found_left = found_match;
    found_match = 0;
    if (std::max(bw4, bh4) <= 16) scan_point(-1, bw4, is_compound);
    if (found_match) found_above = 1;
    close_matches = found_above + found_left;
    const int num_nearest = num_mv_found, num_new = new_mv_count;
    for (int idx = 0; idx < num_nearest; idx++) weight_stack[idx] += REF_CAT_LEVEL;
    zero_mv_ctx = 0;
    if (fh.use_ref_frame_mvs && !fw.mfmv.empty()) temporal_scan(is_compound);
    found_match = 0;
    scan_point(-1, -1, is_compound);
    if (found_match) found_above = 1;
    found_match = 0;
    scan_row(-3, is_compound);
    if (found_match) found_above = 1;
    found_match = 0;
    scan_col(-3, is_compound);
    if (found_match) found_left = 1;
    found_match = 0;
    if (bh4 > 1) scan_row(-5, is_compound);
    if (found_match) found_above = 1;
    found_match = 0;
    if (bw4 > 1) scan_col(-5, is_compound);
    if (found_match) found_left = 1;
    total_matches = found_above + found_left;
    sort_stack(0, num_nearest);
    sort_stack(num_nearest, num_mv_found);
    if (num_mv_found < 2) extra_search(is_compound);
    context_and_clamping(is_compound, num_new);
    b->num_mv_found = (uint8_t)num_mv_found;
}

// ---------------------------------------------------------------- local warp (7.11.3.8)
void TileDecoder::warp_estimation() {
    int64_t A[2][2] = {{0, 0}, {0, 0}}, Bx[2] = {0, 0}, By[2] = {0, 0};
    const int mid_y = mi_row * 4 + bh4 * 2 - 1, mid_x = mi_col * 4 + bw4 * 2 - 1;
    const int suy = mid_y * 8, sux = mid_x * 8;
    const int duy = suy + b->mv[0].row, dux = sux + b->mv[0].col;
    auto ls = [](int a, int c) -> int64_t { return (((int64_t)a * c) >> 2) + (a + c); };
    for (int i = 0; i < num_samples; i++) {
        const int sy = cand_list[i][0] - suy, sx = cand_list[i][1] - sux;
        const int dy = cand_list[i][2] - duy, dx = cand_list[i][3] - dux;
        if (std::abs(sx - dx) < LS_MV_MAX && std::abs(sy - dy) < LS_MV_MAX) {
            A[0][0] += ls(sx, sx) + 8;
            A[0][1] += ls(sx, sy) + 4;
            A[1][1] += ls(sy, sy) + 8;
            Bx[0] += ls(sx, dx) + 8;
            Bx[1] += ls(sy, dx) + 4;
            By[0] += ls(sx, dy) + 4;
            By[1] += ls(sy, dy) + 8;
        }
    }
    const int64_t det = A[0][0] * A[1][1] - A[0][1] * A[0][1];
    b->warp_valid = det != 0;
    if (!b->warp_valid) return;
    int div_shift, div_factor;
    resolve_divisor(det, div_shift, div_factor);
    div_shift -= WARPEDMODEL_PREC_BITS;
    int64_t factor = div_factor;
    if (div_shift < 0) {
        factor = factor * ((int64_t)1 << (-div_shift));
        div_shift = 0;
    }
    const int clampv = 1 << 13;
    auto nondiag = [&](int64_t v) { return (int32_t)std::max<int64_t>(-clampv + 1, std::min<int64_t>(clampv - 1, round2s64(v * factor, div_shift))); };
    auto diag = [&](int64_t v) {
        return (int32_t)std::max<int64_t>((1 << 16) - clampv + 1, std::min<int64_t>((1 << 16) + clampv - 1, round2s64(v * factor, div_shift)));
    };
    int32_t* p = b->warp;
    p[2] = diag(A[1][1] * Bx[0] - A[0][1] * Bx[1]);
    p[3] = nondiag(-A[0][1] * Bx[0] + A[0][0] * Bx[1]);
    p[4] = nondiag(A[1][1] * By[0] - A[0][1] * By[1]);
    p[5] = diag(-A[0][1] * By[0] + A[0][0] * By[1]);
    const int64_t vx = (int64_t)b->mv[0].col * (1 << (WARPEDMODEL_PREC_BITS - 3)) - ((int64_t)mid_x * (p[2] - (1 << 16)) + (int64_t)mid_y * p[3]);
    const int64_t vy = (int64_t)b->mv[0].row * (1 << (WARPEDMODEL_PREC_BITS - 3)) - ((int64_t)mid_x * p[4] + (int64_t)mid_y * (p[5] - (1 << 16)));
    const int64_t tc = 1 << 23;
    p[0] = (int32_t)std::max<int64_t>(-tc, std::min<int64_t>(tc - 1, vx));
    p[1] = (int32_t)std::max<int64_t>(-tc, std::min<int64_t>(tc - 1, vy));
}

// ---------------------------------------------------------------- K2 work-list
void TileDecoder::emit_inter_block() {
    const int subx = seq.subsampling_x, suby = seq.subsampling_y;
    const int is_compound = b->ref_frame[1] > INTRA_FRAME;
    const int bw = kBlockW[b->bsize], bh = kBlockH[b->bsize];
    auto slot_of = [&](int ref) -> int8_t { return (int8_t)fh.ref_frame_idx[ref - LAST_FRAME]; };
    auto base_rec = [&](const BlockInfo* src) {
        InterBlk r;
        memset(&r, 0, sizeof(r));
        r.bsize = src->bsize;
        r.ref[0] = slot_of(src->ref_frame[0]);
        r.ref[1] = src->ref_frame[1] > INTRA_FRAME ? slot_of(src->ref_frame[1]) : (int8_t)-1;
        r.filt[0] = src->interp_filter[0];
        r.filt[1] = src->interp_filter[1];
        for (int l = 0; l < 2; l++) {
            r.mv[l][0] = src->mv[l].row;
            r.mv[l][1] = src->mv[l].col;
        }
        r.warp[0] = r.warp[1] = -1;
        r.comp_type = COMPOUND_AVERAGE;
        return r;
    };
    if (b->motion_mode == WARPED_CAUSAL) {
        warp_estimation();
        if (b->warp_valid) {
            int16_t sh[4];
            b->warp_valid = (uint8_t)setup_shear(b->warp, sh);
        }
    }
    InterBlk r = base_rec(b);
    r.x = (uint16_t)(mi_col * 4);
    r.y = (uint16_t)(mi_row * 4);
    r.w = (uint8_t)bw;
    r.h = (uint8_t)bh;
    r.planes = 1;
    r.comp_type = b->compound_type;
    r.wedge_index = b->wedge_index;
    r.wedge_sign = b->wedge_sign;
    r.mask_type = b->mask_type;
    r.interintra = b->interintra;
    // warp models
    if (b->motion_mode == WARPED_CAUSAL && b->warp_valid) {
        WarpRec wr;
        memcpy(wr.mat, b->warp, sizeof(wr.mat));
        int16_t sh[4];
        setup_shear(b->warp, sh);
        wr.alpha = sh[0]; wr.beta = sh[1]; wr.gamma = sh[2]; wr.delta = sh[3];
        r.warp[0] = (int16_t)(8 + to.warps.size());   // rebased when the tiles are merged
        to.warps.push_back(wr);
    } else if ((b->y_mode == GLOBALMV || b->y_mode == GLOBAL_GLOBALMV) && std::min(bw, bh) >= 8) {
        for (int l = 0; l < 1 + is_compound; l++) {
            const int ref = b->ref_frame[l];
            if (fh.gm_type[ref] > GM_TRANSLATION && fw.gm_warp_valid[ref] && !fw.ref_scaled[ref]) r.warp[l] = (int16_t)ref;
        }
    }
    if (is_compound && b->compound_type == COMPOUND_DISTANCE) {
        int dist[2];
        for (int l = 0; l < 2; l++)
            dist[l] = clip3(0, MAX_FRAME_DISTANCE, std::abs(hp.get_relative_dist(fh.order_hints[b->ref_frame[l]], fh.order_hint)));
        const int d0 = dist[1], d1 = dist[0];
        const int order = d0 <= d1;
        int i;
        if (d0 == 0 || d1 == 0) {
            i = 3;
        } else {
            for (i = 0; i < 3; i++) {
                const int c0 = av1t_quant_dist_weight[i][order], c1 = av1t_quant_dist_weight[i][!order];
                if (order) { if (d0 * c0 > d1 * c1) break; }
                else { if (d0 * c0 < d1 * c1) break; }
            }
        }
        r.fwd_w = av1t_quant_dist_lookup[i][order];
        r.bck_w = av1t_quant_dist_lookup[i][1 - order];
    }
    // overlapped motion compensation neighbours (7.11.3.10)
    if (b->motion_mode == OBMC_CAUSAL) {
        r.obmc_first = (uint32_t)to.obmc.size();
        if (b->has_chroma) r.obmc_chroma_above = plane_residual_size((BlockSize)b->bsize, subx, suby) >= BLOCK_8X8;
        auto push_nb = [&](const BlockInfo* n, int x4, int y4, int step4) {
            ObmcNb o;
            o.x4 = (uint16_t)x4;
            o.y4 = (uint16_t)y4;
            o.step4 = (uint8_t)step4;
            o.ref = slot_of(n->ref_frame[0]);
            o.filt[0] = n->interp_filter[0];
            o.filt[1] = n->interp_filter[1];
            o.mv[0] = n->mv[0].row;
            o.mv[1] = n->mv[0].col;
            to.obmc.push_back(o);
        };
        if (avail_u) {
            int n_count = 0;
            const int n_limit = std::min(4, kBlockWLog2[b->bsize] - 2);
            for (int x4 = mi_col; n_count < n_limit && x4 < std::min(fw.mi_cols, mi_col + bw4);) {
                const BlockInfo* n = blk(mi_row - 1, x4 | 1);
                const int step4 = clip3(2, 16, kBlockW4[n->bsize]);
                if (n->ref_frame[0] > INTRA_FRAME) {
                    n_count++;
                    push_nb(n, x4, mi_row, step4);
                    r.obmc_above++;
                }
                x4 += step4;
            }
        }
        if (avail_l) {
            int n_count = 0;
            const int n_limit = std::min(4, kBlockHLog2[b->bsize] - 2);
            for (int y4 = mi_row; n_count < n_limit && y4 < std::min(fw.mi_rows, mi_row + bh4);) {
                const BlockInfo* n = blk(y4 | 1, mi_col - 1);
                const int step4 = clip3(2, 16, kBlockH4[n->bsize]);
                if (n->ref_frame[0] > INTRA_FRAME) {
                    n_count++;
                    push_nb(n, mi_col, y4, step4);
                    r.obmc_left++;
                }
                y4 += step4;
            }
        }
    }
    {
        uint32_t* th = to.tool_hist;
        th[TOOL_INTER_BLOCKS]++;
        if (is_compound) th[b->compound_type == COMPOUND_AVERAGE ? TOOL_COMPOUND_AVG : b->compound_type == COMPOUND_DISTANCE ? TOOL_COMPOUND_DIST
                            : b->compound_type == COMPOUND_WEDGE ? TOOL_COMPOUND_WEDGE : TOOL_COMPOUND_DIFFWTD]++;
        if (b->interintra) th[b->wedge_interintra ? TOOL_INTERINTRA_WEDGE : TOOL_INTERINTRA]++;
        if (b->motion_mode == OBMC_CAUSAL) th[TOOL_OBMC]++;
        if (b->motion_mode == WARPED_CAUSAL && b->warp_valid) th[TOOL_LOCAL_WARP]++;
        if (b->motion_mode != WARPED_CAUSAL && (r.warp[0] >= 0 || r.warp[1] >= 0)) th[TOOL_GLOBAL_WARP]++;
        if (b->skip_mode) th[TOOL_SKIP_MODE]++;
        if (b->interp_filter[0] != b->interp_filter[1]) th[TOOL_DUAL_FILTER]++;
        if (b->interp_filter[0] != INTERP_EIGHTTAP) th[TOOL_SWITCHABLE_FILTER]++;
        if (has_newmv(b->y_mode)) th[TOOL_NEWMV]++;
    }
    const int sub8 = (bw4 == 1 && subx) || (bh4 == 1 && suby);
    if (b->has_chroma && sub8) to.tool_hist[TOOL_SUB8X8_CHROMA]++;
    if (b->has_chroma && !sub8) r.planes |= 2;
    to.inter.push_back(r);
    {
        const uint64_t area = (uint64_t)bw * bh * ((r.planes & 2) ? 3 : 2) / 2;
        to.inter_samples += area;
        to.inter_ref_samples += area * (1 + is_compound);
    }
    if (b->has_chroma && sub8) {
        // chroma of a group of sub-8x8 luma blocks (spec 7.11.3.1 / compute_prediction)
        const int psz = plane_residual_size((BlockSize)b->bsize, subx, suby);
        const int n4w = kBlockW4[psz], n4h = kBlockH4[psz];
        const int cand_row = (mi_row >> suby) << suby, cand_col = (mi_col >> subx) << subx;
        int some_intra = 0;
        for (int rr = 0; rr < (n4h << suby); rr++)
            for (int cc = 0; cc < (n4w << subx); cc++) {
                const BlockInfo* n = blk(cand_row + rr, cand_col + cc);
                if (!n || n->ref_frame[0] == INTRA_FRAME) some_intra = 1;
            }
        if (some_intra) {
            InterBlk c = base_rec(b);
            c.x = (uint16_t)(cand_col * 4);
            c.y = (uint16_t)(cand_row * 4);
            c.w = (uint8_t)((n4w * 4) << subx);
            c.h = (uint8_t)((n4h * 4) << suby);
            c.planes = 2;
            to.inter.push_back(c);
        } else {
            for (int rr = 0, y = 0; y < n4h * 4; y += bh >> suby, rr++)
                for (int cc = 0, x = 0; x < n4w * 4; x += bw >> subx, cc++) {
                    const BlockInfo* n = blk(cand_row + rr, cand_col + cc);
                    InterBlk c = base_rec(n);
                    c.x = (uint16_t)((cand_col + cc) * 4);
                    c.y = (uint16_t)((cand_row + rr) * 4);
                    c.w = (uint8_t)bw;
                    c.h = (uint8_t)bh;
                    c.planes = 2;
                    to.inter.push_back(c);
                }
        }
        to.inter_samples += (uint64_t)(n4w * 4) * (n4h * 4) * 2;
        to.inter_ref_samples += (uint64_t)(n4w * 4) * (n4h * 4) * 2;
    }
    if (b->interintra) emit_interintra_records();
}

// ---------------------------------------------------------------- variable transform size
void TileDecoder::read_var_tx_size(int row, int col, int txsz, int depth) {
    if (row >= fw.mi_rows || col >= fw.mi_cols) return;
    int split = 0;
    if (txsz != TX_4X4 && depth != 2) {
        auto above_w = [&]() -> int {
            if (row == mi_row) {
                if (!avail_u) return 64;
                const BlockInfo* a = blk(row - 1, col);
                if (a->skip && a->is_inter) return kBlockW[a->bsize];
            }
            return kTxW[fw.inter_tx[(size_t)(row - 1) * fw.mi_cols + col]];
        };
        auto left_h = [&]() -> int {
            if (col == mi_col) {
                if (!avail_l) return 64;
                const BlockInfo* l = blk(row, col - 1);
                if (l->skip && l->is_inter) return kBlockH[l->bsize];
            }
            return kTxH[fw.inter_tx[(size_t)row * fw.mi_cols + col - 1]];
        };
        const int above = above_w() < kTxW[txsz], left = left_h() < kTxH[txsz];
        const int size = std::min(64, std::max((int)kBlockW[b->bsize], (int)kBlockH[b->bsize]));
        const int max_tx = find_tx_size(size, size);
        const int ctx = (kTxSqrUp[txsz] != max_tx) * 3 + (4 - max_tx) * 6 + above + left;
        split = ms.symbol(cdf.txfm_partition[ctx], 2);
        if (split) to.tool_hist[TOOL_VARTX_SPLIT]++;
    }
    const int w4 = kTxW[txsz] / 4, h4 = kTxH[txsz] / 4;
    if (split) {
        const int sub = kSplitTx[txsz];
        const int sw = kTxW[sub] / 4, sh = kTxH[sub] / 4;
        for (int i = 0; i < h4; i += sh)
            for (int j = 0; j < w4; j += sw) read_var_tx_size(row + i, col + j, sub, depth + 1);
    } else {
        for (int i = 0; i < h4; i++)
            for (int j = 0; j < w4; j++)
                if (row + i < fw.mi_rows && col + j < fw.mi_cols) fw.inter_tx[(size_t)(row + i) * fw.mi_cols + col + j] = (uint8_t)txsz;
        b->tx_size = (uint8_t)txsz;
    }
}

void TileDecoder::transform_tree(int start_x, int start_y, int w, int h) {
    const int max_x = fw.mi_cols * 4, max_y = fw.mi_rows * 4;
    if (start_x >= max_x || start_y >= max_y) return;
    const int row = start_y >> 2, col = start_x >> 2;
    const int luma_tx = fw.inter_tx[(size_t)row * fw.mi_cols + col];
    const int lw = kTxW[luma_tx], lh = kTxH[luma_tx];
    if (w <= lw && h <= lh) {
        transform_block(0, start_x, start_y, find_tx_size(w, h), 0, 0);
    } else if (w > h) {
        transform_tree(start_x, start_y, w / 2, h);
        transform_tree(start_x + w / 2, start_y, w / 2, h);
    } else if (w < h) {
        transform_tree(start_x, start_y, w, h / 2);
        transform_tree(start_x, start_y + h / 2, w, h / 2);
    } else {
        transform_tree(start_x, start_y, w / 2, h / 2);
        transform_tree(start_x + w / 2, start_y, w / 2, h / 2);
        transform_tree(start_x, start_y + h / 2, w / 2, h / 2);
        transform_tree(start_x + w / 2, start_y + h / 2, w / 2, h / 2);
    }
}

}  // namespace av1r

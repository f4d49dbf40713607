// Tile parse: partition tree, intra mode info, tx sizes/types, coefficients (AV1 spec 5.11).
// Host, sequential, not the optimised path; its output is the device work-list (worklist.h).
#include "tile.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/av1r.h"
#include "tables/tables_scan.inc"

namespace av1r {

static const uint16_t* default_scan(int txsz) {
    switch (txsz) {
        case TX_4X4: return av1t_default_scan_4x4;
        case TX_8X8: return av1t_default_scan_8x8;
        case TX_16X16: return av1t_default_scan_16x16;
        case TX_32X32: case TX_64X64: case TX_32X64: case TX_64X32: return av1t_default_scan_32x32;
        case TX_4X8: return av1t_default_scan_4x8;
        case TX_8X4: return av1t_default_scan_8x4;
        case TX_8X16: return av1t_default_scan_8x16;
        case TX_16X8: return av1t_default_scan_16x8;
        case TX_16X32: case TX_16X64: return av1t_default_scan_16x32;
        case TX_32X16: case TX_64X16: return av1t_default_scan_32x16;
        case TX_4X16: return av1t_default_scan_4x16;
        case TX_16X4: return av1t_default_scan_16x4;
        case TX_8X32: return av1t_default_scan_8x32;
        case TX_32X8: return av1t_default_scan_32x8;
    }
    return av1t_default_scan_4x4;
}
static const uint16_t* mrow_scan(int txsz) {
    switch (txsz) {
        case TX_4X4: return av1t_mrow_scan_4x4;
        case TX_8X8: return av1t_mrow_scan_8x8;
        case TX_16X16: return av1t_mrow_scan_16x16;
        case TX_4X8: return av1t_mrow_scan_4x8;
        case TX_8X4: return av1t_mrow_scan_8x4;
        case TX_8X16: return av1t_mrow_scan_8x16;
        case TX_16X8: return av1t_mrow_scan_16x8;
        case TX_4X16: return av1t_mrow_scan_4x16;
        case TX_16X4: return av1t_mrow_scan_16x4;
    }
    return nullptr;
}
static const uint16_t* mcol_scan(int txsz) {
    switch (txsz) {
        case TX_4X4: return av1t_mcol_scan_4x4;
        case TX_8X8: return av1t_mcol_scan_8x8;
        case TX_16X16: return av1t_mcol_scan_16x16;
        case TX_4X8: return av1t_mcol_scan_4x8;
        case TX_8X4: return av1t_mcol_scan_8x4;
        case TX_8X16: return av1t_mcol_scan_8x16;
        case TX_16X8: return av1t_mcol_scan_16x8;
        case TX_4X16: return av1t_mcol_scan_4x16;
        case TX_16X4: return av1t_mcol_scan_16x4;
    }
    return nullptr;
}
static const uint16_t* get_scan(int txsz, int txtp) {
    if (txsz == TX_16X64) return av1t_default_scan_16x32;
    if (txsz == TX_64X16) return av1t_default_scan_32x16;
    if (kTxSqrUp[txsz] == TX_64X64) return av1t_default_scan_32x32;
    if (txtp == IDTX) return default_scan(txsz);
    const uint16_t* s = nullptr;
    if (txtp == V_DCT || txtp == V_ADST || txtp == V_FLIPADST) s = mrow_scan(txsz);
    else if (txtp == H_DCT || txtp == H_ADST || txtp == H_FLIPADST) s = mcol_scan(txsz);
    return s ? s : default_scan(txsz);
}

TileDecoder::TileDecoder(const SeqHdr& s, const HeaderParser& h, FrameWork& f, TileOut& t, const CdfCtx& init_cdf)
    : cdf(init_cdf), seq(s), hp(h), fw(f), to(t), fh(f.fh) {
    for (int p = 0; p < 3; p++) {
        above_level[p].assign(fw.mi_cols + 64, 0);
        above_dc[p].assign(fw.mi_cols + 64, 0);
        left_level[p].assign(fw.mi_rows + 64, 0);
        left_dc[p].assign(fw.mi_rows + 64, 0);
    }
    above_seg_pred.assign(fw.mi_cols + 64, 0);
    left_seg_pred.assign(fw.mi_rows + 64, 0);
    memset(levels, 0, sizeof(levels));
    memset(levels3, 0, sizeof(levels3));
}

void TileDecoder::clear_block_decoded_flags(int r, int c, int sb4) {
    // spec 5.11.3: the row above and the column left of the superblock count as decoded where they lie inside the tile, everything
    // else starts undecoded.  Row-wise fills instead of a branch per cell (this runs once per superblock and plane).
    for (int plane = 0; plane < seq.num_planes; plane++) {
        const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        const int sbw4 = (mi_col_end - c) >> sx, sbh4 = (mi_row_end - r) >> sy;
        const int nx = (sb4 >> sx) + 2, ny = (sb4 >> sy) + 2;          // cells x, y = -1 .. sb4 >> s?
        uint8_t (*bd)[35] = block_decoded[plane];
        // y = -1: decoded for x < sbw4 (x = -1 included)
        const int n1 = std::min(nx, std::max(0, sbw4 + 1));
        memset(bd[0], 1, (size_t)n1);
        memset(bd[0] + n1, 0, (size_t)(nx - n1));
        for (int y = 0; y < ny - 1; y++) {
            memset(bd[y + 1], 0, (size_t)nx);
            bd[y + 1][0] = y < sbh4;                                   // x = -1
        }
        bd[(sb4 >> sy) + 1][0] = 0;
    }
}

int TileDecoder::decode_tile(const uint8_t* data, size_t sz, int tile_row, int tile_col) {
    mi_row_start = fh.mi_row_starts[tile_row];
    mi_row_end = fh.mi_row_starts[tile_row + 1];
    mi_col_start = fh.mi_col_starts[tile_col];
    mi_col_end = fh.mi_col_starts[tile_col + 1];
    current_q_index = fh.base_q_idx;
    for (int r = mi_row_start; r < mi_row_end; r++)   // "not yet decoded in this frame" for the tile's rectangle (see FrameWork::init)
        memset(&fw.mi[(size_t)r * fw.mi_cols + mi_col_start], 0, sizeof(fw.mi[0]) * (size_t)(mi_col_end - mi_col_start));
    ms.init(data, sz, fh.disable_cdf_update != 0);
    // clear_above_context
    for (int p = 0; p < 3; p++) {
        std::fill(above_level[p].begin(), above_level[p].end(), 0);
        std::fill(above_dc[p].begin(), above_dc[p].end(), 0);
    }
    std::fill(above_seg_pred.begin(), above_seg_pred.end(), 0);
    for (int i = 0; i < 4; i++) delta_lf[i] = 0;
    for (int p = 0; p < 3; p++) {
        ref_sgr_xqd[p][0] = -32;
        ref_sgr_xqd[p][1] = 31;
        for (int pass = 0; pass < 2; pass++) {
            ref_lr_wiener[p][pass][0] = 3;
            ref_lr_wiener[p][pass][1] = -7;
            ref_lr_wiener[p][pass][2] = 15;
        }
    }
    const int sb_size = seq.use_128x128_superblock ? BLOCK_128X128 : BLOCK_64X64;
    const int sb4 = kBlockW4[sb_size];
    const int sb_shift = seq.use_128x128_superblock ? 5 : 4;
    for (int r = mi_row_start; r < mi_row_end; r += sb4) {
        // clear_left_context
        for (int p = 0; p < 3; p++) {
            std::fill(left_level[p].begin(), left_level[p].end(), 0);
            std::fill(left_dc[p].begin(), left_dc[p].end(), 0);
        }
        std::fill(left_seg_pred.begin(), left_seg_pred.end(), 0);
        for (int c = mi_col_start; c < mi_col_end; c += sb4) {
            read_deltas = fh.delta_q_present;
            // clear_cdef
            {
                int c64 = (fw.mi_cols + 15) >> 4;
                int n = seq.use_128x128_superblock ? 2 : 1;
                for (int yy = 0; yy < n; yy++)
                    for (int xx = 0; xx < n; xx++) {
                        int rr = (r >> 4) + yy, cc = (c >> 4) + xx;
                        if (rr < ((fw.mi_rows + 15) >> 4) && cc < c64) fw.cdef_idx[(size_t)rr * c64 + cc] = -1;
                    }
            }
            clear_block_decoded_flags(r, c, sb4);
            memset(&cur_sb, 0, sizeof(cur_sb));
            cur_sb.sb_row = (uint16_t)(r >> sb_shift);
            cur_sb.sb_col = (uint16_t)(c >> sb_shift);
            cur_sb.tile_sb_col0 = (uint16_t)(mi_col_start >> sb_shift);
            cur_sb.tile_sb_col1 = (uint16_t)((mi_col_end + sb4 - 1) >> sb_shift);
            cur_sb.tile_sb_row0 = (uint16_t)(mi_row_start >> sb_shift);
            cur_unit = -1;
            read_lr(r, c, sb_size);
            if (!decode_partition(r, c, sb_size)) return fail_code ? fail_code : AV1R_EBITSTREAM;
        }
    }
    return fail_code;
}

// ---------------------------------------------------------------- loop restoration side info
int TileDecoder::decode_subexp_bool(int num_syms, int k) {
    int i = 0, mk = 0;
    while (true) {
        int b2 = i ? k + i - 1 : k;
        int a = 1 << b2;
        if (num_syms <= mk + 3 * a) {
            // ns(num_syms - mk) coded with bools
            int n = num_syms - mk;
            int w = 0, x = n;
            while (x) { x >>= 1; w++; }
            int m = (1 << w) - n;
            int v = ms.literal(w - 1);
            if (v < m) return v + mk;
            int extra = ms.literal(1);
            return (v << 1) - m + extra + mk;
        }
        if (ms.literal(1)) {
            i++;
            mk += a;
        } else {
            return ms.literal(b2) + mk;
        }
    }
}

static int inv_recenter(int r, int v) {
    if (v > 2 * r) return v;
    if (v & 1) return r - ((v + 1) >> 1);
    return r + (v >> 1);
}

int TileDecoder::decode_signed_subexp_with_ref_bool(int low, int high, int k, int r) {
    int mx = high - low;
    int rr = r - low;
    int v = decode_subexp_bool(mx, k);
    int x;
    if ((rr << 1) <= mx) x = inv_recenter(rr, v);
    else x = mx - 1 - inv_recenter(mx - 1 - rr, v);
    return x + low;
}

static int count_units_in_frame(int unit_size, int frame_size) { return std::max((frame_size + (unit_size >> 1)) / unit_size, 1); }

void TileDecoder::read_lr(int r, int c, int bsize) {
    if (fh.allow_intrabc) return;
    int w = kBlockW4[bsize], h = kBlockH4[bsize];
    for (int plane = 0; plane < seq.num_planes; plane++) {
        if (fh.lr_type[plane] == RESTORE_NONE) continue;
        int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        int unit_size = fh.lr_size[plane];
        int unit_rows = count_units_in_frame(unit_size, (fh.frame_height + sy) >> sy);
        int unit_cols = count_units_in_frame(unit_size, (fh.upscaled_width + sx) >> sx);
        int unit_row_start = (r * (4 >> sy) + unit_size - 1) / unit_size;
        int unit_row_end = ((r + h) * (4 >> sy) + unit_size - 1) / unit_size;
        int numerator, denominator;
        if (fh.use_superres) {
            numerator = (4 >> sx) * fh.superres_denom;
            denominator = unit_size * 8;
        } else {
            numerator = 4 >> sx;
            denominator = unit_size;
        }
        int unit_col_start = (c * numerator + denominator - 1) / denominator;
        int unit_col_end = ((c + w) * numerator + denominator - 1) / denominator;
        unit_row_end = std::min(unit_row_end, unit_rows);
        unit_col_end = std::min(unit_col_end, unit_cols);
        for (int ur = unit_row_start; ur < unit_row_end; ur++)
            for (int uc = unit_col_start; uc < unit_col_end; uc++) read_lr_unit(plane, ur, uc);
    }
}

void TileDecoder::read_lr_unit(int plane, int ur, int uc) {
    static const int taps_min[3] = {-5, -23, -17}, taps_max[3] = {10, 8, 46}, taps_k[3] = {1, 2, 3};
    static const int xqd_min[2] = {-96, -32}, xqd_max[2] = {31, 95};
    static const uint8_t sgr_r[16][2] = {{2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1},
                                         {2, 1}, {2, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {2, 0}, {2, 0}};
    LrUnit& u = fw.lr[plane][(size_t)ur * fw.lr_cols[plane] + uc];
    int type;
    if (fh.lr_type[plane] == RESTORE_WIENER) type = ms.symbol(cdf.wiener_restore, 2) ? RESTORE_WIENER : RESTORE_NONE;
    else if (fh.lr_type[plane] == RESTORE_SGRPROJ) type = ms.symbol(cdf.sgrproj_restore, 2) ? RESTORE_SGRPROJ : RESTORE_NONE;
    else type = ms.symbol(cdf.switchable_restore, 3);   // 0 none, 1 wiener, 2 sgrproj
    u.type = (uint8_t)type;
    if (type == RESTORE_WIENER) {
        for (int pass = 0; pass < 2; pass++) {
            int first = 0;
            if (plane) {
                first = 1;
                u.wiener[pass][0] = 0;
            }
            for (int j = first; j < 3; j++) {
                int v = decode_signed_subexp_with_ref_bool(taps_min[j], taps_max[j] + 1, taps_k[j], ref_lr_wiener[plane][pass][j]);
                u.wiener[pass][j] = (int8_t)v;
                ref_lr_wiener[plane][pass][j] = v;
            }
        }
    } else if (type == RESTORE_SGRPROJ) {
        int set = ms.literal(4);
        u.sgr_set = (uint8_t)set;
        for (int i = 0; i < 2; i++) {
            int radius = sgr_r[set][i];
            int v;
            if (radius) {
                v = decode_signed_subexp_with_ref_bool(xqd_min[i], xqd_max[i] + 1, 4, ref_sgr_xqd[plane][i]);
            } else {
                v = 0;
                if (i == 1) v = std::max(xqd_min[1], std::min(xqd_max[1], 128 - ref_sgr_xqd[plane][0]));
            }
            u.sgr_xqd[i] = (int8_t)v;
            ref_sgr_xqd[plane][i] = v;
        }
    }
}

// ---------------------------------------------------------------- partition tree
bool TileDecoder::decode_partition(int r, int c, int bsize) {
    if (r >= fw.mi_rows || c >= fw.mi_cols) return true;
    if (fail_code) return false;
    const bool au = is_inside(r - 1, c), al = is_inside(r, c - 1);
    const int num4 = kBlockW4[bsize];
    const int half = num4 >> 1, quarter = half >> 1;
    const bool has_rows = (r + half) < fw.mi_rows, has_cols = (c + half) < fw.mi_cols;
    int partition;
    if (bsize < BLOCK_8X8) {
        partition = PARTITION_NONE;
    } else {
        const int bsl = kBlockWLog2[bsize] - 3;   // 8x8 -> 0 ... 128x128 -> 4
        const int above = au && (kBlockWLog2[blk(r - 1, c)->bsize] < kBlockWLog2[bsize]);
        const int left = al && (kBlockHLog2[blk(r, c - 1)->bsize] < kBlockHLog2[bsize]);
        uint16_t* pc = cdf.partition[bsl * 4 + left * 2 + above];
        const int nsym = bsl == 0 ? 4 : (bsl == 4 ? 8 : 10);
        if (has_rows && has_cols) {
            partition = ms.symbol(pc, nsym);
        } else {
            // inverted cdf: pc[i] = 32768 - CDF(i); P(sym == i) = pc[i-1] - pc[i], pc[-1] = 32768
            auto prob = [&](int i) -> int { return (i ? pc[i - 1] : 32768) - pc[i]; };
            if (has_cols) {   // split_or_horz
                int psum = prob(PARTITION_VERT) + prob(PARTITION_SPLIT);
                if (bsize != BLOCK_8X8) psum += prob(PARTITION_HORZ_A) + prob(PARTITION_VERT_A) + prob(PARTITION_VERT_B);
                if (bsize != BLOCK_8X8 && bsize != BLOCK_128X128) psum += prob(PARTITION_VERT_4);
                // symbol 1 (SPLIT) has probability psum: inverted cdf[0] = P(sym > 0) = psum
                partition = ms.bool_icdf((uint32_t)psum) ? PARTITION_SPLIT : PARTITION_HORZ;
            } else if (has_rows) {   // split_or_vert
                int psum = prob(PARTITION_HORZ) + prob(PARTITION_SPLIT);
                if (bsize != BLOCK_8X8) psum += prob(PARTITION_HORZ_A) + prob(PARTITION_HORZ_B) + prob(PARTITION_VERT_A);
                if (bsize != BLOCK_8X8 && bsize != BLOCK_128X128) psum += prob(PARTITION_HORZ_4);
                partition = ms.bool_icdf((uint32_t)psum) ? PARTITION_SPLIT : PARTITION_VERT;
            } else {
                partition = PARTITION_SPLIT;
            }
        }
    }
    const int sub = partition_subsize(partition, (BlockSize)bsize);
    const int split = partition_subsize(PARTITION_SPLIT, (BlockSize)bsize);
    if (sub == BLOCK_INVALID) return fail(AV1R_EBITSTREAM, "invalid partition for block size");
    switch (partition) {
        case PARTITION_NONE: return decode_block(r, c, sub);
        case PARTITION_HORZ:
            if (!decode_block(r, c, sub)) return false;
            if (has_rows) return decode_block(r + half, c, sub);
            return true;
        case PARTITION_VERT:
            if (!decode_block(r, c, sub)) return false;
            if (has_cols) return decode_block(r, c + half, sub);
            return true;
        case PARTITION_SPLIT:
            return decode_partition(r, c, sub) && decode_partition(r, c + half, sub) && decode_partition(r + half, c, sub) &&
                   decode_partition(r + half, c + half, sub);
        case PARTITION_HORZ_A:
            return decode_block(r, c, split) && decode_block(r, c + half, split) && decode_block(r + half, c, sub);
        case PARTITION_HORZ_B:
            return decode_block(r, c, sub) && decode_block(r + half, c, split) && decode_block(r + half, c + half, split);
        case PARTITION_VERT_A:
            return decode_block(r, c, split) && decode_block(r + half, c, split) && decode_block(r, c + half, sub);
        case PARTITION_VERT_B:
            return decode_block(r, c, sub) && decode_block(r, c + half, split) && decode_block(r + half, c + half, split);
        case PARTITION_HORZ_4:
            if (!decode_block(r, c, sub) || !decode_block(r + quarter, c, sub) || !decode_block(r + quarter * 2, c, sub)) return false;
            if (r + quarter * 3 < fw.mi_rows) return decode_block(r + quarter * 3, c, sub);
            return true;
        case PARTITION_VERT_4:
            if (!decode_block(r, c, sub) || !decode_block(r, c + quarter, sub) || !decode_block(r, c + quarter * 2, sub)) return false;
            if (c + quarter * 3 < fw.mi_cols) return decode_block(r, c + quarter * 3, sub);
            return true;
    }
    return false;
}

// ---------------------------------------------------------------- block
bool TileDecoder::decode_block(int r, int c, int bsize) {
    if (fail_code) return false;
    to.blocks.emplace_back();
    b = &to.blocks.back();
    memset(b, 0, sizeof(*b));
    mi_row = r;
    mi_col = c;
    b->mi_row = (uint16_t)r;
    b->mi_col = (uint16_t)c;
    b->bsize = (uint8_t)bsize;
    bw4 = kBlockW4[bsize];
    bh4 = kBlockH4[bsize];
    int has_chroma;
    if (bh4 == 1 && seq.subsampling_y && (r & 1) == 0) has_chroma = 0;
    else if (bw4 == 1 && seq.subsampling_x && (c & 1) == 0) has_chroma = 0;
    else has_chroma = seq.num_planes > 1;
    b->has_chroma = (uint8_t)has_chroma;
    avail_u = is_inside(r - 1, c);
    avail_l = is_inside(r, c - 1);
    avail_u_chroma = avail_u;
    avail_l_chroma = avail_l;
    if (has_chroma) {
        if (seq.subsampling_y && bh4 == 1) avail_u_chroma = is_inside(r - 2, c);
        if (seq.subsampling_x && bw4 == 1) avail_l_chroma = is_inside(r, c - 2);
    } else {
        avail_u_chroma = avail_l_chroma = 0;
    }
    b->ref_frame[0] = INTRA_FRAME;
    b->ref_frame[1] = -1;
    ftype_cache[0] = ftype_cache[1] = -1;
    if (fh.frame_is_intra) intra_frame_mode_info();
    else inter_frame_mode_info();
    if (fail_code) return false;
    if (b->pal_size[0] || b->pal_size[1]) {
        to.tool_hist[TOOL_PALETTE]++;
        palette_tokens();
    }
    read_block_tx_size();
    if (b->skip) reset_block_context();
    b->qidx = (uint8_t)get_qidx(fh, 0, b->segment_id, current_q_index);
    for (int i = 0; i < 4; i++) b->delta_lf[i] = (int8_t)delta_lf[i];
    if (fh.lf.level[0] || fh.lf.level[1]) {
        // filter level of the block per edge class (spec 7.14.4), computed once here instead of per 4x4 cell later
        for (int i = 0; i < 4; i++) {
            const int dlf = fh.delta_lf_multi ? b->delta_lf[i] : b->delta_lf[0];
            int lvl = std::max(0, std::min(63, dlf + fh.lf.level[i]));
            if (fh.seg.enabled && fh.seg.feature_enabled[b->segment_id][1 + i])
                lvl = std::max(0, std::min(63, lvl + fh.seg.feature_data[b->segment_id][1 + i]));
            if (fh.lf.delta_enabled) {
                const int nshift = lvl >> 5;
                const int ref = b->ref_frame[0];
                if (ref == INTRA_FRAME) {
                    lvl += fh.lf.ref_deltas[INTRA_FRAME] << nshift;
                } else {
                    const int mode = b->y_mode;
                    const int mode_type = (mode >= NEARESTMV && mode != GLOBALMV && mode != GLOBAL_GLOBALMV) ? 1 : 0;
                    lvl += (fh.lf.ref_deltas[ref] << nshift) + (fh.lf.mode_deltas[mode_type] << nshift);
                }
                lvl = std::max(0, std::min(63, lvl));
            }
            b->lf_lvl[i] = (uint8_t)lvl;
        }
    }
    // publish the block in the per-mi maps (clipped to the frame)
    const int rmax = std::min(r + bh4, fw.mi_rows), cmax = std::min(c + bw4, fw.mi_cols);
    const bool lf_frame = fh.lf.level[0] || fh.lf.level[1];
    if (fw.host_lf) {
        LfMi lm;
        memcpy(lm.lvl, b->lf_lvl, 4);
        lm.bsize = b->bsize;
        lm.filt_inside = (uint8_t)(!b->skip || b->ref_frame[0] <= INTRA_FRAME);
        lm.valid = 1;
        lm.pad = 0;
        // four per-mi maps, one contiguous run per row each: library fills (vectorised) instead of a four-stream scalar loop
        const size_t n = (size_t)(cmax - c);
        for (int y = r; y < rmax; y++) {
            const size_t o = (size_t)y * fw.mi_cols + c;
            std::fill_n(&fw.mi[o], n, b);
            memset(&fw.skip_mi[o], b->skip, n);
            memset(&fw.seg_ids[o], b->segment_id, n);
            std::fill_n(&fw.lf_mi[o], n, lm);
        }
    } else {
        // device-side edge classification: the block goes into a 12-byte list, the per-mi level map is not needed on the host
        if (lf_frame) {
            LfBlk lb;
            lb.mi_row = (uint16_t)r;
            lb.mi_col = (uint16_t)c;
            lb.bsize = b->bsize;
            lb.filt_inside = (uint8_t)(!b->skip || b->ref_frame[0] <= INTRA_FRAME);
            memcpy(lb.lvl, b->lf_lvl, 4);
            lb.pad = 0;
            to.lf_blocks.push_back(lb);
        }
        const size_t n = (size_t)(cmax - c);
        for (int y = r; y < rmax; y++) {
            const size_t o = (size_t)y * fw.mi_cols + c;
            std::fill_n(&fw.mi[o], n, b);
            memset(&fw.skip_mi[o], b->skip, n);
            memset(&fw.seg_ids[o], b->segment_id, n);
        }
    }
    if (b->is_inter && !b->use_intrabc) emit_inter_block();   // (intra block copy is predicted by the wavefront kernel, record by record)
    residual();
    return fail_code == 0;
}

void TileDecoder::intra_frame_mode_info() {
    b->skip = 0;
    if (fh.seg.seg_id_pre_skip) intra_segment_id();
    b->skip_mode = 0;
    read_skip();
    if (!fh.seg.seg_id_pre_skip) intra_segment_id();
    read_cdef();
    read_delta_qindex();
    read_delta_lf();
    read_deltas = 0;
    b->ref_frame[0] = INTRA_FRAME;
    b->ref_frame[1] = -1;
    if (fh.allow_intrabc) {
        b->use_intrabc = (uint8_t)ms.symbol(cdf.intrabc, 2);
    }
    if (b->use_intrabc) {
        // spec 5.11.7: the block is coded like a single-reference NEWMV inter block whose reference is the frame being
        // decoded (before any filter); bilinear interpolation, no palette, DC_PRED for the neighbours' contexts
        b->is_inter = 1;
        b->y_mode = DC_PRED;
        b->uv_mode = DC_PRED;
        b->motion_mode = 0;
        b->compound_type = COMPOUND_AVERAGE;
        b->interp_filter[0] = b->interp_filter[1] = INTERP_BILINEAR;
        to.tool_hist[TOOL_INTRABC]++;
        find_mv_stack(0);
        assign_dv();
        return;
    }
    b->is_inter = 0;
    const int above_mode = avail_u ? blk(mi_row - 1, mi_col)->y_mode : DC_PRED;
    const int left_mode = avail_l ? blk(mi_row, mi_col - 1)->y_mode : DC_PRED;
    const int actx = kIntraModeCtx[above_mode < INTRA_MODES ? above_mode : DC_PRED];
    const int lctx = kIntraModeCtx[left_mode < INTRA_MODES ? left_mode : DC_PRED];
    b->y_mode = (uint8_t)ms.symbol(cdf.kf_y_mode[actx][lctx], 13);
    intra_angle_info_y();
    intra_mode_tail();
}

void TileDecoder::intra_mode_tail() {
    if (b->has_chroma) {
        int cfl_allowed;
        if (b->lossless && plane_residual_size((BlockSize)b->bsize, seq.subsampling_x, seq.subsampling_y) == BLOCK_4X4) cfl_allowed = 1;
        else if (!b->lossless && std::max(kBlockW[b->bsize], kBlockH[b->bsize]) <= 32) cfl_allowed = 1;
        else cfl_allowed = 0;
        b->uv_mode = (uint8_t)ms.symbol(cdf.uv_mode[cfl_allowed][b->y_mode], 13 + cfl_allowed);
        if (b->uv_mode == UV_CFL_PRED) read_cfl_alphas();
        intra_angle_info_uv();
    }
    b->pal_size[0] = b->pal_size[1] = 0;
    if (b->bsize >= BLOCK_8X8 && kBlockW[b->bsize] <= 64 && kBlockH[b->bsize] <= 64 && fh.allow_screen_content_tools) {
        palette_mode_info();
    }
    filter_intra_mode_info();
}

// ---------------------------------------------------------------- palette (spec 5.11.46, 5.11.49, 7.11.4)
int TileDecoder::get_palette_cache(int plane, uint16_t* cache) const {
    int above_n = 0, left_n = 0;
    const BlockInfo *a = nullptr, *l = nullptr;
    if (((mi_row * 4) % 64) && avail_u) {
        a = blk(mi_row - 1, mi_col);
        above_n = a->pal_size[plane];
    }
    if (avail_l) {
        l = blk(mi_row, mi_col - 1);
        left_n = l->pal_size[plane];
    }
    int ai = 0, li = 0, n = 0;
    while (ai < above_n && li < left_n) {
        const int ac = a->pal_colors[plane][ai], lc = l->pal_colors[plane][li];
        if (lc < ac) {
            if (n == 0 || lc != cache[n - 1]) cache[n++] = (uint16_t)lc;
            li++;
        } else {
            if (n == 0 || ac != cache[n - 1]) cache[n++] = (uint16_t)ac;
            ai++;
            if (lc == ac) li++;
        }
    }
    while (ai < above_n) {
        const int v = a->pal_colors[plane][ai++];
        if (n == 0 || v != cache[n - 1]) cache[n++] = (uint16_t)v;
    }
    while (li < left_n) {
        const int v = l->pal_colors[plane][li++];
        if (n == 0 || v != cache[n - 1]) cache[n++] = (uint16_t)v;
    }
    return n;
}

static int ceil_log2(int x) {
    if (x < 2) return 0;
    int i = 1, p = 2;
    while (p < x) { i++; p <<= 1; }
    return i;
}

void TileDecoder::palette_mode_info() {
    const int bd = seq.bit_depth, pixmax = (1 << bd) - 1;
    const int bsize_ctx = kBlockWLog2[b->bsize] + kBlockHLog2[b->bsize] - 6;   // Mi_Width_Log2 + Mi_Height_Log2 - 2
    uint16_t cache[16];
    if (b->y_mode == DC_PRED) {
        int ctx = 0;
        if (avail_u && blk(mi_row - 1, mi_col)->pal_size[0] > 0) ctx++;
        if (avail_l && blk(mi_row, mi_col - 1)->pal_size[0] > 0) ctx++;
        if (ms.symbol(cdf.palette_y_mode[bsize_ctx][ctx], 2)) {
            const int n = ms.symbol(cdf.palette_y_size[bsize_ctx], 7) + 2;
            b->pal_size[0] = (uint8_t)n;
            uint16_t* col = b->pal_colors[0];
            const int cache_n = get_palette_cache(0, cache);
            int idx = 0;
            for (int i = 0; i < cache_n && idx < n; i++)
                if (ms.literal(1)) col[idx++] = cache[i];
            if (idx < n) {
                col[idx++] = (uint16_t)ms.literal(bd);
                if (idx < n) {
                    int bits = bd - 3 + ms.literal(2);
                    while (idx < n) {
                        const int delta = ms.literal(bits) + 1;
                        col[idx] = (uint16_t)std::min(pixmax, col[idx - 1] + delta);
                        const int range = (1 << bd) - col[idx] - 1;
                        bits = std::min(bits, ceil_log2(range));
                        idx++;
                    }
                }
            }
            std::sort(col, col + n);
        }
    }
    if (b->has_chroma && b->uv_mode == DC_PRED) {
        const int ctx = b->pal_size[0] > 0;
        if (ms.symbol(cdf.palette_uv_mode[ctx], 2)) {
            const int n = ms.symbol(cdf.palette_uv_size[bsize_ctx], 7) + 2;
            b->pal_size[1] = (uint8_t)n;
            uint16_t* cu = b->pal_colors[1];
            uint16_t* cv = b->pal_colors[2];
            const int cache_n = get_palette_cache(1, cache);
            int idx = 0;
            for (int i = 0; i < cache_n && idx < n; i++)
                if (ms.literal(1)) cu[idx++] = cache[i];
            if (idx < n) {
                cu[idx++] = (uint16_t)ms.literal(bd);
                if (idx < n) {
                    int bits = bd - 3 + ms.literal(2);
                    while (idx < n) {
                        const int delta = ms.literal(bits);
                        cu[idx] = (uint16_t)std::min(pixmax, cu[idx - 1] + delta);
                        const int range = (1 << bd) - cu[idx];
                        bits = std::min(bits, ceil_log2(range));
                        idx++;
                    }
                }
            }
            std::sort(cu, cu + n);
            if (ms.literal(1)) {   // delta_encode_palette_colors_v
                const int max_val = 1 << bd;
                const int bits = bd - 4 + ms.literal(2);
                cv[0] = (uint16_t)ms.literal(bd);
                for (idx = 1; idx < n; idx++) {
                    int delta = ms.literal(bits);
                    if (delta && ms.literal(1)) delta = -delta;
                    int val = cv[idx - 1] + delta;
                    if (val < 0) val += max_val;
                    if (val >= max_val) val -= max_val;
                    cv[idx] = (uint16_t)std::min(pixmax, std::max(0, val));
                }
            } else {
                for (idx = 0; idx < n; idx++) cv[idx] = (uint16_t)ms.literal(bd);
            }
        }
    }
}

// Colour index maps.  Layout of one palette entry in FrameWork::pal (TxRec::pal_off points at it):
//   uint16 colours[8] | uint16 origin_x, origin_y (plane samples) | uint16 stride | uint16 0 | uint8 map[rows * stride]
void TileDecoder::palette_tokens() {
    static const int8_t kColorCtx[9] = {-1, -1, 0, -1, -1, 4, 3, 2, 1};
    for (int pi = 0; pi < 2; pi++) {
        const int n = b->pal_size[pi];
        pal_entry[pi] = pal_entry[2] = 0;
        if (!n) continue;
        const int sx = pi ? seq.subsampling_x : 0, sy = pi ? seq.subsampling_y : 0;
        int bw = kBlockW[b->bsize] >> sx, bh = kBlockH[b->bsize] >> sy;
        int ow = std::min((int)kBlockW[b->bsize], (fw.mi_cols - mi_col) * 4) >> sx, oh = std::min((int)kBlockH[b->bsize], (fw.mi_rows - mi_row) * 4) >> sy;
        if (pi && bw < 4) { bw += 2; ow += 2; }
        if (pi && bh < 4) { bh += 2; oh += 2; }
        std::vector<uint8_t> map((size_t)bw * bh, 0);
        auto at = [&](int r, int c) -> uint8_t& { return map[(size_t)r * bw + c]; };
        {   // color_index_map: NS(n)
            int w = 0, x = n;
            while (x) { x >>= 1; w++; }
            const int m = (1 << w) - n;
            int v = ms.literal(w - 1);
            if (v >= m) v = (v << 1) - m + ms.literal(1);
            at(0, 0) = (uint8_t)v;
        }
        for (int i = 1; i < oh + ow - 1; i++)
            for (int j = std::min(i, ow - 1); j >= std::max(0, i - oh + 1); j--) {
                const int r = i - j, c = j;
                int scores[8] = {0, 0, 0, 0, 0, 0, 0, 0}, order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                if (c > 0) scores[at(r, c - 1)] += 2;
                if (r > 0 && c > 0) scores[at(r - 1, c - 1)] += 1;
                if (r > 0) scores[at(r - 1, c)] += 2;
                for (int a = 0; a < 3; a++) {
                    int max_score = scores[a], max_idx = a;
                    for (int k = a + 1; k < n; k++)
                        if (scores[k] > max_score) { max_score = scores[k]; max_idx = k; }
                    if (max_idx != a) {
                        const int mo = order[max_idx];
                        for (int k = max_idx; k > a; k--) { scores[k] = scores[k - 1]; order[k] = order[k - 1]; }
                        scores[a] = max_score;
                        order[a] = mo;
                    }
                }
                const int hash = scores[0] * 1 + scores[1] * 2 + scores[2] * 2;
                const int ctx = kColorCtx[hash];
                uint16_t* c_ = pi ? cdf.palette_uv_color_index[n - 2][ctx] : cdf.palette_y_color_index[n - 2][ctx];
                int sym;
                switch (n) {   // the symbol decoder wants a compile-time alphabet size
                    case 2: sym = ms.symbol(c_, 2); break;
                    case 3: sym = ms.symbol(c_, 3); break;
                    case 4: sym = ms.symbol(c_, 4); break;
                    case 5: sym = ms.symbol(c_, 5); break;
                    case 6: sym = ms.symbol(c_, 6); break;
                    case 7: sym = ms.symbol(c_, 7); break;
                    default: sym = ms.symbol(c_, 8); break;
                }
                at(r, c) = (uint8_t)order[sym];
            }
        for (int i = 0; i < oh; i++)
            for (int j = ow; j < bw; j++) at(i, j) = at(i, ow - 1);
        for (int i = oh; i < bh; i++)
            for (int j = 0; j < bw; j++) at(i, j) = at(oh - 1, j);
        // emit one entry per plane (U and V share the map)
        for (int plane = pi ? 1 : 0; plane <= (pi ? 2 : 0); plane++) {
            while (to.pal.size() & 3) to.pal.push_back(0);
            pal_entry[plane] = (uint32_t)to.pal.size();
            uint16_t hdr[12];
            for (int k = 0; k < 8; k++) hdr[k] = b->pal_colors[plane][k];
            hdr[8] = (uint16_t)((mi_col >> sx) * 4);
            hdr[9] = (uint16_t)((mi_row >> sy) * 4);
            hdr[10] = (uint16_t)bw;
            hdr[11] = 0;
            const uint8_t* hb = reinterpret_cast<const uint8_t*>(hdr);
            to.pal.insert(to.pal.end(), hb, hb + sizeof(hdr));
            to.pal.insert(to.pal.end(), map.begin(), map.end());
        }
    }
}

void TileDecoder::intra_segment_id() {
    if (fh.seg.enabled) read_segment_id();
    else b->segment_id = 0;
    b->lossless = (uint8_t)fh.lossless_array[b->segment_id];
}

static int neg_deinterleave(int diff, int ref, int max) {
    if (!ref) return diff;
    if (ref >= (max - 1)) return max - diff - 1;
    if (2 * ref < max) {
        if (diff <= 2 * ref) {
            if (diff & 1) return ref + ((diff + 1) >> 1);
            return ref - (diff >> 1);
        }
        return diff;
    }
    if (diff <= 2 * (max - ref - 1)) {
        if (diff & 1) return ref + ((diff + 1) >> 1);
        return ref - (diff >> 1);
    }
    return max - (diff + 1);
}

void TileDecoder::read_segment_id() {
    int prev_ul = -1, prev_u = -1, prev_l = -1;
    if (avail_u && avail_l) prev_ul = fw.seg_ids[(size_t)(mi_row - 1) * fw.mi_cols + mi_col - 1];
    if (avail_u) prev_u = fw.seg_ids[(size_t)(mi_row - 1) * fw.mi_cols + mi_col];
    if (avail_l) prev_l = fw.seg_ids[(size_t)mi_row * fw.mi_cols + mi_col - 1];
    int pred;
    if (prev_u == -1) pred = prev_l == -1 ? 0 : prev_l;
    else if (prev_l == -1) pred = prev_u;
    else pred = (prev_ul == prev_u) ? prev_u : prev_l;
    if (b->skip) {
        b->segment_id = (uint8_t)pred;
    } else {
        int ctx;
        if (prev_ul < 0) ctx = 0;
        else if (prev_ul == prev_u && prev_ul == prev_l) ctx = 2;
        else if (prev_ul == prev_u || prev_ul == prev_l || prev_u == prev_l) ctx = 1;
        else ctx = 0;
        int v = ms.symbol(cdf.seg_spatial[ctx], 8);
        v = neg_deinterleave(v, pred, fh.seg.last_active_seg_id + 1);
        b->segment_id = (uint8_t)std::max(0, std::min(fh.seg.last_active_seg_id, v));
    }
}

void TileDecoder::read_skip() {
    if (fh.seg.seg_id_pre_skip && fh.seg.enabled && fh.seg.feature_enabled[b->segment_id][SEG_LVL_SKIP]) {
        b->skip = 1;
        return;
    }
    int ctx = 0;
    if (avail_u) ctx += blk(mi_row - 1, mi_col)->skip;
    if (avail_l) ctx += blk(mi_row, mi_col - 1)->skip;
    b->skip = (uint8_t)ms.symbol(cdf.skip[ctx], 2);
}

void TileDecoder::read_cdef() {
    if (b->skip || fh.coded_lossless || !seq.enable_cdef || fh.allow_intrabc) return;
    const int c64 = (fw.mi_cols + 15) >> 4, r64 = (fw.mi_rows + 15) >> 4;
    const int r = mi_row >> 4, c = mi_col >> 4;
    int8_t& idx = fw.cdef_idx[(size_t)r * c64 + c];
    if (idx == -1) {
        idx = (int8_t)ms.literal(fh.cdef_bits);
        for (int y = r; y < std::min(r64, (mi_row + bh4 + 15) >> 4); y++)
            for (int x = c; x < std::min(c64, (mi_col + bw4 + 15) >> 4); x++) fw.cdef_idx[(size_t)y * c64 + x] = idx;
    }
}

void TileDecoder::read_delta_qindex() {
    const int sb_size = seq.use_128x128_superblock ? BLOCK_128X128 : BLOCK_64X64;
    if (b->bsize == sb_size && b->skip) return;
    if (read_deltas) {
        int abs_ = ms.symbol(cdf.delta_q, 4);
        if (abs_ == 3) {
            int rem_bits = ms.literal(3) + 1;
            abs_ = ms.literal(rem_bits) + (1 << rem_bits) + 1;
        }
        if (abs_) {
            int sign = ms.literal(1);
            int reduced = sign ? -abs_ : abs_;
            current_q_index = std::max(1, std::min(255, current_q_index + (reduced << fh.delta_q_res)));
        }
    }
}

void TileDecoder::read_delta_lf() {
    const int sb_size = seq.use_128x128_superblock ? BLOCK_128X128 : BLOCK_64X64;
    if (b->bsize == sb_size && b->skip) return;
    if (read_deltas && fh.delta_lf_present) {
        int count = 1;
        if (fh.delta_lf_multi) count = seq.num_planes > 1 ? 4 : 2;
        for (int i = 0; i < count; i++) {
            uint16_t* c = fh.delta_lf_multi ? cdf.delta_lf_multi[i] : cdf.delta_lf;
            int abs_ = ms.symbol(c, 4);
            if (abs_ == 3) {
                int n = ms.literal(3) + 1;
                abs_ = ms.literal(n) + (1 << n) + 1;
            }
            if (abs_) {
                int sign = ms.literal(1);
                int reduced = sign ? -abs_ : abs_;
                delta_lf[i] = std::max(-63, std::min(63, delta_lf[i] + (reduced << fh.delta_lf_res)));
            }
        }
    }
}

void TileDecoder::intra_angle_info_y() {
    b->angle_y = 0;
    if (b->bsize >= BLOCK_8X8 && is_directional_mode(b->y_mode))
        b->angle_y = (int8_t)(ms.symbol(cdf.angle_delta[b->y_mode - V_PRED], 7) - 3);
}

void TileDecoder::intra_angle_info_uv() {
    b->angle_uv = 0;
    if (b->bsize >= BLOCK_8X8 && is_directional_mode(b->uv_mode))
        b->angle_uv = (int8_t)(ms.symbol(cdf.angle_delta[b->uv_mode - V_PRED], 7) - 3);
}

void TileDecoder::read_cfl_alphas() {
    int signs = ms.symbol(cdf.cfl_sign, 8);
    int sign_u = (signs + 1) / 3, sign_v = (signs + 1) % 3;
    if (sign_u) {
        int ctx = (sign_u - 1) * 3 + sign_v;
        int a = 1 + ms.symbol(cdf.cfl_alpha[ctx], 16);
        b->cfl_alpha_u = (int8_t)(sign_u == 1 ? -a : a);
    } else {
        b->cfl_alpha_u = 0;
    }
    if (sign_v) {
        int ctx = (sign_v - 1) * 3 + sign_u;
        int a = 1 + ms.symbol(cdf.cfl_alpha[ctx], 16);
        b->cfl_alpha_v = (int8_t)(sign_v == 1 ? -a : a);
    } else {
        b->cfl_alpha_v = 0;
    }
}

void TileDecoder::filter_intra_mode_info() {
    b->use_filter_intra = 0;
    if (seq.enable_filter_intra && b->y_mode == DC_PRED && b->pal_size[0] == 0 && std::max(kBlockW[b->bsize], kBlockH[b->bsize]) <= 32) {
        b->use_filter_intra = (uint8_t)ms.symbol(cdf.filter_intra[b->bsize], 2);
        if (b->use_filter_intra) b->fi_mode = (uint8_t)ms.symbol(cdf.filter_intra_mode, 5);
    }
}

// ---------------------------------------------------------------- transform size
void TileDecoder::read_tx_size(int allow_select) {
    if (b->lossless) {
        b->tx_size = TX_4X4;
        return;
    }
    const int max_rect = kMaxTxRect[b->bsize];
    const int max_depth = kMaxTxDepth[b->bsize];
    int tx = max_rect;
    if (b->bsize > BLOCK_4X4 && allow_select && fh.tx_mode == TX_MODE_SELECT) {
        const int max_w = kTxW[max_rect], max_h = kTxH[max_rect];
        int above_w = 0, left_h = 0;
        if (avail_u) {
            const BlockInfo* a = blk(mi_row - 1, mi_col);
            if (a->is_inter) above_w = kBlockW[a->bsize];
            else if (a->skip && a->is_inter) above_w = kBlockW[a->bsize];
            else above_w = kTxW[fw.inter_tx[(size_t)(mi_row - 1) * fw.mi_cols + mi_col]];
        }
        if (avail_l) {
            const BlockInfo* l = blk(mi_row, mi_col - 1);
            if (l->is_inter) left_h = kBlockH[l->bsize];
            else left_h = kTxH[fw.inter_tx[(size_t)mi_row * fw.mi_cols + mi_col - 1]];
        }
        const int ctx = (above_w >= max_w) + (left_h >= max_h);
        const int cat = max_depth - 1;
        const int depth = ms.symbol(cdf.tx_size[cat][ctx], cat == 0 ? 2 : 3);
        for (int i = 0; i < depth; i++) tx = kSplitTx[tx];
    }
    b->tx_size = (uint8_t)tx;
}

void TileDecoder::read_block_tx_size() {
    if (fh.tx_mode == TX_MODE_SELECT && b->bsize > BLOCK_4X4 && b->is_inter && !b->skip && !b->lossless) {
        const int max_tx = kMaxTxRect[b->bsize];
        const int tw = kTxW[max_tx] / 4, th = kTxH[max_tx] / 4;
        for (int row = mi_row; row < mi_row + bh4; row += th)
            for (int col = mi_col; col < mi_col + bw4; col += tw) read_var_tx_size(row, col, max_tx, 0);
        b->tx_size = (uint8_t)max_tx;
    } else {
        read_tx_size(!b->skip || !b->is_inter);
        const int rmax = std::min(mi_row + bh4, fw.mi_rows), cmax = std::min(mi_col + bw4, fw.mi_cols);
        for (int row = mi_row; row < rmax; row++)
            for (int col = mi_col; col < cmax; col++) fw.inter_tx[(size_t)row * fw.mi_cols + col] = b->tx_size;
    }
}

void TileDecoder::reset_block_context() {
    for (int plane = 0; plane < 1 + 2 * b->has_chroma; plane++) {
        int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        for (int i = mi_col >> sx; i <= ((mi_col + bw4 - 1) >> sx); i++) {
            above_level[plane][i] = 0;
            above_dc[plane][i] = 0;
        }
        for (int i = mi_row >> sy; i <= ((mi_row + bh4 - 1) >> sy); i++) {
            left_level[plane][i] = 0;
            left_dc[plane][i] = 0;
        }
    }
}

// ---------------------------------------------------------------- residual
static int uv_tx_size(int bsize, int sx, int sy) {
    int uv = kMaxTxRect[plane_residual_size((BlockSize)bsize, sx, sy)];
    if (kTxW[uv] == 64 || kTxH[uv] == 64) {
        if (kTxW[uv] == 16) return TX_16X32;
        if (kTxH[uv] == 16) return TX_32X16;
        return TX_32X32;
    }
    return uv;
}

// A skipped inter block codes no coefficient and emits no record: all its walk over the transform blocks leaves behind is the
// transform size in the deblocking maps and the "decoded" flags the intra edge availability of later blocks looks at.  Both are
// rectangles (the transform size of a skipped inter block is uniform), written here with row fills instead of one
// transform_block() call per transform block (a million calls per 4K clip of 60 frames).
void TileDecoder::mark_skipped_inter_block() {
    const int sb_mask = seq.use_128x128_superblock ? 31 : 15;
    for (int plane = 0; plane < 1 + b->has_chroma * 2; plane++) {
        const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        const int txsz = plane ? uv_tx_size(b->bsize, sx, sy) : b->tx_size;
        const int step_x = kTxW[txsz] >> 2, step_y = kTxH[txsz] >> 2;
        const int plane_sz = plane_residual_size((BlockSize)b->bsize, sx, sy);
        const int x0 = mi_col >> sx, y0 = mi_row >> sy;
        const int pw4 = fw.plane_w4(plane), ph4 = fw.plane_h4(plane);
        // transform blocks that start inside the frame (the others are not visited)
        const int cw = std::min((int)kBlockW4[plane_sz], pw4 - x0), chh = std::min((int)kBlockH4[plane_sz], ph4 - y0);
        if (cw <= 0 || chh <= 0) continue;
        const int cw_tx = (cw + step_x - 1) / step_x * step_x, ch_tx = (chh + step_y - 1) / step_y * step_y;
        const int wv = std::min(cw_tx, pw4 - x0), hv = std::min(ch_tx, ph4 - y0);
        for (int i = 0; i < hv; i++) memset(&fw.lf_tx[plane][(size_t)(y0 + i) * pw4 + x0], txsz, (size_t)wv);
        const int bx0 = (((x0 << sx) & sb_mask) >> sx) + 1, by0 = (((y0 << sy) & sb_mask) >> sy) + 1;
        const int n = std::min(cw_tx, 35 - bx0);
        for (int i = 0; i < ch_tx && by0 + i < 35 && n > 0; i++) memset(&block_decoded[plane][by0 + i][bx0], 1, (size_t)n);
    }
}

void TileDecoder::residual() {
    const int sb_mask = seq.use_128x128_superblock ? 31 : 15;
    (void)sb_mask;
    if (b->skip && b->is_inter && !b->use_intrabc && !b->lossless) {
        mark_skipped_inter_block();
        return;
    }
    const int width_chunks = std::max(1, kBlockW[b->bsize] >> 6), height_chunks = std::max(1, kBlockH[b->bsize] >> 6);
    const int mi_size_chunk = (width_chunks > 1 || height_chunks > 1) ? BLOCK_64X64 : b->bsize;
    for (int chunk_y = 0; chunk_y < height_chunks; chunk_y++)
        for (int chunk_x = 0; chunk_x < width_chunks; chunk_x++) {
            const int mi_row_chunk = mi_row + (chunk_y << 4), mi_col_chunk = mi_col + (chunk_x << 4);
            for (int plane = 0; plane < 1 + b->has_chroma * 2; plane++) {
                const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
                const int txsz = b->lossless ? TX_4X4 : (plane ? uv_tx_size(b->bsize, sx, sy) : b->tx_size);
                const int step_x = kTxW[txsz] >> 2, step_y = kTxH[txsz] >> 2;
                const int plane_sz = plane_residual_size((BlockSize)mi_size_chunk, sx, sy);
                const int num4w = kBlockW4[plane_sz], num4h = kBlockH4[plane_sz];
                const int base_x = (mi_col_chunk >> sx) * 4, base_y = (mi_row_chunk >> sy) * 4;
                if (b->is_inter && !b->lossless && !plane) {
                    transform_tree(base_x, base_y, num4w * 4, num4h * 4);
                } else {
                    const int base_xb = (mi_col >> sx) * 4, base_yb = (mi_row >> sy) * 4;
                    for (int y = 0; y < num4h; y += step_y)
                        for (int x = 0; x < num4w; x += step_x)
                            transform_block(plane, base_xb, base_yb, txsz, x + ((chunk_x << 4) >> sx), y + ((chunk_y << 4) >> sy));
                }
                if (fail_code) return;
            }
        }
}

int TileDecoder::filter_type(int plane) const {
    auto is_smooth = [&](int r, int c) -> int {
        const BlockInfo* n = blk(r, c);
        int mode;
        if (plane == 0) {
            mode = n->y_mode;
        } else {
            if (n->ref_frame[0] > INTRA_FRAME) return 0;
            mode = n->uv_mode;
        }
        return mode == SMOOTH_PRED || mode == SMOOTH_V_PRED || mode == SMOOTH_H_PRED;
    };
    int above_smooth = 0, left_smooth = 0;
    if (plane == 0 ? avail_u : avail_u_chroma) {
        int r = mi_row - 1, c = mi_col;
        if (plane > 0) {
            if (seq.subsampling_x && !(mi_col & 1)) c++;
            if (seq.subsampling_y && (mi_row & 1)) r--;
        }
        if (c < fw.mi_cols && r >= 0 && blk(r, c)) above_smooth = is_smooth(r, c);
    }
    if (plane == 0 ? avail_l : avail_l_chroma) {
        int r = mi_row, c = mi_col - 1;
        if (plane > 0) {
            if (seq.subsampling_x && (mi_col & 1)) c--;
            if (seq.subsampling_y && !(mi_row & 1)) r++;
        }
        if (r < fw.mi_rows && c >= 0 && blk(r, c)) left_smooth = is_smooth(r, c);
    }
    return above_smooth || left_smooth;
}

void TileDecoder::transform_block(int plane, int base_x, int base_y, int txsz, int x, int y) {
    const int start_x = base_x + 4 * x, start_y = base_y + 4 * y;
    const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
    const int row = (start_y << sy) >> 2, col = (start_x << sx) >> 2;
    const int sb_mask = seq.use_128x128_superblock ? 31 : 15;
    const int sub_row = row & sb_mask, sub_col = col & sb_mask;
    const int step_x = kTxW[txsz] >> 2, step_y = kTxH[txsz] >> 2;
    const int max_x = (fw.mi_cols * 4) >> sx, max_y = (fw.mi_rows * 4) >> sy;
    if (start_x >= max_x || start_y >= max_y) return;
    TxRec rec;
    memset(&rec, 0, sizeof(rec));
    rec.x4 = (uint16_t)(start_x >> 2);
    rec.y4 = (uint16_t)(start_y >> 2);
    rec.plane = (uint8_t)plane;
    rec.txsz = (uint8_t)txsz;
    rec.qidx = b->qidx;
    rec.seg_id = b->segment_id;
    rec.qm_level = (uint8_t)fh.seg_qm_level[plane][b->segment_id];
    rec.coef_off = (uint32_t)to.coefs.size();
    if (b->lossless) rec.flags |= TXF_LOSSLESS;
    if (!b->is_inter) {
        const int have_left = (plane == 0 ? avail_l : avail_l_chroma) || start_x > base_x;
        const int have_above = (plane == 0 ? avail_u : avail_u_chroma) || start_y > base_y;
        const int have_ar = block_decoded[plane][(sub_row >> sy) - 1 + 1][(sub_col >> sx) + step_x + 1];
        const int have_bl = block_decoded[plane][(sub_row >> sy) + step_y + 1][(sub_col >> sx) - 1 + 1];
        if (have_left) rec.flags |= TXF_HAVE_LEFT;
        if (have_above) rec.flags |= TXF_HAVE_ABOVE;
        if (have_ar) rec.flags |= TXF_HAVE_ABOVE_RIGHT;
        if (have_bl) rec.flags |= TXF_HAVE_BELOW_LEFT;
        if (seq.enable_intra_edge_filter) {
            int& ft = ftype_cache[plane > 0];
            if (ft < 0) ft = filter_type(plane);
            if (ft) rec.flags |= TXF_SMOOTH_EDGE;
        }
        if (b->pal_size[plane > 0]) {
            rec.mode = TXM_PALETTE;
            rec.pal_off = pal_entry[plane];
            if (plane == 0) {
                max_luma_w = start_x + step_x * 4;
                max_luma_h = start_y + step_y * 4;
            }
        } else if (plane == 0) {
            if (b->use_filter_intra) {
                rec.mode = TXM_FILTER_INTRA;
                rec.fi_mode = b->fi_mode;
            } else {
                rec.mode = b->y_mode;
                rec.angle_delta = b->angle_y;
            }
            max_luma_w = start_x + step_x * 4;
            max_luma_h = start_y + step_y * 4;
        } else if (b->uv_mode == UV_CFL_PRED) {
            rec.mode = TXM_CFL;
            rec.cfl_alpha = plane == 1 ? b->cfl_alpha_u : b->cfl_alpha_v;
            rec.cfl_max_w4 = (uint16_t)(max_luma_w >> 2);
            rec.cfl_max_h4 = (uint16_t)(max_luma_h >> 2);
        } else {
            rec.mode = b->uv_mode;
            rec.angle_delta = b->angle_uv;
        }
    } else if (b->use_intrabc) {
        // every transform block of an intra-block-copy block is its own predict + reconstruct record of the wavefront kernel:
        // the predictor is a pure function of the sample position and the block vector (chroma uses the block's own vector,
        // also for sub-8x8 groups: spec 7.11.3.1 finds RefFrame[0] == INTRA_FRAME there)
        rec.mode = TXM_INTRABC;
        rec.cfl_max_w4 = (uint16_t)b->mv[0].col;
        rec.cfl_max_h4 = (uint16_t)b->mv[0].row;
    } else {
        rec.mode = TXM_INTER;
    }
    int eob = 0;
    if (!b->skip) {
        eob = coeffs(plane, start_x, start_y, txsz, rec);
        if (fail_code) return;
    }
    rec.eob = (uint16_t)eob;
    rec.ntok = (uint16_t)(to.coefs.size() - rec.coef_off);
    if (b->is_inter && b->interintra) rec.flags |= TXF_II;
    if (!b->is_inter || eob > 0 || b->use_intrabc) push_record(rec, (start_x << sx) >> 6, (start_y << sy) >> 6);
    if (eob > 0) to.coded_samples += (uint64_t)kTxW[txsz] * kTxH[txsz];
    // LoopfilterTxSizes + BlockDecoded
    const int pw4 = fw.plane_w4(plane), ph4 = fw.plane_h4(plane);
    for (int i = 0; i < step_y; i++)
        for (int j = 0; j < step_x; j++) {
            const int yy = (row >> sy) + i, xx = (col >> sx) + j;
            if (yy < ph4 && xx < pw4) fw.lf_tx[plane][(size_t)yy * pw4 + xx] = (uint8_t)txsz;
            const int by = (sub_row >> sy) + i + 1, bx = (sub_col >> sx) + j + 1;
            if (by < 35 && bx < 35) block_decoded[plane][by][bx] = 1;
        }
}

// records are grouped by 64x64 luma unit (decode order visits each unit contiguously)
void TileDecoder::push_record(const TxRec& rec, int ux, int uy) {
    const int key = (uy << 16) | ux;
    if (key != cur_unit) {
        cur_unit = key;
        SbRange sr = cur_sb;
        sr.first = (uint32_t)to.tx.size();
        sr.count = 0;
        sr.ux = (uint16_t)ux;
        sr.uy = (uint16_t)uy;
        to.sbs.push_back(sr);
    }
    to.sbs.back().count++;
    to.tx.push_back(rec);
    to.tx_blocks++;
}

// inter-intra (spec 7.11.3.1 / compute_prediction): one record per plane asks the wavefront kernel to form the intra
// predictor of the whole block from reconstructed neighbours and blend it over the inter predictor
void TileDecoder::emit_interintra_records() {
    static const uint8_t ii_to_intra[4] = {DC_PRED, V_PRED, H_PRED, SMOOTH_PRED};
    const int sb_mask = seq.use_128x128_superblock ? 31 : 15;
    const int sub_row = mi_row & sb_mask, sub_col = mi_col & sb_mask;
    for (int plane = 0; plane < 1 + 2 * b->has_chroma; plane++) {
        const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        const int psz = plane_residual_size((BlockSize)b->bsize, sx, sy);
        const int n4w = kBlockW4[psz], n4h = kBlockH4[psz];
        int txsz = TX_4X4;
        for (int t = 0; t < TX_SIZES_ALL; t++)
            if (kTxW[t] == n4w * 4 && kTxH[t] == n4h * 4) txsz = t;
        TxRec rec;
        memset(&rec, 0, sizeof(rec));
        rec.x4 = (uint16_t)(mi_col >> sx);
        rec.y4 = (uint16_t)(mi_row >> sy);
        rec.plane = (uint8_t)plane;
        rec.txsz = (uint8_t)txsz;
        rec.mode = ii_to_intra[b->interintra_mode];
        rec.flags = TXF_II;
        if (plane == 0 ? avail_l : avail_l_chroma) rec.flags |= TXF_HAVE_LEFT;
        if (plane == 0 ? avail_u : avail_u_chroma) rec.flags |= TXF_HAVE_ABOVE;
        if (block_decoded[plane][(sub_row >> sy) - 1 + 1][(sub_col >> sx) + n4w + 1]) rec.flags |= TXF_HAVE_ABOVE_RIGHT;
        if (block_decoded[plane][(sub_row >> sy) + n4h + 1][(sub_col >> sx) - 1 + 1]) rec.flags |= TXF_HAVE_BELOW_LEFT;
        rec.cfl_alpha = ii_pack(b->wedge_interintra, b->wedge_index, b->interintra_mode, b->bsize);
        rec.coef_off = (uint32_t)to.coefs.size();
        push_record(rec, (mi_col * 4) >> 6, (mi_row * 4) >> 6);
    }
}

int TileDecoder::get_tx_set(int txsz) const {
    const int sqr = kTxSqr[txsz], sqr_up = kTxSqrUp[txsz];
    if (sqr_up > TX_32X32) return 0;
    if (b->is_inter) {
        if (fh.reduced_tx_set || sqr_up == TX_32X32) return 3;
        if (sqr == TX_16X16) return 2;
        return 1;
    }
    if (sqr_up == TX_32X32) return 0;
    if (fh.reduced_tx_set) return 2;
    if (sqr == TX_16X16) return 2;
    return 1;
}

void TileDecoder::read_transform_type(int x4, int y4, int txsz) {
    const int set = get_tx_set(txsz);
    int tx_type = DCT_DCT;
    const int qidx = fh.seg.enabled ? get_qidx(fh, 1, b->segment_id, 0) : fh.base_q_idx;
    if (set > 0 && qidx > 0) {
        const int sqr = kTxSqr[txsz];
        if (b->is_inter) {
            if (set == 1) tx_type = kTxTypeInterInvSet1[ms.symbol(cdf.inter_ext_tx[1][sqr], 16)];
            else if (set == 2) tx_type = kTxTypeInterInvSet2[ms.symbol(cdf.inter_ext_tx[2][sqr], 12)];
            else tx_type = kTxTypeInterInvSet3[ms.symbol(cdf.inter_ext_tx[3][sqr], 2)];
        } else {
            const int intra_dir = b->use_filter_intra ? kFilterIntraModeToIntraDir[b->fi_mode] : b->y_mode;
            if (set == 1) tx_type = kTxTypeIntraInvSet1[ms.symbol(cdf.intra_ext_tx[1][sqr][intra_dir], 7)];
            else tx_type = kTxTypeIntraInvSet2[ms.symbol(cdf.intra_ext_tx[2][sqr][intra_dir], 5)];
        }
    }
    const int w4 = kTxW[txsz] >> 2, h4 = kTxH[txsz] >> 2;
    for (int j = 0; j < h4; j++)
        for (int i = 0; i < w4; i++)
            if (y4 + j < fw.mi_rows && x4 + i < fw.mi_cols) fw.tx_types[(size_t)(y4 + j) * fw.mi_cols + x4 + i] = (uint8_t)tx_type;
}

static bool tx_type_in_set(int set, bool is_inter, int t) {
    if (set == 0) return t == DCT_DCT;
    if (is_inter) {
        if (set == 1) return true;
        if (set == 2) return t <= IDTX || t == V_DCT || t == H_DCT;
        return t == IDTX || t == DCT_DCT;
    }
    if (set == 1) return t == IDTX || t == DCT_DCT || t == V_DCT || t == H_DCT || t == ADST_ADST || t == ADST_DCT || t == DCT_ADST;
    return t == IDTX || t == DCT_DCT || t == ADST_ADST || t == ADST_DCT || t == DCT_ADST;
}

int TileDecoder::compute_tx_type(int plane, int txsz, int block_x, int block_y) const {
    const int sqr_up = kTxSqrUp[txsz];
    if (b->lossless) return WHT_WHT;
    if (sqr_up > TX_32X32) return DCT_DCT;
    const int set = get_tx_set(txsz);
    if (plane == 0) return fw.tx_types[(size_t)block_y * fw.mi_cols + block_x];
    if (b->is_inter) {
        const int x4 = std::max(mi_col, block_x << seq.subsampling_x), y4 = std::max(mi_row, block_y << seq.subsampling_y);
        const int t = fw.tx_types[(size_t)y4 * fw.mi_cols + x4];
        return tx_type_in_set(set, true, t) ? t : DCT_DCT;
    }
    const int t = kModeToTxfm[b->uv_mode];
    return tx_type_in_set(set, false, t) ? t : DCT_DCT;
}

int TileDecoder::coeffs(int plane, int start_x, int start_y, int txsz, TxRec& rec) {
    const int x4 = start_x >> 2, y4 = start_y >> 2;
    const int w4 = kTxW[txsz] >> 2, h4 = kTxH[txsz] >> 2;
    const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
    const int tx_ctx = (kTxSqr[txsz] + kTxSqrUp[txsz] + 1) >> 1;
    const int ptype = plane > 0;
    int max_x4 = fw.mi_cols, max_y4 = fw.mi_rows;
    if (plane) {
        max_x4 = (max_x4 + sx) >> sx;
        max_y4 = (max_y4 + sy) >> sy;
    }
    // ---- all_zero context
    int ctx;
    {
        const int w = kTxW[txsz], h = kTxH[txsz];
        const int bsz = plane_residual_size((BlockSize)b->bsize, sx, sy);
        const int bw = kBlockW[bsz], bh = kBlockH[bsz];
        if (plane == 0) {
            int top = 0, left = 0;
            for (int k = 0; k < w4; k++)
                if (x4 + k < max_x4) top = std::max(top, (int)above_level[0][x4 + k]);
            for (int k = 0; k < h4; k++)
                if (y4 + k < max_y4) left = std::max(left, (int)left_level[0][y4 + k]);
            top = std::min(top, 255);
            left = std::min(left, 255);
            if (bw == w && bh == h) ctx = 0;
            else if (top == 0 && left == 0) ctx = 1;
            else if (top == 0 || left == 0) ctx = 2 + (std::max(top, left) > 3);
            else if (std::max(top, left) <= 3) ctx = 4;
            else if (std::min(top, left) <= 3) ctx = 5;
            else ctx = 6;
        } else {
            int above = 0, left = 0;
            for (int k = 0; k < w4; k++)
                if (x4 + k < max_x4) above |= above_level[plane][x4 + k] | above_dc[plane][x4 + k];
            for (int k = 0; k < h4; k++)
                if (y4 + k < max_y4) left |= left_level[plane][y4 + k] | left_dc[plane][y4 + k];
            ctx = (above != 0) + (left != 0);
            ctx += 7;
            if (bw * bh > w * h) ctx += 3;
        }
    }
    // the arithmetic decoder runs from a local copy inside this function: the byte stores of the level maps (and the vector stores
    // of the CDF adaptation) may alias anything behind `this`, and would force its window / range / count through memory per symbol
    Msac m = ms;
    const int all_zero = m.symbol(cdf.txb_skip[tx_ctx][ctx], 2);
    int eob = 0, cul_level = 0, dc_category = 0;
    if (all_zero) {
        if (plane == 0) {
            for (int j = 0; j < h4; j++)
                for (int i = 0; i < w4; i++)
                    if (y4 + j < fw.mi_rows && x4 + i < fw.mi_cols) fw.tx_types[(size_t)(y4 + j) * fw.mi_cols + x4 + i] = DCT_DCT;
        }
        rec.txtp = b->lossless ? WHT_WHT : DCT_DCT;
    } else {
        if (plane == 0) {
            ms = m;
            read_transform_type(x4, y4, txsz);
            m = ms;
        }
        const int txtp = compute_tx_type(plane, txsz, x4, y4);
        rec.txtp = (uint8_t)txtp;
        const int cls = tx_class_of(txtp);
        const uint16_t* scan = get_scan(txsz, txtp);
        const int adj = kAdjTx[txsz];
        const int bwl = kTxWLog2[adj];
        const int width = 1 << bwl, height = kTxH[adj];
        const int eob_multi = std::min((int)kTxWLog2[txsz], 5) + std::min((int)kTxHLog2[txsz], 5) - 4;
        const int eob_ctx = cls == TX_CLASS_2D ? 0 : 1;
        int eob_pt;
        switch (eob_multi) {
            case 0: eob_pt = m.symbol(cdf.eob_pt_16[ptype][eob_ctx], 5) + 1; break;
            case 1: eob_pt = m.symbol(cdf.eob_pt_32[ptype][eob_ctx], 6) + 1; break;
            case 2: eob_pt = m.symbol(cdf.eob_pt_64[ptype][eob_ctx], 7) + 1; break;
            case 3: eob_pt = m.symbol(cdf.eob_pt_128[ptype][eob_ctx], 8) + 1; break;
            case 4: eob_pt = m.symbol(cdf.eob_pt_256[ptype][eob_ctx], 9) + 1; break;
            case 5: eob_pt = m.symbol(cdf.eob_pt_512[ptype][eob_ctx], 10) + 1; break;
            default: eob_pt = m.symbol(cdf.eob_pt_1024[ptype][eob_ctx], 11) + 1; break;
        }
        eob = eob_pt < 2 ? eob_pt : ((1 << (eob_pt - 2)) + 1);
        int eob_shift = eob_pt >= 3 ? eob_pt - 3 : -1;
        if (eob_shift >= 0) {
            if (m.symbol(cdf.eob_extra[tx_ctx][ptype][eob_pt - 3], 2)) eob += 1 << eob_shift;
            for (int i = 1; i < std::max(0, eob_pt - 2); i++) {
                eob_shift = std::max(0, eob_pt - 2) - 1 - i;
                if (m.literal(1)) eob += 1 << eob_shift;
            }
        }
        if (eob > width * height) { ms = m; fail(AV1R_EBITSTREAM, "eob exceeds transform size"); return 0; }
        // levels, reverse scan.  lv[] / lv3[] are zero-padded (stride = width + 4) byte maps of min(level, 15) and min(level, 3):
        // the neighbour sums of the context derivation need neither bounds checks nor clamps; the positions of the non-zero
        // levels are chained in nz[] so that the sign pass (forward scan) visits and clears only those.
        const int ls = width + 4;
        uint8_t* lv = levels;
        uint8_t* lv3 = levels3;
        uint16_t nz[1024];
        int nnz = 0;
        uint16_t (*cb_cdf)[5] = cdf.coeff_base[tx_ctx][ptype];
        uint16_t (*br_cdf)[5] = cdf.coeff_br[std::min(tx_ctx, 3)][ptype];
        {   // last coefficient of the scan: coeff_base_eob
            const int c = eob - 1;
            const int pos = scan[c];
            const int row = pos >> bwl, col = pos - (row << bwl);
            int ectx;
            if (c == 0) ectx = 0;
            else if (c <= (height << bwl) / 8) ectx = 1;
            else if (c <= (height << bwl) / 4) ectx = 2;
            else ectx = 3;
            int level = m.symbol(cdf.coeff_base_eob[tx_ctx][ptype][ectx], 3) + 1;
            if (level > 2) {
                int rctx;
                if (pos == 0) rctx = 0;
                else if (cls == TX_CLASS_2D) rctx = (row < 2 && col < 2) ? 7 : 14;
                else if (cls == TX_CLASS_HORIZ) rctx = col == 0 ? 7 : 14;
                else rctx = row == 0 ? 7 : 14;
                uint16_t* bc = br_cdf[rctx];
                for (int idx = 0; idx < 4; idx++) {
                    const int br = m.symbol(bc, 4);
                    level += br;
                    if (br < 3) break;
                }
            }
            const int off = pos + (row << 2);
            lv[off] = (uint8_t)level;
            lv3[off] = (uint8_t)std::min(level, 3);
            nz[nnz++] = (uint16_t)pos;
        }
        auto pass1 = [&](auto cls_c) {
            constexpr int CLS = decltype(cls_c)::value;
            // neighbour offsets of the base context (5 positions) and of the base-range context (3 positions), spec 8.3.2
            const int s2 = CLS == TX_CLASS_2D ? ls + 1 : (CLS == TX_CLASS_HORIZ ? 2 : 2 * ls);
            const int s3 = CLS == TX_CLASS_2D ? 2 : (CLS == TX_CLASS_HORIZ ? 3 : 3 * ls);
            const int s4 = CLS == TX_CLASS_2D ? 2 * ls : (CLS == TX_CLASS_HORIZ ? 4 : 4 * ls);
            const int m2 = CLS == TX_CLASS_2D ? ls + 1 : (CLS == TX_CLASS_HORIZ ? 2 : 2 * ls);
            for (int c = eob - 2; c >= 0; c--) {
                const int pos = scan[c];
                const int row = pos >> bwl, col = pos - (row << bwl);
                const int off = pos + (row << 2);
                const uint8_t* l3 = lv3 + off;
                const int mag3 = l3[1] + l3[ls] + l3[s2] + l3[s3] + l3[s4];
                int bctx = std::min((mag3 + 1) >> 1, 4);
                if (CLS == TX_CLASS_2D) {
                    if (pos == 0) bctx = 0;
                    else bctx += av1t_coeff_base_ctx_offset[txsz][std::min(row, 4)][std::min(col, 4)];
                } else {
                    const int idx = CLS == TX_CLASS_VERT ? row : col;
                    bctx += 26 + 5 * std::min(idx, 2);
                }
                int level = m.symbol(cb_cdf[bctx], 4);
                if (level == 0) continue;
                if (level > 2) {
                    const uint8_t* lp = lv + off;
                    const int mag = std::min((lp[1] + lp[ls] + lp[m2] + 1) >> 1, 6);
                    int rctx;
                    if (pos == 0) rctx = mag;
                    else if (CLS == TX_CLASS_2D) rctx = (row < 2 && col < 2) ? mag + 7 : mag + 14;
                    else if (CLS == TX_CLASS_HORIZ) rctx = col == 0 ? mag + 7 : mag + 14;
                    else rctx = row == 0 ? mag + 7 : mag + 14;
                    uint16_t* bc = br_cdf[rctx];
                    for (int idx = 0; idx < 4; idx++) {
                        const int br = m.symbol(bc, 4);
                        level += br;
                        if (br < 3) break;
                    }
                }
                lv[off] = (uint8_t)level;
                lv3[off] = (uint8_t)std::min(level, 3);
                nz[nnz++] = (uint16_t)pos;
            }
        };
        if (cls == TX_CLASS_2D) pass1(std::integral_constant<int, TX_CLASS_2D>());
        else if (cls == TX_CLASS_HORIZ) pass1(std::integral_constant<int, TX_CLASS_HORIZ>());
        else pass1(std::integral_constant<int, TX_CLASS_VERT>());
        // signs + golomb, forward scan over the non-zero levels (tokens are written through a raw pointer: nnz of them)
        uint32_t* tok_out = to.coefs.tail((size_t)nnz);
        int n_tok = 0;
        for (int k = nnz - 1; k >= 0; k--) {
            const int pos = nz[k];
            const int off = pos + ((pos >> bwl) << 2);
            int level = lv[off];
            lv[off] = 0;
            lv3[off] = 0;
            int sign;
            if (pos == 0) {
                int dcs = 0;
                for (int i = 0; i < w4; i++)
                    if (x4 + i < max_x4) {
                        const int s = above_dc[plane][x4 + i];
                        if (s == 1) dcs--;
                        else if (s == 2) dcs++;
                    }
                for (int i = 0; i < h4; i++)
                    if (y4 + i < max_y4) {
                        const int s = left_dc[plane][y4 + i];
                        if (s == 1) dcs--;
                        else if (s == 2) dcs++;
                    }
                const int dctx = dcs < 0 ? 1 : (dcs > 0 ? 2 : 0);
                sign = m.symbol(cdf.dc_sign[ptype][dctx], 2);
            } else {
                sign = m.bit();
            }
            if (level > 14) {
                int length = 0, bit;
                do {
                    length++;
                    bit = m.bit();
                    if (length > 32) {
                        for (int kk = k - 1; kk >= 0; kk--) {
                            const int o2 = nz[kk] + ((nz[kk] >> bwl) << 2);
                            lv[o2] = 0;
                            lv3[o2] = 0;
                        }
                        ms = m;
                        fail(AV1R_EBITSTREAM, "golomb too long");
                        return 0;
                    }
                } while (!bit);
                int x = 1;
                for (int i = length - 2; i >= 0; i--) x = (x << 1) + m.bit();
                level = x + 14;
            }
            if (pos == 0) dc_category = sign ? 1 : 2;
            level &= 0xFFFFF;
            cul_level += level;
            if (cul_level > 63) cul_level = 63;
            tok_out[n_tok++] = coef_token(pos, sign ? -level : level);
        }
        to.coefs.n += (size_t)n_tok;
        to.coef_tokens += eob;
    }
    for (int i = 0; i < w4; i++) {
        above_level[plane][x4 + i] = (uint8_t)cul_level;
        above_dc[plane][x4 + i] = (uint8_t)dc_category;
    }
    for (int i = 0; i < h4; i++) {
        left_level[plane][y4 + i] = (uint8_t)cul_level;
        left_dc[plane][y4 + i] = (uint8_t)dc_category;
    }
    ms = m;
    return eob;
}

}  // namespace av1r

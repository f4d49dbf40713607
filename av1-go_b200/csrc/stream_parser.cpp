#include "stream_parser.h"

#include <chrono>
#include <cstddef>

#include "../../include/av1r.h"
#include "tile.h"

namespace av1r {

void cdf_clear_counters(CdfCtx& c) {
    // every cdf is an array [nsyms+1]; the counter sits right after the *actual* symbols.
#define CLR(field, ns)                                                                  \
    {                                                                                   \
        uint16_t* p = reinterpret_cast<uint16_t*>(&c.field);                            \
        const size_t stride = sizeof(c.field) / sizeof(uint16_t);                       \
        (void)stride;                                                                   \
        p[ns] = 0;                                                                      \
    }
    uint16_t* base = reinterpret_cast<uint16_t*>(&c);
    auto clr_array = [&](size_t off_bytes, size_t total_bytes, int stride, auto ns_of) {
        uint16_t* p = base + off_bytes / 2;
        size_t n = total_bytes / 2 / stride;
        for (size_t k = 0; k < n; k++) p[k * stride + ns_of(k)] = 0;
    };
#define ARR(field, stride, ns_expr) \
    clr_array(offsetof(CdfCtx, field), sizeof(c.field), stride, [&](size_t k) -> int { (void)k; return ns_expr; })
    ARR(txb_skip, 3, 2); ARR(eob_extra, 3, 2); ARR(dc_sign, 3, 2);
    ARR(eob_pt_16, 6, 5); ARR(eob_pt_32, 7, 6); ARR(eob_pt_64, 8, 7); ARR(eob_pt_128, 9, 8);
    ARR(eob_pt_256, 10, 9); ARR(eob_pt_512, 11, 10); ARR(eob_pt_1024, 12, 11);
    ARR(coeff_base_eob, 4, 3); ARR(coeff_base, 5, 4); ARR(coeff_br, 5, 4);
    ARR(newmv, 3, 2); ARR(zeromv, 3, 2); ARR(refmv, 3, 2); ARR(drl, 3, 2);
    ARR(inter_compound_mode, 9, 8); ARR(compound_type, 3, 2); ARR(wedge_idx, 17, 16);
    ARR(interintra, 3, 2); ARR(wedge_interintra, 3, 2); ARR(interintra_mode, 5, 4);
    ARR(motion_mode, 4, 3); ARR(obmc, 3, 2);
    ARR(palette_y_size, 8, 7); ARR(palette_uv_size, 8, 7);
    ARR(palette_y_color_index, 9, (int)(k / 5) + 2); ARR(palette_uv_color_index, 9, (int)(k / 5) + 2);
    ARR(palette_y_mode, 3, 2); ARR(palette_uv_mode, 3, 2);
    ARR(comp_inter, 3, 2); ARR(single_ref, 3, 2); ARR(comp_ref_type, 3, 2); ARR(uni_comp_ref, 3, 2);
    ARR(comp_ref, 3, 2); ARR(comp_bwdref, 3, 2); ARR(txfm_partition, 3, 2); ARR(compound_index, 3, 2);
    ARR(comp_group_idx, 3, 2); ARR(skip_mode, 3, 2); ARR(skip, 3, 2); ARR(intra_inter, 3, 2);
#define NMV(pfx)                                                                                               \
    ARR(pfx##joints, 5, 4);                                                                                    \
    ARR(pfx##c0_classes, 12, 11); ARR(pfx##c0_class0_fp, 5, 4); ARR(pfx##c0_fp, 5, 4); ARR(pfx##c0_sign, 3, 2); \
    ARR(pfx##c0_class0_hp, 3, 2); ARR(pfx##c0_hp, 3, 2); ARR(pfx##c0_class0, 3, 2); ARR(pfx##c0_bits, 3, 2);    \
    ARR(pfx##c1_classes, 12, 11); ARR(pfx##c1_class0_fp, 5, 4); ARR(pfx##c1_fp, 5, 4); ARR(pfx##c1_sign, 3, 2); \
    ARR(pfx##c1_class0_hp, 3, 2); ARR(pfx##c1_hp, 3, 2); ARR(pfx##c1_class0, 3, 2); ARR(pfx##c1_bits, 3, 2);
    NMV(mv_) NMV(dv_)
    ARR(intrabc, 3, 2); ARR(seg_pred, 3, 2); ARR(seg_spatial, 9, 8);
    ARR(filter_intra, 3, 2); ARR(filter_intra_mode, 6, 5);
    ARR(switchable_restore, 4, 3); ARR(wiener_restore, 3, 2); ARR(sgrproj_restore, 3, 2);
    ARR(y_mode, 14, 13); ARR(uv_mode, 15, k < 13 ? 13 : 14);
    ARR(partition, 11, k < 4 ? 4 : (k < 16 ? 10 : 8));
    ARR(switchable_interp, 4, 3); ARR(kf_y_mode, 14, 13); ARR(angle_delta, 8, 7);
    ARR(tx_size, 4, k < 3 ? 2 : 3);
    ARR(delta_q, 5, 4); ARR(delta_lf_multi, 5, 4); ARR(delta_lf, 5, 4);
    ARR(intra_ext_tx, 17, (k / 52) == 1 ? 7 : ((k / 52) == 2 ? 5 : 16));
    ARR(inter_ext_tx, 17, (k / 4) == 1 ? 16 : ((k / 4) == 2 ? 12 : ((k / 4) == 3 ? 2 : 16)));
    ARR(cfl_sign, 9, 8); ARR(cfl_alpha, 17, 16);
#undef ARR
#undef NMV
#undef CLR
}

StreamParser::StreamParser() {
    for (auto& c : slot_cdf_) cdf_load_defaults(c, 0);
}

int StreamParser::begin_frame(const FrameHdr& fh) {
    cur_ = std::make_shared<FrameWork>();
    cur_->init(hp.seq, fh);
    cur_fh_ = fh;
    tiles_done_ = 0;
    have_frame_ = true;
    if (fh.primary_ref_frame == PRIMARY_REF_NONE) {
        cdf_load_defaults(cur_init_cdf_, fh.base_q_idx);
    } else {
        cur_init_cdf_ = slot_cdf_[fh.ref_frame_idx[fh.primary_ref_frame]];
    }
    if (hp.seq.mono_chrome) return fail(AV1R_ENOSYS, "monochrome streams are not supported yet");
    if (fh.use_superres) return fail(AV1R_ENOSYS, "super-resolution is not supported yet");
    if (fh.frame_width != hp.seq.max_frame_width || fh.frame_height != hp.seq.max_frame_height) {
        // reference scaling is not implemented; intra frames of a different size would still work
        if (!fh.frame_is_intra) return fail(AV1R_ENOSYS, "scaled reference frames are not supported yet");
    }
    return 0;
}

int StreamParser::tile_group(const uint8_t* payload, size_t size, size_t offset) {
    BitReader br(payload + offset, size - offset);
    TileGroupInfo tg;
    if (!hp.parse_tile_group_header(br, cur_fh_, tg)) return fail(AV1R_EBITSTREAM, hp.error);
    size_t pos = offset + tg.data_offset;
    auto t0 = std::chrono::steady_clock::now();
    for (int tile = tg.tg_start; tile <= tg.tg_end; tile++) {
        const int tile_row = tile / cur_fh_.tile_cols, tile_col = tile % cur_fh_.tile_cols;
        size_t tile_size;
        if (tile == tg.tg_end) {
            tile_size = size - pos;
        } else {
            if (pos + cur_fh_.tile_size_bytes > size) return fail(AV1R_EBITSTREAM, "truncated tile size");
            tile_size = 0;
            for (int i = 0; i < cur_fh_.tile_size_bytes; i++) tile_size |= (size_t)payload[pos + i] << (8 * i);
            tile_size += 1;
            pos += cur_fh_.tile_size_bytes;
        }
        if (pos + tile_size > size) return fail(AV1R_EBITSTREAM, "tile exceeds tile group");
        TileDecoder td(hp.seq, hp, *cur_, cur_init_cdf_);
        int rc = td.decode_tile(payload + pos, tile_size, tile_row, tile_col);
        if (rc) return fail(rc, td.err);
        if (tile == cur_fh_.context_update_tile_id) {
            cur_->end_cdf = td.cdf;
            cur_->have_end_cdf = true;
        }
        pos += tile_size;
        tiles_done_++;
    }
    cur_->parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

int StreamParser::finish_frame(int64_t pts, std::vector<ParsedFrame>& out) {
    auto t0 = std::chrono::steady_clock::now();
    build_loopfilter_edges(hp.seq, *cur_);
    cur_->parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CdfCtx save = cur_init_cdf_;
    if (!cur_fh_.disable_frame_end_update_cdf && cur_->have_end_cdf) {
        save = cur_->end_cdf;
        cdf_clear_counters(save);
    }
    for (int i = 0; i < NUM_REF_FRAMES; i++)
        if ((cur_fh_.refresh_frame_flags >> i) & 1) {
            slot_cdf_[i] = save;
            slot_fw_[i] = cur_;
        }
    hp.reference_update(cur_fh_);
    ParsedFrame pf;
    pf.fw = cur_;
    pf.fh = cur_fh_;
    pf.pts = pts;
    out.push_back(pf);
    cur_.reset();
    have_frame_ = false;
    return 0;
}

int StreamParser::parse_tu(const uint8_t* data, size_t len, int64_t pts, std::vector<ParsedFrame>& out) {
    std::vector<ObuUnit> obus;
    if (!hp.split_obus(data, len, obus)) return fail(AV1R_EBITSTREAM, hp.error);
    for (const ObuUnit& u : obus) {
        switch (u.type) {
            case OBU_SEQUENCE_HEADER:
                if (!hp.parse_sequence_header(u.data, u.size)) return fail(AV1R_EBITSTREAM, hp.error);
                break;
            case OBU_TEMPORAL_DELIMITER:
                break;
            case OBU_FRAME_HEADER:
            case OBU_REDUNDANT_FRAME_HEADER:
            case OBU_FRAME: {
                if (have_frame_) {
                    if (u.type == OBU_FRAME) return fail(AV1R_EBITSTREAM, "frame OBU while another frame is open");
                    break;   // redundant copy of the active header
                }
                BitReader br(u.data, u.size);
                FrameHdr fh;
                if (!hp.parse_frame_header(br, fh, u.temporal_id, u.spatial_id)) return fail(AV1R_EBITSTREAM, hp.error);
                if (fh.show_existing_frame) {
                    ParsedFrame pf;
                    pf.show_existing_slot = fh.frame_to_show_map_idx;
                    pf.fh = fh;
                    pf.pts = pts;
                    if (fh.frame_type == KEY_FRAME) {
                        // frame loading process (7.21): the shown key frame refreshes every slot
                        const int s = fh.frame_to_show_map_idx;
                        RefHdrState r = hp.refs[s];
                        CdfCtx c = slot_cdf_[s];
                        auto f = slot_fw_[s];
                        for (int i = 0; i < NUM_REF_FRAMES; i++) {
                            hp.refs[i] = r;
                            slot_cdf_[i] = c;
                            slot_fw_[i] = f;
                        }
                    }
                    out.push_back(pf);
                    break;
                }
                int rc = begin_frame(fh);
                if (rc) return rc;
                if (u.type == OBU_FRAME) {
                    br.byte_align();
                    rc = tile_group(u.data, u.size, br.byte_pos());
                    if (rc) return rc;
                    if (tiles_done_ == cur_fh_.tile_cols * cur_fh_.tile_rows) {
                        rc = finish_frame(pts, out);
                        if (rc) return rc;
                    }
                }
                break;
            }
            case OBU_TILE_GROUP: {
                if (!have_frame_) return fail(AV1R_EBITSTREAM, "tile group without frame header");
                int rc = tile_group(u.data, u.size, 0);
                if (rc) return rc;
                if (tiles_done_ == cur_fh_.tile_cols * cur_fh_.tile_rows) {
                    rc = finish_frame(pts, out);
                    if (rc) return rc;
                }
                break;
            }
            case OBU_TILE_LIST:
                return fail(AV1R_ENOSYS, "large-scale tile lists are not supported");
            default:
                break;   // metadata, padding
        }
    }
    return 0;
}

void build_loopfilter_edges(const SeqHdr& seq, FrameWork& fw) {
    const FrameHdr& fh = fw.fh;
    if (!fh.lf.level[0] && !fh.lf.level[1]) return;
    for (int plane = 0; plane < seq.num_planes; plane++) {
        if (plane > 0 && !fh.lf.level[1 + plane]) continue;
        const int sx = plane ? seq.subsampling_x : 0, sy = plane ? seq.subsampling_y : 0;
        const int pw4 = fw.plane_w4(plane), ph4 = fw.plane_h4(plane);
        auto level_of = [&](const BlockInfo* b, int pass) -> int {
            const int i = plane == 0 ? pass : plane + 1;
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
            return lvl;
        };
        for (int r4 = 0; r4 < ph4; r4++)
            for (int c4 = 0; c4 < pw4; c4++) {
                // luma mi position visited by the spec's loop for this plane unit
                const int row = r4 << sy, col = c4 << sx;
                const int x = col * 4, y = row * 4;
                if (row >= fw.mi_rows || col >= fw.mi_cols) continue;
                LfEdge& e = fw.lf[plane][(size_t)r4 * pw4 + c4];
                e = LfEdge{0, 0, 0, 0};
                if (x >= fh.frame_width || y >= fh.frame_height) continue;
                // for sub-sampled planes the spec addresses mode info at the odd (bottom-right) luma mi
                const int mrow = std::min(fw.mi_rows - 1, row | sy), mcol = std::min(fw.mi_cols - 1, col | sx);
                const BlockInfo* b = fw.mi[(size_t)mrow * fw.mi_cols + mcol];
                if (!b) continue;
                const int txsz = fw.lf_tx[plane][(size_t)r4 * pw4 + c4];
                const int psz = plane_residual_size((BlockSize)b->bsize, sx, sy);
                const int xp = c4 * 4, yp = r4 * 4;
                const int is_intra = b->ref_frame[0] <= INTRA_FRAME;
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 0 && c4 == 0) continue;
                    if (pass == 1 && r4 == 0) continue;
                    const int is_block_edge = pass == 0 ? (xp % kBlockW[psz] == 0) : (yp % kBlockH[psz] == 0);
                    const int is_tx_edge = pass == 0 ? (xp % kTxW[txsz] == 0) : (yp % kTxH[txsz] == 0);
                    if (!is_tx_edge) continue;
                    if (!(is_block_edge || !b->skip || is_intra)) continue;
                    const int pr4 = pass == 0 ? r4 : r4 - 1, pc4 = pass == 0 ? c4 - 1 : c4;
                    const int prev_tx = fw.lf_tx[plane][(size_t)pr4 * pw4 + pc4];
                    const int base = pass == 0 ? std::min(kTxW[prev_tx], kTxW[txsz]) : std::min(kTxH[prev_tx], kTxH[txsz]);
                    const int fsz = plane == 0 ? std::min(16, base) : std::min(8, base);
                    int lvl = level_of(b, pass);
                    if (lvl == 0) {
                        const int prow = std::min(fw.mi_rows - 1, (pr4 << sy) | sy), pcol = std::min(fw.mi_cols - 1, (pc4 << sx) | sx);
                        const BlockInfo* pb = fw.mi[(size_t)prow * fw.mi_cols + pcol];
                        if (pb) lvl = level_of(pb, pass);
                    }
                    if (lvl == 0) continue;
                    if (pass == 0) { e.len_v = (uint8_t)fsz; e.lvl_v = (uint8_t)lvl; }
                    else { e.len_h = (uint8_t)fsz; e.lvl_h = (uint8_t)lvl; }
                }
            }
    }
}

}  // namespace av1r
